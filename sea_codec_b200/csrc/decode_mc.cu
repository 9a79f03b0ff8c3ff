// decode_mc.cu -- throughput decode for uniform CBR batches with MORE than two channels (4, 6 or 8; BASELINE config 3 is an
// 8-channel stream): decode_mc_kernel<CT, B>.
//
// Lane mapping: one lane per chunk with ALL its channels (MCfg::CPL == CT; the first versions gave a lane one channel pair, then
// a quad: see MCfg).  In the [frame][channel] bit stream (chunk.rs:254-278) a frame's CT codes are adjacent, so the lane walks
// the residual section front to back with CT independent LMS chains in registers (the ILP of the stereo kernel and more) and
// owns whole PCM frames: they leave as full 32-byte sectors.  Everything else is borrowed from decode_unrolled_kernel /
// decode_vbr_kernel: per-lane cp.async ring, a window of big-endian words per body pre-shifted once so that every field
// position inside the body is a compile-time constant, I2IP pack-saturate clamp, LMS signs carried in registers.
#include "sea_device.cuh"

namespace sea {

using namespace dev;

namespace {


template <int V>
struct ParTag {
    static constexpr int value = V;
};

template <int CT, int B>
struct MCfg {
    static constexpr int F = 20;
#ifndef SEA_MC_WHOLE
#define SEA_MC_WHOLE 1
#endif
    // Channels per lane.  SEA_MC_WHOLE (default): all of them -- the lane owns whole frames, so its PCM goes out as full 16/32-byte
    // stores (a pair or a quad per lane wrote 4/8 bytes to 32 different rows per instruction: the L1 tag stage, not the math,
    // bounded 6 channels at 0.67 and 8 at 0.93 Tsamples/s) and CT independent LMS chains give the lane its ILP.  Otherwise a quad
    // where the count allows, else a pair (the first version of this kernel).
    static constexpr int CPL = SEA_MC_WHOLE ? CT : ((CT % 4 == 0) ? 4 : 2);
    static constexpr int U = CT / CPL;                  // lanes per chunk
    static constexpr int kChunksPerWarp = 32 / U;       // pairs of CT = 6: 10 chunks, two idle lanes
    // frames per looped body (divides F): bounded by the window registers (body bits / 32) and, for whole frames, a body must
    // be a whole number of stores
#ifndef SEA_MC_HF8
#define SEA_MC_HF8 10
#endif
    static constexpr int HF = CPL != CT ? (CT >= 6 ? 10 : 20) : (CT == 4 ? 20 : (CT == 6 ? 4 : (B <= 4 ? SEA_MC_HF8 : 4)));
    static constexpr int WPF = CT / 2;                  // 32-bit words per frame
    // Whole frames go out as 32-byte stores.  A body of 6 channels is an odd number of 16-byte halves (240 or 48 bytes), so its
    // bodies alternate between two store phases (kPhaseWords = 4): the last four words of an even body wait in registers for
    // the first four of the odd one.  The host only sends 6-channel batches here when N % 40 == 0 (even body count, 32-byte rows).
    static constexpr int kPhaseWords = (HF * WPF) % 8;
    static_assert(CPL != CT || kPhaseWords == 0 || kPhaseWords == 4, "a body must be a whole number of 16-byte halves");
    static constexpr int kBodyBits = HF * CT * B;       // bits of the stream one body walks through
    static constexpr int kNW = (kBodyBits - (CT - CPL) * B + 31 + 31) / 32;  // window words from my first field to my last (any phase)
    static constexpr int kBodyBytesMax = (kBodyBits + 7) / 8 + 1;
    static constexpr int kRingWords = 64;               // 256-byte ring per lane: two bodies (<= 80 bytes each) plus slack
    static constexpr int kTopUp = (kBodyBytesMax + 15) / 16 + 1;  // granules issued per body at most
    // The ring is kept full, so the bytes of a body were issued (256 - 32) / body bytes - 1 bodies before it is decoded: that many
    // of the newest groups may still be in flight.  (Waiting for all but the newest one stalled every body on loads it would
    // not need for another 4-5 bodies: long_scoreboard 1.35 per issue in profiles/r01_decode_mc_v2.)
    static constexpr int kAhead = (256 - 32) / kBodyBytesMax - 1;
    static constexpr int kKeep = kAhead < 1 ? 1 : (kAhead > 4 ? 4 : kAhead);
    static constexpr int kPitch = 256 + 16;
    static constexpr int kWarpBytes = 32 * kPitch + 64;
#ifndef SEA_MC_WARPS8
#define SEA_MC_WARPS8 12
#endif
    static constexpr int kWarps = CT == 4 ? 16 : (CPL == CT ? SEA_MC_WARPS8 : 12);  // measured: 4 channels 2.14 ms at 16 warps (2.46 at 12); quads of 8: 7.36 ms at 12 (8.75 at 16)
    static_assert(2 * kBodyBytesMax + 32 <= 256, "ring too small for two bodies");
};

}  // namespace

template <int CT, int B>
__global__ void __launch_bounds__(MCfg<CT, B>::kWarps * 32, 1)
decode_mc_kernel(const uint8_t *__restrict__ sea, int16_t *__restrict__ pcm, const DecStream *__restrict__ streams, DecFastParams p,
                 const int32_t *__restrict__ tab, int *err)
{
    using Cfg = MCfg<CT, B>;
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t s = 4;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

    // dequant rows of size B as uploaded: lut[sf][code], at the start of the (1024-byte aligned) shared window so that a row's
    // base address has its low B + 2 bits clear and "row | code << 2" needs no add
    const uint32_t nwarps = blockDim.x >> 5;  // chosen per launch (launch_mc): fewer warps per CTA when the grid is only a few waves
    const uint32_t smem_sh = smem_u32(smem), lut_sh = (smem_sh + 1023u) & ~1023u;
    int32_t *lut = reinterpret_cast<int32_t *>(smem + (lut_sh - smem_sh));
    for (uint32_t i = threadIdx.x; i < (1u << (s + B)); i += blockDim.x) lut[i] = tab[tab_dqt_off(s, B) + i];
    __syncthreads();
    const uint32_t rings_off = (lut_sh - smem_sh) + (4u << (s + B));

    constexpr int CPL = Cfg::CPL;
    const uint32_t pr = lane % Cfg::U;                              // my channel group (pair or quad)
    uint64_t g = ((uint64_t)blockIdx.x * nwarps + warp) * Cfg::kChunksPerWarp + lane / Cfg::U;  // global chunk index
    const bool valid = lane < (uint32_t)(Cfg::kChunksPerWarp * Cfg::U) && g < p.total_chunks;
    if (!valid) g = p.total_chunks - 1;  // idle lanes shadow the last chunk and never store

    const DecStream st = streams[find_stream(streams, p.n_streams, g * CT)];
    const uint32_t k = (uint32_t)(g - st.chain_begin / CT);
    const uint64_t ck_off = st.data_off + (uint64_t)k * p.chunk_size;
    const uint8_t *ck = sea + ck_off;
    {
        const uint32_t word = (uint32_t)ck[0] | ((uint32_t)ck[1] << 8) | ((uint32_t)ck[2] << 16) | ((uint32_t)ck[3] << 24);
        if (word != p.hdr_word) report(err, kDevFallback);  // not what this kernel was specialised for: host reruns generically
    }
    int32_t w[CPL][4], h[CPL][4], sg[CPL][4];
#pragma unroll
    for (int c = 0; c < CPL; c++) {
        const uint8_t *l = ck + 4u + 16u * (CPL * pr + c);  // lms.rs:80-94
#pragma unroll
        for (int i = 0; i < 4; i++) {
            h[c][i] = (int16_t)(l[2 * i] | (l[2 * i + 1] << 8));
            w[c][i] = (int16_t)(l[8 + 2 * i] | (l[8 + 2 * i + 1] << 8));
            sg[c][i] = (h[c][i] >> 31) | 1;
        }
    }
    const uint32_t items = (p.N / Cfg::F) * CT;
    const uint64_t sf_off = ck_off + 4u + 16u * CT;              // chunk.rs:108-113
    const uint64_t res_off = sf_off + items / 2u;                // s == 4: two scale factors per byte, items even
    const uint8_t *sfp = sea + sf_off + pr * (CPL / 2);          // my group's CPL/2 bytes of block b: sfp[b * CT/2 ...]
    uint8_t *out = reinterpret_cast<uint8_t *>(pcm + st.pcm_off + (uint64_t)k * p.N * CT) + 2u * CPL * pr;

    // ---- per-lane ring.  Word w of the 16-byte aligned stream sits at ring word (w & 63).
    const uint64_t a0 = res_off & ~(uint64_t)15;
    const uint8_t *src0 = sea + a0;
    const uint32_t ring_sh = smem_u32(smem + rings_off + warp * Cfg::kWarpBytes) + lane * Cfg::kPitch + (lane >> 3) * 16u;
    uint32_t fetched = 0;                                        // granules issued so far
    uint32_t posg = (uint32_t)(res_off - a0) * 8u;               // bit position of the current body's first field, from a0
#pragma unroll
    for (int t = 0; t < 16; t++) cp_async16_if(true, ring_sh + t * 16, src0 + t * 16);
    fetched = 16;
    cp_async_commit();
    cp_async_commit();
    cp_async_wait<0>();

    const uint32_t n_bodies = p.N / Cfg::HF;
    constexpr int kBodiesPerBlock = Cfg::F / Cfg::HF;
    // scale-factor nibbles of a block for my channels, first channel in the top nibble of the CPL*4-bit value.  The bytes of the
    // NEXT block are requested when a block starts and only combined when the next one does: consumed right after the load
    // (the first version) every block waited out a global-memory round trip -- 22 % of the stall samples of
    // profiles/r01_decode_mc_v3 sat on the shift behind that load.
    uint32_t sf_raw[CPL / 2];
    auto request_sf = [&](uint32_t blk) {
        const uint8_t *q = sfp + (size_t)blk * (CT / 2);
#pragma unroll
        for (int j = 0; j < CPL / 2; j++) asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(sf_raw[j]) : "l"(q + j));
    };
    auto combine_sf = [&]() -> uint32_t {
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < CPL / 2; j++) v = (v << 8) | sf_raw[j];
        return v;
    };
    request_sf(0);
    uint32_t sf_cur = 0;
    const uint32_t n_blocks = p.N / Cfg::F;

    uint32_t ow[8];  // whole frames: the 32-byte store being assembled (carried across bodies when kPhaseWords != 0)
    auto body = [&](uint32_t bd, auto parity_tag) {
        constexpr int kPar = decltype(parity_tag)::value;
        // ---- top the ring up, then wait for everything but that (the bytes of this body were issued a body ago)
        {
            const uint32_t wq = posg >> 5;
#pragma unroll
            for (int t = 0; t < Cfg::kTopUp; t++) {
                const bool room = fetched * 4u + 4u <= wq + (uint32_t)Cfg::kRingWords;
                cp_async16_if(room, ring_sh + (fetched & 15u) * 16u, src0 + (size_t)fetched * 16u);
                fetched += room ? 1u : 0u;
            }
            cp_async_commit();
            cp_async_wait<Cfg::kKeep>();
        }
        // scale factors of this body's block (one byte per block and pair); the next block's byte is fetched a body ahead
        if ((bd % kBodiesPerBlock) == 0) {
            const uint32_t blk = bd / kBodiesPerBlock;
            sf_cur = combine_sf();
            if (blk + 1u < n_blocks) request_sf(blk + 1u);
        }
        uint32_t rowbase[CPL];
#pragma unroll
        for (int c = 0; c < CPL; c++) rowbase[c] = lut_sh + (((sf_cur >> (4 * (CPL - 1 - c))) & 15u) << (B + 2));

        // ---- window: big-endian words from my first field of this body on, pre-shifted so that it starts at bit 0 of W[0]
        const uint32_t my = posg + (uint32_t)(CPL * B) * pr;
        const uint32_t w0 = my >> 5, sh = my & 31u;
        uint32_t V[Cfg::kNW + 1], W[Cfg::kNW];
#pragma unroll
        for (int t = 0; t < Cfg::kNW + 1; t++) V[t] = __byte_perm(lds_u32(ring_sh + ((w0 + t) & 63u) * 4u), 0, 0x0123);
#pragma unroll
        for (int t = 0; t < Cfg::kNW; t++) W[t] = __funnelshift_l(V[t + 1], V[t], sh);
        posg += Cfg::kBodyBits;

        uint8_t *ob = out + (size_t)bd * (Cfg::HF * CT * 2);
#pragma unroll
        for (int fi = 0; fi < Cfg::HF; fi++) {
            constexpr int kGB = CPL * B;
            const int bit = fi * CT * B;  // compile-time position of my group of codes in W[]
            const int wd = bit >> 5, off = bit & 31;
            uint32_t x = 0;  // my CPL codes in the low CPL*B bits, first channel highest (groups of up to 32 bits)
            if (kGB <= 32) {
                if (off + kGB <= 32) x = W[wd] >> (32 - off - kGB);
                else x = __funnelshift_r(W[wd + 1], W[wd], (64 - off - kGB) & 31);
            }
            int32_t y[CPL], d[CPL], sgn[CPL];
#pragma unroll
            for (int c = 0; c < CPL; c++) {
                // the code lands at bit 2 (the table's 4-byte stride) in ONE shift; mask and row base join it in one LOP3
                uint32_t code4;
                constexpr uint32_t kMask4 = ((1u << B) - 1u) << 2;
                if (kGB <= 32) {
                    const int sh2 = B * (CPL - 1 - c) - 2;
                    code4 = sh2 >= 0 ? (x >> (sh2 & 31)) : (x << ((-sh2) & 31));
                } else {  // wider groups: every field on its own, still at a compile-time position
                    const int cb = bit + c * B, cw = cb >> 5, co = cb & 31;
                    if (co + B + 2 <= 32) code4 = W[cw] >> (32 - co - B - 2);
                    else if (co + B <= 32) code4 = W[cw] << ((co + B + 2 - 32) & 31);
                    else code4 = __funnelshift_r(W[cw + 1], W[cw], (64 - co - B - 2) & 31);
                }
                d[c] = lds_s32((code4 & kMask4) | rowbase[c]);
                const uint32_t acc = (uint32_t)w[c][0] * (uint32_t)h[c][0] + (uint32_t)w[c][1] * (uint32_t)h[c][1] +
                                     (uint32_t)w[c][2] * (uint32_t)h[c][2] + (uint32_t)w[c][3] * (uint32_t)h[c][3];
                y[c] = (int32_t)((uint32_t)((int32_t)acc >> 13) + (uint32_t)d[c]);  // codec/decoder.rs:38, before the clamp
                sgn[c] = (y[c] >> 31) | 1;                                            // the clamp keeps the sign
            }
            uint32_t packed[CPL / 2];
#pragma unroll
            for (int q = 0; q < CPL / 2; q++) {  // clamp_i16 x2 + interleave
                asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(packed[q]) : "r"(y[2 * q + 1]), "r"(y[2 * q]));
                y[2 * q] = (int32_t)(int16_t)(packed[q] & 0xffffu);
                y[2 * q + 1] = (int32_t)packed[q] >> 16;
            }
#pragma unroll
            for (int c = 0; c < CPL; c++) {
                const int32_t delta = d[c] >> 4;  // lms.rs:43-51
                w[c][0] += delta * sg[c][0];
                w[c][1] += delta * sg[c][1];
                w[c][2] += delta * sg[c][2];
                w[c][3] += delta * sg[c][3];
                h[c][0] = h[c][1]; h[c][1] = h[c][2]; h[c][2] = h[c][3]; h[c][3] = y[c];
                sg[c][0] = sg[c][1]; sg[c][1] = sg[c][2]; sg[c][2] = sg[c][3]; sg[c][3] = sgn[c];
            }
            if (Cfg::U == 1) {
                // the lane owns whole frames: consecutive frames fill 32-byte sectors -> 256-bit stores as in the stereo kernel
                // (128-bit where a body is not a multiple of 32 bytes: 6 channels); 8-byte stores to 32 different rows per
                // instruction choked the L1 tag stage (26 % issue)
#pragma unroll
                for (int q = 0; q < Cfg::WPF; q++) {
                    const int widx = kPar * Cfg::kPhaseWords + fi * Cfg::WPF + q;  // word position counted from the last 32-byte boundary before the body
                    ow[widx % 8] = packed[q];
                    if (widx % 8 == 7 && valid) {
                        uint8_t *dst = ob + (widx / 8) * 32 - kPar * Cfg::kPhaseWords * 4;
                        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(ow[0]), "r"(ow[1]), "r"(ow[2]),
                                     "r"(ow[3]), "r"(ow[4]), "r"(ow[5]), "r"(ow[6]), "r"(ow[7])
                                     : "memory");
                    }
                }
            } else if (valid) {
                if (CPL == 4) *reinterpret_cast<uint2 *>(ob + fi * (CT * 2)) = make_uint2(packed[0], packed[CPL / 2 - 1]);
                else *reinterpret_cast<uint32_t *>(ob + fi * (CT * 2)) = packed[0];
            }
        }
    };
    if (Cfg::U == 1 && Cfg::kPhaseWords != 0) {
        for (uint32_t bd = 0; bd < n_bodies; bd += 2) {
            body(bd, ParTag<0>{});
            body(bd + 1u, ParTag<1>{});
        }
    } else {
        for (uint32_t bd = 0; bd < n_bodies; bd++) body(bd, ParTag<0>{});
    }
}

bool decode_mc_supported(const DecFastParams &p)
{
    if (p.channels != 4 && p.channels != 6 && p.channels != 8) return false;
    if ((p.hdr_word & 0xffu) != 1u) return false;  // CBR chunks only
    if (p.F != 20 || p.s != 4 || p.b < 1 || p.b > 8) return false;
    if (p.channels == 6 && p.N % 40u != 0) return false;  // 12-byte frames: 32-byte rows and an even number of bodies (see MCfg)
    return p.N != 0 && p.N % 20u == 0;
}

template <int CT, int B>
static cudaError_t launch_mc(const uint8_t *d_sea, int16_t *d_pcm, const DecStream *d_streams, const DecFastParams &p, const int32_t *tab,
                             int *d_err, cudaStream_t stream)
{
    using Cfg = MCfg<CT, B>;
    const uint32_t warps = pick_cta_warps(p.total_chunks, Cfg::kChunksPerWarp, Cfg::kWarps);
    const size_t smem = (size_t)warps * Cfg::kWarpBytes + ((size_t)4u << (4 + B)) + 1024u;
    cudaError_t e = cudaFuncSetAttribute(decode_mc_kernel<CT, B>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)((size_t)Cfg::kWarps * Cfg::kWarpBytes + ((size_t)4u << (4 + B)) + 1024u));
    if (e != cudaSuccess) return e;
    const uint64_t chunks_per_cta = (uint64_t)warps * Cfg::kChunksPerWarp;
    const uint64_t blocks = (p.total_chunks + chunks_per_cta - 1) / chunks_per_cta;
    decode_mc_kernel<CT, B><<<(unsigned)blocks, warps * 32, smem, stream>>>(d_sea, d_pcm, d_streams, p, tab, d_err);
    return cudaGetLastError();
}

template <int CT>
static cudaError_t launch_mc_b(const uint8_t *d_sea, int16_t *d_pcm, const DecStream *d_streams, const DecFastParams &p, const int32_t *tab,
                               int *d_err, cudaStream_t stream)
{
    switch (p.b) {
        case 1: return launch_mc<CT, 1>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        case 2: return launch_mc<CT, 2>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        case 3: return launch_mc<CT, 3>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        case 4: return launch_mc<CT, 4>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        case 5: return launch_mc<CT, 5>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        case 6: return launch_mc<CT, 6>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        case 7: return launch_mc<CT, 7>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        default: return launch_mc<CT, 8>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
    }
}

cudaError_t launch_decode_mc(const uint8_t *d_sea, int16_t *d_pcm, const DecStream *d_streams, const DecFastParams &p, DevTables tabs,
                             int *d_err, cudaStream_t stream)
{
    if (p.total_chunks == 0) return cudaSuccess;
    const int32_t *tab = tabs.by_s[p.s];
    switch (p.channels) {
        case 4: return launch_mc_b<4>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        case 6: return launch_mc_b<6>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        default: return launch_mc_b<8>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
    }
}

}  // namespace sea
