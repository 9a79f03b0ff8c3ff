// decode_fast.cu -- the throughput decode kernel for uniform CBR batches (1 or 2 channels, scale_factor_frames = 20,
// full chunks): decode_unrolled_kernel<C, B>.
//
// Same work as decode_staged_kernel (chunk.rs:69-213 parse, bits.rs:34-50 unpack, codec/decoder.rs:20-50 reconstruct), laid
// out for the B200 issue-rate bound (the chain recurrence costs ~18 integer instructions per sample; HBM needs 2.4 B/sample):
//   * one chain (chunk, channel) per lane, a warp owns 32/C consecutive chunks, a CTA of 32 warps owns one SM;
//   * packed residuals arrive by 16-byte cp.async (LDGSTS) issued by the lanes of each chunk row, double buffered one round
//     ahead; PCM leaves as whole interleaved row tiles with 128-bit shared loads / global stores.  (Round 1 first used TMA bulk
//     copies for both directions: 48 bulk ops per warp-round, each serialised through the uniform datapath by an
//     elect/broadcast loop, cost ~6 extra ALU-pipe instructions per sample; see profiles/r01_decode_unrolled_tma_v2_*.)
//   * a round is RF frames with RF*C*B a multiple of 32 bits, so every field position inside a round is a compile-time
//     constant: a field costs one shift and one LOP3 that also forms the look-up address;
//   * the dequant row table is replicated per bank in shared memory (lane l reads bank l) when it fits, so the one dependent
//     shared-memory load per sample is conflict free.
#include "sea_kernels.h"

namespace sea {

namespace {

__device__ __forceinline__ void report_f(int *err, int code) { atomicCAS(err, 0, code); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ int32_t lds_s32(uint32_t addr)
{
    int32_t v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint32_t find_stream_f(const DecStream *streams, uint32_t n_streams, uint64_t chain)
{
    uint32_t lo = 0, hi = n_streams;  // last stream whose chain_begin <= chain
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((uint64_t)streams[mid].chain_begin <= chain) lo = mid;
        else hi = mid;
    }
    return lo;
}

constexpr int round_up16(int v) { return (v + 15) / 16 * 16; }
constexpr int pitch16_odd(int bytes)  // multiple of 16 with an odd number of 16-byte units: rows spread over all bank groups
{
    int p = round_up16(bytes);
    return ((p / 16) % 2) ? p : p + 16;
}

template <int C, int B>
struct UCfg {
    static constexpr int F = 20;                                   // scale_factor_frames this kernel is unrolled for
    static constexpr int kRows = 32 / C;                           // chunks per warp
    static constexpr int RF = ((C * B) % 2 == 0) ? 80 : 160;       // frames per round: RF*C*B % 32 == 0 and RF % F == 0
    static constexpr int kRoundBits = RF * C * B;
    static constexpr int kRoundBytes = kRoundBits / 8;
    static constexpr int HF = 40;                                  // frames per output tile (two blocks)
    static constexpr int kHalves = RF / HF;
    static constexpr int kHalfBits = HF * C * B;
    static constexpr int kNW = ((kHalfBits + 8 + 31) >> 5) + 1;    // words one half can touch (channel shift + straddle)
    // row buffer: up to 12 bytes of 16-byte alignment slack, then every word the last half reads (realign + funnel over-read)
    // input tile: word-major [word][row] so that "word w of every row" is one conflict-free wavefront; the words come in by
    // 4-byte cp.async from the 4-byte aligned start of the round (byte phase is undone by the PRMT that also swaps bytes)
    static constexpr int kInWords = (((kHalves - 1) * kHalfBits) >> 5) + kNW + 2;
    static constexpr int kInBytes = kInWords * 4;  // per row and buffer
    // output tile rows are dense (pitch = row bytes: 40 or 20 words, so LDS/STS of 4 or 8 consecutive rows tile all 32 banks);
    // the 4 rows that would share a bank group rotate the words inside each 16-byte granule by rho = (row / kOutPeriod) & 3
    static constexpr int kOutBytes = HF * C * 2;
    static constexpr int kOutPitch = kOutBytes;
    static constexpr int kOutPeriod = C == 2 ? 4 : 8;
    static constexpr int kWarpBytes = 2 * kRows * kInBytes + kRows * kOutPitch;
    // warps per CTA (one CTA per SM): as many as fit next to <= 37 KB of look-up table, a multiple of 4 (one per SMSP)
    static constexpr int kWarpsFit = (190 * 1024 / kWarpBytes) / 4 * 4;
    static constexpr int kWarps = kWarpsFit < 32 ? kWarpsFit : 32;
};

}  // namespace

template <int C, int B, bool REPL>
__global__ void __launch_bounds__(UCfg<C, B>::kWarps * 32, 1)
decode_unrolled_kernel(const uint8_t *__restrict__ sea, int16_t *__restrict__ pcm, const DecStream *__restrict__ streams,
                       DecFastParams p, const int32_t *__restrict__ tab, uint32_t lut_align, int *err)
{
    using Cfg = UCfg<C, B>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t s = p.s;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

    // ---- dequant rows of residual size B: lut[sf][code], replicated per bank when REPL (entry stride 128 B, lane l at +4l)
    constexpr int kShift = REPL ? 7 : 2;  // log2 of the byte stride between consecutive codes
    // the table sits after the warp tiles at an address aligned to its own size, so "row base | code offset" never carries
    const uint32_t smem_sh = smem_u32(smem);
    const uint32_t lut_abs = (smem_sh + Cfg::kWarps * Cfg::kWarpBytes + lut_align - 1u) & ~(lut_align - 1u);
    int32_t *lut = reinterpret_cast<int32_t *>(smem + (lut_abs - smem_sh));
    {
        const uint32_t entries = 1u << (s + B);
        const int32_t *src = tab + tab_dqt_off(s, B);
        if (REPL) {
            for (uint32_t i = threadIdx.x; i < entries * 32u; i += blockDim.x) lut[i] = src[i >> 5];
        } else {
            for (uint32_t i = threadIdx.x; i < entries; i += blockDim.x) lut[i] = src[i];
        }
    }
    __syncthreads();
    const uint32_t lut_sh = smem_u32(lut) + (REPL ? lane * 4u : 0u);

    uint8_t *wbase = smem + warp * Cfg::kWarpBytes;
    uint8_t *in_rows = wbase;  // [2][kInWords][kRows] words
    uint8_t *out_rows = wbase + 2 * Cfg::kRows * Cfg::kInBytes;

    const uint32_t j = lane / C, c = lane % C;
    uint64_t g = ((uint64_t)blockIdx.x * Cfg::kWarps + warp) * Cfg::kRows + j;  // global chunk index
    const bool valid = g < p.total_chunks;
    if (!valid) g = p.total_chunks - 1;  // idle rows shadow the last chunk and never store

    const DecStream st = streams[find_stream_f(streams, p.n_streams, g * C)];
    const uint32_t k = (uint32_t)(g - st.chain_begin / C);
    const uint64_t ck_off = st.data_off + (uint64_t)k * p.chunk_size;
    const uint8_t *ck = sea + ck_off;
    {
        const uint32_t word = (uint32_t)ck[0] | ((uint32_t)ck[1] << 8) | ((uint32_t)ck[2] << 16) | ((uint32_t)ck[3] << 24);
        if (word != p.hdr_word) report_f(err, kDevFallback);  // not what this kernel was specialised for: host reruns generically
    }
    int32_t w[4], h[4], sg[4];
    {
        const uint8_t *l = ck + 4u + 16u * c;  // lms.rs:80-94
#pragma unroll
        for (int i = 0; i < 4; i++) {
            h[i] = (int16_t)(l[2 * i] | (l[2 * i + 1] << 8));
            w[i] = (int16_t)(l[8 + 2 * i] | (l[8 + 2 * i + 1] << 8));
            sg[i] = (h[i] >> 31) | 1;
        }
    }
    const uint32_t items = (p.N / Cfg::F) * C;
    const uint64_t sf_off = ck_off + 4u + 16u * C;
    const uint64_t res_off = sf_off + div_ceil_u32(items * s, 8u);
    const uint8_t *sfp = sea + sf_off;
    int16_t *out = pcm + st.pcm_off + (uint64_t)k * p.N * C;

    // byte phase of the residual section inside its 16-byte granule: constant over rounds modulo 4 (kRoundBytes % 4 == 0)
    const uint32_t bp = (uint32_t)res_off & 3u;
    const uint32_t prmt_sel = 0x0123u + bp * 0x1111u;  // byte swap + byte realign in one PRMT
    const uint32_t cB = c * B;                         // my channel's bit offset inside a frame

    const uint32_t n_rounds = p.N / Cfg::RF;
    constexpr int kBlocks = Cfg::RF / Cfg::F;
    constexpr int kOutGran = Cfg::kOutBytes / 16;
    constexpr int kInPerLane = (Cfg::kInWords + C - 1) / C, kOutPerLane = (kOutGran + C - 1) / C;

    // The C lanes of a row fetch its next slice word by word (4-byte cp.async, lane c takes words c, c+C, ...), one round ahead.
    const uint32_t my_in_sh = smem_u32(in_rows) + (c * Cfg::kRows + j) * 4u;
    auto issue_round = [&](uint32_t r) {
        const uint8_t *src = sea + ((res_off + (uint64_t)r * Cfg::kRoundBytes) & ~(uint64_t)3) + c * 4u;
        const uint32_t dst = my_in_sh + (r & 1u) * (Cfg::kRows * Cfg::kInBytes);
#pragma unroll
        for (int t = 0; t < kInPerLane; t++)
            if (t * C + (int)c < Cfg::kInWords) cp_async4(dst + t * C * Cfg::kRows * 4, src + t * C * 4);
        cp_async_commit();
    };
    // scale-factor bytes are prefetched one round ahead too (they come straight from global memory)
    uint32_t sf_raw[kBlocks], sf_raw2[kBlocks];
    auto fetch_sf = [&](uint32_t r, uint32_t *raw, uint32_t *raw2) {
#pragma unroll
        for (int q = 0; q < kBlocks; q++) {
            raw2[q] = 0;
            if (s == 4u && C == 2) {
                raw[q] = __ldg(sfp + r * kBlocks + q);
            } else if (s == 4u && C == 1) {
                raw[q] = __ldg(sfp + ((r * kBlocks + q) >> 1));
            } else {
                const uint32_t bit = ((r * kBlocks + q) * C + c) * s;
                raw[q] = __ldg(sfp + (bit >> 3));
                if ((bit & 7u) + s > 8u) raw2[q] = __ldg(sfp + (bit >> 3) + 1);
            }
        }
    };
    issue_round(0);
    fetch_sf(0, sf_raw, sf_raw2);

    // my row of the output tile: logical word m of every granule lives at physical word (m + rho) & 3
    const uint32_t rho = (j / Cfg::kOutPeriod) & 3u;
    uint8_t *my_row = out_rows + j * Cfg::kOutPitch;
    uint8_t *st_base[4];       // per logical word-in-granule: where my samples go (stereo: + 2c inside the word)
    const uint32_t *ld_base[4];  // per logical word-in-granule: where the copy-out reads it (my first granule is granule c)
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const int32_t rot = (int32_t)((m + rho) & 3u) - m;
        st_base[m] = my_row + rot * 4 + (C == 2 ? 2 * c : 0);
        ld_base[m] = reinterpret_cast<const uint32_t *>(my_row) + ((m + rho) & 3u) + c * 4u;
    }
    uint4 *my_dst = reinterpret_cast<uint4 *>(out) + c;

    for (uint32_t r = 0; r < n_rounds; r++) {
        // all lanes are done with buffer (r+1)&1 (read in round r-1): refill it, then wait for this round's slice
        __syncwarp();
        if (r + 1 < n_rounds) issue_round(r + 1);
        else cp_async_commit();
        cp_async_wait1();
        __syncwarp();
        const uint32_t *words = reinterpret_cast<const uint32_t *>(in_rows + (r & 1u) * (Cfg::kRows * Cfg::kInBytes)) + j;

        // scale factors of this round's blocks (bytes fetched during the previous round), then prefetch the next round's
        uint32_t sfv[kBlocks];
#pragma unroll
        for (int q = 0; q < kBlocks; q++) {
            if (s == 4u && C == 2) {
                sfv[q] = (sf_raw[q] >> (4u * (1u - c))) & 15u;
            } else if (s == 4u && C == 1) {
                sfv[q] = (sf_raw[q] >> (4u * (1u - (q & 1)))) & 15u;
            } else {
                const uint32_t sh = (((r * kBlocks + q) * C + c) * s) & 7u;
                sfv[q] = (((sf_raw[q] << 8) | sf_raw2[q]) >> (16u - sh - s)) & ((1u << s) - 1u);
            }
        }
        if (r + 1 < n_rounds) fetch_sf(r + 1, sf_raw, sf_raw2);

#pragma unroll
        for (int hh = 0; hh < Cfg::kHalves; hh++) {
            // window of this half: big-endian words, realigned so that my field i sits at bit i*C*B of W[0..]
            constexpr int kHB = Cfg::kHalfBits;
            const int wlo = (hh * kHB) >> 5;
            constexpr int kNW = Cfg::kNW;
            uint32_t V[kNW + 2], W[kNW + 1];
#pragma unroll
            for (int t = 0; t < kNW + 2; t++) V[t] = words[(wlo + t) * Cfg::kRows];
#pragma unroll
            for (int t = 0; t < kNW + 1; t++) W[t] = __byte_perm(V[t], V[t + 1], prmt_sel);
            if (C == 2) {
#pragma unroll
                for (int t = 0; t < kNW; t++) W[t] = __funnelshift_l(W[t + 1], W[t], cB);
            }
            __syncwarp();  // the previous tile has been copied out by every lane of the row
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint32_t sf = sfv[hh * 2 + q];
                const uint32_t rowbase = lut_sh + (sf << (B + kShift));
#pragma unroll
                for (int i = 0; i < Cfg::F; i++) {
                    const int fi = q * Cfg::F + i;                      // frame inside the half
                    const int bit = (hh * kHB) - (wlo << 5) + fi * C * B;  // compile-time position of my field in W[]
                    const int wd = bit >> 5, off = bit & 31;
                    uint32_t x;
                    if (off + B <= 32) {
                        const int rs = 32 - off - B - kShift;
                        x = rs >= 0 ? (W[wd] >> (rs & 31)) : (W[wd] << ((-rs) & 31));
                    } else {
                        x = __funnelshift_r(W[wd + 1], W[wd], (64 - off - B - kShift) & 31);
                    }
                    const uint32_t addr = (x & (((1u << B) - 1u) << kShift)) | rowbase;
                    const int32_t d = lds_s32(addr);
                    const uint32_t acc = (uint32_t)w[0] * (uint32_t)h[0] + (uint32_t)w[1] * (uint32_t)h[1] + (uint32_t)w[2] * (uint32_t)h[2] +
                                         (uint32_t)w[3] * (uint32_t)h[3];
                    const int32_t y = clamp_i16((int32_t)((uint32_t)((int32_t)acc >> 13) + (uint32_t)d));
                    {
                        const int lbyte = (fi * C) * 2;  // logical byte of (frame fi, channel 0); my channel adds 2c (in st_base)
                        *reinterpret_cast<int16_t *>(st_base[(lbyte >> 2) & 3] + (lbyte >> 2) * 4 + (C == 1 ? (lbyte & 2) : 0)) = (int16_t)y;
                    }
                    const int32_t delta = d >> 4;
                    w[0] += delta * sg[0];
                    w[1] += delta * sg[1];
                    w[2] += delta * sg[2];
                    w[3] += delta * sg[3];
                    h[0] = h[1]; h[1] = h[2]; h[2] = h[3]; h[3] = y;
                    sg[0] = sg[1]; sg[1] = sg[2]; sg[2] = sg[3]; sg[3] = (y >> 31) | 1;
                }
            }
            // copy-out: the C lanes of a row move its tile with 128-bit loads/stores (adjacent lanes, adjacent 16 bytes)
            __syncwarp();
            if (valid) {
                uint4 *dst = my_dst + (size_t)(r * Cfg::RF + hh * Cfg::HF) * C * 2 / 16;
#pragma unroll
                for (int t = 0; t < kOutPerLane; t++)
                    if (t * C + (int)c < kOutGran) {
                        uint4 q;
                        q.x = ld_base[0][t * C * 4];
                        q.y = ld_base[1][t * C * 4];
                        q.z = ld_base[2][t * C * 4];
                        q.w = ld_base[3][t * C * 4];
                        dst[t * C] = q;
                    }
            }
        }
    }
}

// Shared-memory plan of decode_unrolled_kernel<C, B> for scale_factor_bits s: replicated table when it is <= 16 KB.
template <int C, int B>
static bool plan_unrolled(uint32_t s, bool *repl, uint32_t *lut_align, size_t *smem)
{
    using Cfg = UCfg<C, B>;
    const uint32_t entries = 1u << (s + B);
    *repl = entries * 128u <= 16384u;
    const uint32_t lut_bytes = *repl ? entries * 128u : entries * 4u;
    *lut_align = lut_bytes < 1024u ? 1024u : lut_bytes;  // power of two >= the table size
    *smem = (size_t)Cfg::kWarps * Cfg::kWarpBytes + *lut_align + lut_bytes;
    return Cfg::kWarps >= 8 && *smem <= 227u * 1024u;
}

template <int C, int B>
static cudaError_t launch_unrolled(const uint8_t *d_sea, int16_t *d_pcm, const DecStream *d_streams, const DecFastParams &p,
                                   const int32_t *tab, int *d_err, cudaStream_t stream)
{
    using Cfg = UCfg<C, B>;
    bool repl;
    uint32_t lut_align;
    size_t smem;
    if (!plan_unrolled<C, B>(p.s, &repl, &lut_align, &smem)) return cudaErrorInvalidConfiguration;
    const uint64_t chunks_per_cta = (uint64_t)Cfg::kWarps * Cfg::kRows;
    const uint64_t blocks = (p.total_chunks + chunks_per_cta - 1) / chunks_per_cta;
    cudaError_t e;
    if (repl) {
        e = cudaFuncSetAttribute(decode_unrolled_kernel<C, B, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        decode_unrolled_kernel<C, B, true><<<(unsigned)blocks, Cfg::kWarps * 32, smem, stream>>>(d_sea, d_pcm, d_streams, p, tab, lut_align, d_err);
    } else {
        e = cudaFuncSetAttribute(decode_unrolled_kernel<C, B, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        decode_unrolled_kernel<C, B, false><<<(unsigned)blocks, Cfg::kWarps * 32, smem, stream>>>(d_sea, d_pcm, d_streams, p, tab, lut_align, d_err);
    }
    return cudaGetLastError();
}

template <int C>
static bool plan_unrolled_b(uint32_t b, uint32_t s)
{
    bool repl;
    uint32_t la;
    size_t smem;
    switch (b) {
        case 1: return plan_unrolled<C, 1>(s, &repl, &la, &smem);
        case 2: return plan_unrolled<C, 2>(s, &repl, &la, &smem);
        case 3: return plan_unrolled<C, 3>(s, &repl, &la, &smem);
        case 4: return plan_unrolled<C, 4>(s, &repl, &la, &smem);
        case 5: return plan_unrolled<C, 5>(s, &repl, &la, &smem);
        case 6: return plan_unrolled<C, 6>(s, &repl, &la, &smem);
        case 7: return plan_unrolled<C, 7>(s, &repl, &la, &smem);
        case 8: return plan_unrolled<C, 8>(s, &repl, &la, &smem);
        default: return false;
    }
}

bool decode_unrolled_supported(const DecFastParams &p)
{
    if (p.channels != 1 && p.channels != 2) return false;
    if ((p.hdr_word & 0xffu) != 1u) return false;  // CBR only
    if (p.F != 20 || p.b < 1 || p.b > 8 || p.s < 1 || p.s > 8) return false;
    const uint32_t rf = ((p.channels * p.b) % 2 == 0) ? 80u : 160u;
    if (p.N % rf != 0 || p.N == 0) return false;
    return p.channels == 1 ? plan_unrolled_b<1>(p.b, p.s) : plan_unrolled_b<2>(p.b, p.s);
}

cudaError_t launch_decode_unrolled(const uint8_t *d_sea, int16_t *d_pcm, const DecStream *d_streams, const DecFastParams &p, DevTables tabs,
                                   int *d_err, cudaStream_t stream)
{
    if (p.total_chunks == 0) return cudaSuccess;
    const int32_t *tab = tabs.by_s[p.s];
#define SEA_UNROLLED(CC, BB) return launch_unrolled<CC, BB>(d_sea, d_pcm, d_streams, p, tab, d_err, stream)
    if (p.channels == 1) {
        switch (p.b) {
            case 1: SEA_UNROLLED(1, 1);
            case 2: SEA_UNROLLED(1, 2);
            case 3: SEA_UNROLLED(1, 3);
            case 4: SEA_UNROLLED(1, 4);
            case 5: SEA_UNROLLED(1, 5);
            case 6: SEA_UNROLLED(1, 6);
            case 7: SEA_UNROLLED(1, 7);
            default: SEA_UNROLLED(1, 8);
        }
    }
    switch (p.b) {
        case 1: SEA_UNROLLED(2, 1);
        case 2: SEA_UNROLLED(2, 2);
        case 3: SEA_UNROLLED(2, 3);
        case 4: SEA_UNROLLED(2, 4);
        case 5: SEA_UNROLLED(2, 5);
        case 6: SEA_UNROLLED(2, 6);
        case 7: SEA_UNROLLED(2, 7);
        default: SEA_UNROLLED(2, 8);
    }
#undef SEA_UNROLLED
}

}  // namespace sea
