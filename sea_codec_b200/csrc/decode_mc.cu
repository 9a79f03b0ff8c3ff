// decode_mc.cu -- decode_mc_kernel (decode_mc.cuh) for the even channel counts 4, 6 and 8, and the routing of all of them.
#include "decode_mc.cuh"

namespace sea {

// decode_mc_odd.cu
cudaError_t launch_decode_mc_odd(const uint8_t *d_sea, int16_t *d_pcm, const DecStream *d_streams, const DecFastParams &p, const int32_t *tab,
                                 int *d_err, cudaStream_t stream);

bool decode_mc_supported(const DecFastParams &p)
{
    if (p.channels < 3 || p.channels > 8) return false;
    if ((p.hdr_word & 0xffu) != 1u) return false;  // CBR chunks only
    if (p.F != 20 || p.s < 1 || p.s > 6 || p.b < 1 || p.b > 8) return false;
    // whole blocks and a whole cycle of store phases (MCfg): 20 frames for 4 and 8 channels, 40 for 6, 80 for the odd counts
    const uint32_t mult = p.channels == 4 ? mc_frame_multiple<4>() : p.channels == 6 ? mc_frame_multiple<6>() : p.channels == 8 ? mc_frame_multiple<8>() : 80u;
    return p.N != 0 && p.N % mult == 0;
}

template <int CT>
static cudaError_t launch_mc_s(const uint8_t *d_sea, int16_t *d_pcm, const DecStream *d_streams, const DecFastParams &p, const int32_t *tab,
                               int *d_err, cudaStream_t stream)
{
    if (p.s == 4u) return launch_mc_b<CT, true>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
    return launch_mc_b<CT, false>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
}

cudaError_t launch_decode_mc(const uint8_t *d_sea, int16_t *d_pcm, const DecStream *d_streams, const DecFastParams &p, DevTables tabs,
                             int *d_err, cudaStream_t stream)
{
    if (p.total_chunks == 0) return cudaSuccess;
    const int32_t *tab = tabs.by_s[p.s];
    switch (p.channels) {
        case 4: return launch_mc_s<4>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        case 6: return launch_mc_s<6>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        case 8: return launch_mc_s<8>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        default: return launch_decode_mc_odd(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
    }
}

}  // namespace sea
