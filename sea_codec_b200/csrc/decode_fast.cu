// decode_fast.cu -- decode_unrolled_kernel (decode_fast.cuh) for stereo streams, and the routing of both channel counts.
#include "decode_fast.cuh"

namespace sea {

// decode_fast_mono.cu
bool plan_unrolled_mono(uint32_t b, uint32_t s);
cudaError_t launch_decode_unrolled_mono(const uint8_t *d_sea, int16_t *d_pcm, const DecStream *d_streams, const DecFastParams &p, const int32_t *tab,
                                        int *d_err, cudaStream_t stream);

bool decode_unrolled_supported(const DecFastParams &p)
{
    if (p.channels != 1 && p.channels != 2) return false;
    if ((p.hdr_word & 0xffu) != 1u) return false;  // CBR only
    if (p.F != 20 || p.b < 1 || p.b > 8 || p.s < 1 || p.s > 8) return false;
    // whole halves (UCfg::HF = 40 stereo / 80 mono frames): that is also what keeps every chunk's PCM rows 32-byte aligned
    // given N % 20 == 0 -- e.g. 5000-frame stereo chunks (62.5 rounds) end with a round of one half
    const uint32_t hf = 80u / p.channels;
    if (p.N % hf != 0 || p.N == 0) return false;
    return p.channels == 1 ? plan_unrolled_mono(p.b, p.s) : plan_unrolled_b<2>(p.b, p.s);
}

cudaError_t launch_decode_unrolled(const uint8_t *d_sea, int16_t *d_pcm, const DecStream *d_streams, const DecFastParams &p, DevTables tabs,
                                   int *d_err, cudaStream_t stream)
{
    if (p.total_chunks == 0) return cudaSuccess;
    const int32_t *tab = tabs.by_s[p.s];
    if (p.channels == 1) return launch_decode_unrolled_mono(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
    return launch_unrolled_c<2>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
}

}  // namespace sea
