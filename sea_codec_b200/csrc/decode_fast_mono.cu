// decode_fast_mono.cu -- decode_unrolled_kernel (decode_fast.cuh) for mono streams.
#include "decode_fast.cuh"

namespace sea {

bool plan_unrolled_mono(uint32_t b, uint32_t s) { return plan_unrolled_b<1>(b, s); }

cudaError_t launch_decode_unrolled_mono(const uint8_t *d_sea, int16_t *d_pcm, const DecStream *d_streams, const DecFastParams &p, const int32_t *tab,
                                        int *d_err, cudaStream_t stream)
{
    return launch_unrolled_c<1>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
}

}  // namespace sea
