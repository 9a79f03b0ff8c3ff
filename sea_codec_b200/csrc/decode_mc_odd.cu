// decode_mc_odd.cu -- decode_mc_kernel (decode_mc.cuh) for 3, 5 and 7 channels: four store phases per cycle of bodies, one PCM
// word straddling every pair of frames.  3 channels is what the reference's own tests decode (tests/test.rs:10,38).
#include "decode_mc.cuh"

namespace sea {

static_assert(mc_frame_multiple<3>() == 80u && mc_frame_multiple<5>() == 80u && mc_frame_multiple<7>() == 80u, "decode_mc_supported assumes 80");

cudaError_t launch_decode_mc_odd(const uint8_t *d_sea, int16_t *d_pcm, const DecStream *d_streams, const DecFastParams &p, const int32_t *tab,
                                 int *d_err, cudaStream_t stream)
{
    switch (p.channels) {
        case 3: return launch_mc_b<3, false>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        case 5: return launch_mc_b<5, false>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        case 7: return launch_mc_b<7, false>(d_sea, d_pcm, d_streams, p, tab, d_err, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}

}  // namespace sea
