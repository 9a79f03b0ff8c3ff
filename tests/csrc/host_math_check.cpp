// Host-side instantiation of the kernels' shared arithmetic (sea_codec_b200/csrc/sea_common.cuh) so that the CPU test
// suite can check the closed forms against the oracle's tables without a GPU.  Test-only; not part of libsea_b200.so.
#include "../../sea_codec_b200/csrc/sea_common.cuh"

extern "C" {
unsigned hm_quant_code(int r, int recip, unsigned b) { return sea::quant_code(r, recip, b); }
int hm_predict(const int *w, const int *h) { return sea::lms_predict(w, h); }
void hm_update(int *w, int *h, int y, int d) { sea::lms_update(w, h, y, d); }
unsigned long long hm_penalty(const int *w) { return sea::lms_penalty(w); }
int hm_clamp(int v) { return sea::clamp_i16(v); }
}
extern "C" unsigned long long hm_rank_step(unsigned long long rank, int err, const int *w)
{
    // the narrow form must agree with the wide one wherever its precondition holds
    unsigned long long wide = sea::rank_step<false>(rank, err, w);
    if (sea::weights_stay_narrow(w, 0) && sea::rank_step<true>(rank, err, w) != wide) return ~wide;
    return wide;
}
