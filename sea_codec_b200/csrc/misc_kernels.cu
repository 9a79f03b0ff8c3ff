// misc_kernels.cu -- measurement micro-kernel: INT32 issue-rate ceiling used as the encoder roofline denominator
// ("a measured INT32 peak from a micro-kernel of independent IADD3/LOP3/IMAD chains").
#include "sea_kernels.h"

namespace sea {

constexpr int kPeakIters = 4096;
constexpr int kPeakChains = 8;

// MODE 0: mad.lo (fma pipe)   MODE 1: lop3 + add (alu pipe)   MODE 2: both interleaved
template <int MODE>
__global__ void __launch_bounds__(512) int32_peak_kernel(uint32_t *sink)
{
    uint32_t a[kPeakChains], b[kPeakChains];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < kPeakChains; i++) {
        a[i] = t * 2654435761u + i;
        b[i] = t ^ (0x9e3779b9u * (i + 1));
    }
    const uint32_t m = t | 1u, k = t + 12345u;
    for (int it = 0; it < kPeakIters; it++) {
#pragma unroll
        for (int i = 0; i < kPeakChains; i++) {
            if (MODE == 0 || MODE == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(k));
            if (MODE == 1 || MODE == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(m), "r"(k));
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < kPeakChains; i++) acc ^= a[i] ^ b[i];
    if (acc == 0x12345678u) sink[0] = acc;  // keep the chains alive
}

cudaError_t launch_int32_peak(int mode, uint32_t *d_sink, uint64_t *lane_ops, cudaStream_t stream)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t blocks = (uint32_t)sms * 4u, threads = 512;
    const uint64_t per_thread = (uint64_t)kPeakIters * kPeakChains * (mode == 2 ? 2 : 1);
    *lane_ops = per_thread * blocks * threads;
    if (mode == 0) int32_peak_kernel<0><<<blocks, threads, 0, stream>>>(d_sink);
    else if (mode == 1) int32_peak_kernel<1><<<blocks, threads, 0, stream>>>(d_sink);
    else int32_peak_kernel<2><<<blocks, threads, 0, stream>>>(d_sink);
    return cudaGetLastError();
}

}  // namespace sea
