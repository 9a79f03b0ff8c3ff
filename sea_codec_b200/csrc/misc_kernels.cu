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

// ---- synthetic PCM (SURVEY.md 8d "Synthetic input"; the integer recipe of sea_codec_b200/synth.py, sample for sample) ----------
//   x[n] = clamp((A * sine[phase >> 20] >> 15) + noise),  phase = (c * 0x1F3D5B79 + n * step_k) mod 2^32,
//   noise = (splitmix64((seed + k*256 + c) * 0x2545F4914F6CDD1D + n) >> 40) mod (2*amp + 1) - amp
// One thread per frame of a stream (all channels), grid.y = stream: 16-bit stores coalesce along the interleaved frame order.
__global__ void __launch_bounds__(256) synth_kernel(int16_t *__restrict__ pcm, uint64_t stream_stride, uint32_t n_frames, uint32_t channels,
                                                    const uint32_t *__restrict__ ids, const uint32_t *__restrict__ steps,
                                                    const int32_t *__restrict__ sine, uint64_t seed, int32_t amp, int32_t noise_amp)
{
    const uint32_t k = ids[blockIdx.y], step = steps[blockIdx.y];
    int16_t *dst = pcm + (uint64_t)blockIdx.y * stream_stride;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < n_frames; n += gridDim.x * blockDim.x) {
        for (uint32_t c = 0; c < channels; c++) {
            const uint32_t phase = c * 0x1F3D5B79u + n * step;
            const int32_t tone = (int32_t)(((int64_t)amp * (int64_t)__ldg(sine + (phase >> 20))) >> 15);
            uint64_t z = (seed + (uint64_t)k * 256u + c) * 0x2545F4914F6CDD1Dull + n + 0x9E3779B97F4A7C15ull;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            const int32_t noise = (int32_t)((uint32_t)(z >> 40) % (uint32_t)(2 * noise_amp + 1)) - noise_amp;  // z >> 40 < 2^24
            const int32_t v = tone + noise;
            dst[(uint64_t)n * channels + c] = (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v));
        }
    }
}

cudaError_t launch_synth(int16_t *d_pcm, uint64_t stream_stride, uint32_t n_streams, uint32_t n_frames, uint32_t channels, const uint32_t *d_ids,
                         const uint32_t *d_steps, const int32_t *d_sine, uint64_t seed, int32_t amp, int32_t noise_amp, cudaStream_t stream)
{
    if (n_streams == 0 || n_frames == 0) return cudaSuccess;
    uint32_t bx = (n_frames + 255u) / 256u;
    if (bx > 1024u) bx = 1024u;
    for (uint32_t s0 = 0; s0 < n_streams; s0 += 65535u) {  // grid.y limit
        const uint32_t ns = n_streams - s0 < 65535u ? n_streams - s0 : 65535u;
        synth_kernel<<<dim3(bx, ns), 256, 0, stream>>>(d_pcm + (uint64_t)s0 * stream_stride, stream_stride, n_frames, channels, d_ids + s0,
                                                       d_steps + s0, d_sine, seed, amp, noise_amp);
    }
    return cudaGetLastError();
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
