// int_ops.cu -- throughput of candidate integer instructions on sm_100a and which issue pipe they share.
// Each kernel runs CH independent dependency chains per thread of one op (and optionally an interleaved LOP3 / IMAD chain set);
// reports warp-instructions per cycle per SM sub-partition.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o int_ops int_ops.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int CH = 8, ITERS = 4096;

#define OP_I2IP(x, y)  asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define OP_DP2A(x, y)  asm volatile("dp2a.lo.s32.s32 %0, %1, %1, %0;" : "+r"(x) : "r"(y))
#define OP_LOP3(x, y)  asm volatile("lop3.b32 %0, %0, %1, 0x55aa55aa, 0x96;" : "+r"(x) : "r"(y))
#define OP_IMAD(x, y)  asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(x) : "r"(y))
#define OP_MNMX(x, y)  asm volatile("max.s32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define OP_PRMT(x, y)  asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(x) : "r"(y))
#define OP_SHF(x, y)   asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x) : "r"(y))
#define OP_SHR(x, y)   asm volatile("shr.s32 %0, %0, 1;" : "+r"(x))
#define OP_I2I(x, y)   asm volatile("cvt.sat.s16.s32 %0, %0;" : "+r"(x))
#define OP_IMADHI(x, y) asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define OP_VIADDMAX(x, y) x = __viaddmax_s32(x, y, -32768)
#define OP_VIMINRELU(x, y) x = __vimin_s32_relu(x, y)
#define OP_IADD3(x, y) asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define OP_NONE(x, y)

#define KERNEL(NAME, OPA, OPB)                                                        \
    __global__ void NAME(int *out, int seed, long long *cyc)                          \
    {                                                                                 \
        int a[CH], b[CH];                                                             \
        for (int i = 0; i < CH; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i + threadIdx.x; } \
        int y = seed | 1;                                                             \
        long long t0 = clock64();                                                     \
        for (int it = 0; it < ITERS; it++) {                                          \
            _Pragma("unroll") for (int i = 0; i < CH; i++) { OPA(a[i], y); OPB(b[i], y); } \
        }                                                                             \
        long long t1 = clock64();                                                     \
        int s = 0;                                                                    \
        for (int i = 0; i < CH; i++) s += a[i] ^ b[i];                                \
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;                               \
        if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;                      \
    }

KERNEL(k_lop3, OP_LOP3, OP_NONE)
KERNEL(k_imad, OP_IMAD, OP_NONE)
KERNEL(k_lop3_imad, OP_LOP3, OP_IMAD)
KERNEL(k_i2ip, OP_I2IP, OP_NONE)
KERNEL(k_i2ip_lop3, OP_I2IP, OP_LOP3)
KERNEL(k_i2ip_imad, OP_I2IP, OP_IMAD)
KERNEL(k_dp2a, OP_DP2A, OP_NONE)
KERNEL(k_dp2a_lop3, OP_DP2A, OP_LOP3)
KERNEL(k_dp2a_imad, OP_DP2A, OP_IMAD)
KERNEL(k_mnmx, OP_MNMX, OP_NONE)
KERNEL(k_mnmx_lop3, OP_MNMX, OP_LOP3)
KERNEL(k_prmt, OP_PRMT, OP_NONE)
KERNEL(k_shf, OP_SHF, OP_NONE)
KERNEL(k_shr, OP_SHR, OP_NONE)
KERNEL(k_shr_lop3, OP_SHR, OP_LOP3)
KERNEL(k_i2i, OP_I2I, OP_NONE)
KERNEL(k_i2i_lop3, OP_I2I, OP_LOP3)
KERNEL(k_imadhi, OP_IMADHI, OP_NONE)
KERNEL(k_imadhi_lop3, OP_IMADHI, OP_LOP3)
KERNEL(k_viaddmax, OP_VIADDMAX, OP_NONE)
KERNEL(k_viaddmax_lop3, OP_VIADDMAX, OP_LOP3)
KERNEL(k_viminrelu, OP_VIMINRELU, OP_NONE)
KERNEL(k_iadd, OP_IADD3, OP_NONE)
KERNEL(k_iadd_lop3, OP_IADD3, OP_LOP3)
KERNEL(k_iadd_imad, OP_IADD3, OP_IMAD)

template <typename K>
void run(const char *name, K k, int nops, int *d_out, long long *d_cyc)
{
    const int warps = 16;  // 4 warps per sub-partition
    k<<<1, warps * 32>>>(d_out, 12345, d_cyc);
    k<<<1, warps * 32>>>(d_out, 12345, d_cyc);
    cudaDeviceSynchronize();
    long long cyc;
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    const double winstr = (double)warps / 4 * CH * ITERS * nops;  // warp-instructions per sub-partition
    printf("%-18s %9lld cycles  %.3f warp-instr/clk/SMSP  (%.2f clk per op-pair)\n", name, cyc, winstr / cyc, cyc / ((double)warps / 4 * CH * ITERS));
}

int main()
{
    int *d_out;
    long long *d_cyc;
    cudaMalloc(&d_out, 1 << 20);
    cudaMalloc(&d_cyc, 8);
#define RUN(k, n) run(#k, k, n, d_out, d_cyc)
    RUN(k_lop3, 1); RUN(k_imad, 1); RUN(k_lop3_imad, 2);
    RUN(k_i2ip, 1); RUN(k_i2ip_lop3, 2); RUN(k_i2ip_imad, 2);
    RUN(k_dp2a, 1); RUN(k_dp2a_lop3, 2); RUN(k_dp2a_imad, 2);
    RUN(k_mnmx, 1); RUN(k_mnmx_lop3, 2); RUN(k_prmt, 1); RUN(k_shf, 1); RUN(k_shr, 1); RUN(k_shr_lop3, 2);
    RUN(k_i2i, 1); RUN(k_i2i_lop3, 2); RUN(k_imadhi, 1); RUN(k_imadhi_lop3, 2);
    RUN(k_viaddmax, 1); RUN(k_viaddmax_lop3, 2); RUN(k_viminrelu, 1);
    RUN(k_iadd, 1); RUN(k_iadd_lop3, 2); RUN(k_iadd_imad, 2);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
