// sea_device.cuh -- device-side helpers shared by the decode kernels (decode_fast / decode_vbr / decode_mc / decode_latency /
// decode_kernels): error word, stream lookup, shared-window loads, cp.async staging, 256-bit stores.
#pragma once
#include "sea_kernels.h"

namespace sea {
namespace dev {

// first failure wins (device error codes: sea_kernels.h kDev*)
__device__ __forceinline__ void report(int *err, int code) { atomicCAS(err, 0, code); }

// last stream whose chain_begin <= chain (chains (chunk, channel) of all streams are numbered consecutively)
__device__ __forceinline__ uint32_t find_stream(const DecStream *streams, uint32_t n_streams, uint64_t chain)
{
    uint32_t lo = 0, hi = n_streams;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((uint64_t)streams[mid].chain_begin <= chain) lo = mid;
        else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// loads on a 32-bit shared-window address (typed ld.shared: a generic pointer would make them LD.E)
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int32_t lds_s32(uint32_t addr)
{
    int32_t v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int32_t lds_s16(uint32_t addr)
{
    int32_t v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// 16-byte global -> shared copies that bypass L1 (each lane stages its own chunk row), optionally predicated.  SEA_CP_CA (tuning
// builds): allocate in L1 instead, so that the second granule of a 32-byte sector hits the line the first one brought in.
#ifdef SEA_CP_CA
#define SEA_CP_ASYNC16 "cp.async.ca.shared.global"
#else
#define SEA_CP_ASYNC16 "cp.async.cg.shared.global"
#endif
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile(SEA_CP_ASYNC16 " [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_if(bool pred, uint32_t dst, const void *src)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %0, 0;\n\t@p " SEA_CP_ASYNC16 " [%1], [%2], 16;\n\t}" ::"r"((int)pred), "r"(dst), "l"(src)
        : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one full 32-byte sector per lane
__device__ __forceinline__ void st_global_256(void *p, const uint32_t (&v)[8])
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
                 "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

}  // namespace dev
}  // namespace sea
