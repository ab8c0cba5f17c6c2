// nns_common.cuh -- shared device helpers: packed f32x2 arithmetic, mbarrier / bulk-copy (TMA)
// PTX wrappers, packed (dist, idx) keys.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nns {

typedef unsigned long long u64;

constexpr int LB = 128;                                // reference block (NNS_B200_REF_BLOCK)
constexpr u64 KEY_INIT = 0x7F80000000000000ull;        // (+INF, index 0)

__host__ __device__ __forceinline__ u64 pack_key(float dist, int idx)
{
#ifdef __CUDA_ARCH__
    return ((u64)__float_as_uint(dist) << 32) | (u64)(unsigned)idx;
#else
    union { float f; unsigned u; } c; c.f = dist;
    return ((u64)c.u << 32) | (u64)(unsigned)idx;
#endif
}

// ---- packed FP32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2): one issue slot, two lanes ----
__device__ __forceinline__ u64 pk2(float lo, float hi)
{
    u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void upk2u(u64 v, unsigned& lo, unsigned& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b)
{
    u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
// 3-input minimum (FMNMX3).  NaN operands are ignored unless all are NaN (IEEE minNum).
__device__ __forceinline__ float min3(float a, float b, float c)
{
    float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}

// ---- shared-memory / mbarrier / bulk async copy ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
// Blocks until the phase with the given parity completes.  try_wait suspends the thread in
// hardware for up to the time hint (ns) instead of polling, so a waiting producer lane does not
// steal issue slots from the compute warps that share its scheduler.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(0x989680u)
            : "memory");
    } while (!done);
}
// 1-D bulk copy global -> shared through the TMA unit (SASS: UBLKCP), completion on an mbarrier.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ ulonglong2 lds_v2u64(const float* p)
{
    return *reinterpret_cast<const ulonglong2*>(p);
}

__device__ __forceinline__ float nan_f() { return __int_as_float(0x7fc00000); }
__device__ __forceinline__ float inf_f() { return __int_as_float(0x7f800000); }

}  // namespace nns
