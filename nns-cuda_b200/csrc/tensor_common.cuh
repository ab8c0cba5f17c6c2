// tensor_common.cuh -- pieces shared by the tcgen05 kernels (tensor_search.cu: contractions of up to 144
// columns, compile-time geometry; tensor_longk.cu: K-loop for 128 < k <= 509): CTA shape, tcgen05 / TMEM /
// mbarrier wrappers, UMMA descriptors, operand-image geometry, the error bound and the candidate buffer.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "nns_internal.h"

namespace nns {

constexpr int T_BM = 256;     // query rows per CTA (two M = 128 accumulator halves)
constexpr int T_BN = 128;     // references per tile == one index block
constexpr int T_MAX_STAGES = 24;  // upper bound of the B ring depth (the mbarrier area is sized for it)
// experiment knobs (tools/tensor_tune.sh builds the variants; the defaults are the measured best)
#ifndef NNS_T_TEAMS
#define NNS_T_TEAMS 2       // epilogue teams of 8 warps; team i owns TMEM buffer i and reduces the tiles t % 2 == i
#endif
#ifndef NNS_T_SUB
#define NNS_T_SUB 2         // accumulator units per reference tile for the short contractions (KB = 0): 1 or 2
#endif
#ifndef NNS_T_LD64
#define NNS_T_LD64 (-1)     // TMEM columns per epilogue load: 0 = 32, 1 = 64, -1 = 64 for 64-column units (SUB = 2) else 32
#endif
#ifndef NNS_T_SPIN
#define NNS_T_SPIN 2        // bit 0: the epilogue warps poll their mbarrier, bit 1: the MMA issuer polls
#endif
#ifndef NNS_T_SPIN_F16
#define NNS_T_SPIN_F16 0    // the same for the F16 screens: three polling issuers cost more (power-capped clock) than their wake-up saves
#endif
#ifndef NNS_T_EPI_NORM
#define NNS_T_EPI_NORM 0    // F16 mode, k = 62..64 / 126..128: 1 = the epilogue adds |r'|^2 (no norm columns, one MMA step less).
                            // Measured on C4 (k = 128, power-capped at 1 kW): 232 ms at 1.49 GHz vs 229 ms at 1.55 GHz with the norm
                            // columns -- the MMA step saved is paid back in clock, so it stays off; parity-tested in both settings
#endif
#ifndef NNS_T_PIPE
#define NNS_T_PIPE 2        // epilogue of the 64-reference units: 0 = one unit per loop trip, a candidate test per 32-column chunk;
#endif                      // 2 = NNS_T_TRIP units per trip and ONE test per trip (12 % faster on C2: profiles/r2_tune_trip.txt);
                            // 1 = additionally two rotating 32-column register sets with a load in flight under every reduction (slower)
#ifndef NNS_T_ISS
#define NNS_T_ISS 2         // MMA-issuing threads of the short-contraction screens (k <= 9, plain mid-k): 1 or 2
#endif
#ifndef NNS_T_ISS_F16
#define NNS_T_ISS_F16 3     // the same for the F16-accumulator variants: 3 = A in TMEM, three buffers, one issuer each; 2 / 4 = SS form, four buffers
#endif
#ifndef NNS_T_TEAMS_F16
#define NNS_T_TEAMS_F16 3   // epilogue teams of the F16 short-contraction screen (NNS_T_ISS_F16 = 3): 2 or 3
#endif
#ifndef NNS_T_TRIP
#define NNS_T_TRIP 2        // NNS_T_PIPE = 2: units per loop trip / candidate test
#endif
#ifndef NNS_T_TS
#define NNS_T_TS 1          // A operand in tensor memory: 0 = never, 1 = contractions of 64 / 80 columns, 2 = also 128 / 144
#endif
#ifndef NNS_T_EXPERIMENT
#define NNS_T_EXPERIMENT 0  // timing experiments only (wrong results): 1 = epilogue reduces 2 of 32 columns,
#endif                      // 2 = epilogue does not read TMEM at all, 3 = additionally no MMA is issued
#ifdef NNS_T_TRACE  // = first tile of the 32-tile window: clock64() timeline of CTA (0,0), tools/tensor_trace.py
__device__ long long g_ttrace[36][32];
#define T_TRACE(ev, t) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (t) >= NNS_T_TRACE && (t) < NNS_T_TRACE + 32) g_ttrace[ev][(t) - NNS_T_TRACE] = clock64(); } while (0)
extern "C" int nns_b200_debug_trace(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_ttrace, sizeof(g_ttrace)); }
#else
#define T_TRACE(ev, t) ((void)0)
#endif
constexpr int T_TEAMS = NNS_T_TEAMS;
// Every accumulator mbarrier must have ONE waiting team, in phase order: a waiter that is more than one
// phase ahead of an mbarrier misreads its parity.  Two teams and 2 or 4 buffers satisfy that (buffer
// parity = team); three teams over 2 or 4 buffers do not (deadlocked on B200).
static_assert(T_TEAMS == 1 || T_TEAMS == 2, "epilogue teams: 1 or 2");
constexpr int T_TEAM_WARPS = 8;                        // one warp per (TMEM lane quarter, accumulator half)
constexpr int T_SERVICE_WARPS = 2;                     // warp 0 = TMA producer, warp 1 = MMA issuer
constexpr int T_THREADS = 32 * (T_SERVICE_WARPS + T_TEAMS * T_TEAM_WARPS);
// one CTA per SM: the whole register file.  Registers are per scheduler (16384 each), and the
// busiest one hosts ceil(warps / 4) warps
constexpr int T_MAX_REGS = (16384 / (32 * ((T_THREADS / 32 + 3) / 4))) & ~7;

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc512(uint32_t smem_dst)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_dst) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc512(uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, BF16 x BF16 -> FP32, M = 128, N = 128, K = 16
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, u64 adesc, u64 bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same with the A operand in tensor memory (lane = row, 8 columns = 16 bf16 of one K step): only B is
// fetched from shared memory, half the operand traffic of the shared-memory form
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, u64 bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// registers -> tensor memory: each thread writes 8 consecutive columns of its own lane
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4 lo, const uint4 hi)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr)
        : "memory");
}
// 64 columns of 16-bit accumulators (F16 accumulators occupy the low half of their 32-bit cell), two per register:
// register i = columns 2i (low half) and 2i + 1 (high half)
__device__ __forceinline__ void tmem_ld64_pack16(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ uint32_t hmin2(uint32_t a, uint32_t b)  // HMNMX2: per-half minimum, NaN loses
{
    uint32_t d;
    asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c)  // HFMA2: a * b + c per half, one rounding
{
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ float hmin2_to_float(uint32_t a)  // the smaller half, exactly, as FP32
{
    float lo, hi;
    asm("{.reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(lo), "=f"(hi) : "r"(a));
    return fminf(lo, hi);
}
// true in exactly one (the lowest active) lane of the warp; must be called with all 32 lanes converged
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the same, tied to the 32 destination registers of a load that may still be in flight: the compiler sees the
// wait as the definition of v[], so no use of v[] can be scheduled above it (the pipelined epilogue keeps a
// load in flight across a whole reduction of the other register set)
__device__ __forceinline__ void tmem_ld_wait_for(uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.wait::ld.sync.aligned;"
        : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
          "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
          "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
          "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
        :
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_for64(uint32_t (&v)[64])
{
    asm volatile(
        "tcgen05.wait::ld.sync.aligned;"
        : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
          "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]),
          "+r"(v[32]), "+r"(v[33]), "+r"(v[34]), "+r"(v[35]), "+r"(v[36]), "+r"(v[37]), "+r"(v[38]), "+r"(v[39]), "+r"(v[40]), "+r"(v[41]), "+r"(v[42]), "+r"(v[43]), "+r"(v[44]), "+r"(v[45]), "+r"(v[46]), "+r"(v[47]),
          "+r"(v[48]), "+r"(v[49]), "+r"(v[50]), "+r"(v[51]), "+r"(v[52]), "+r"(v[53]), "+r"(v[54]), "+r"(v[55]), "+r"(v[56]), "+r"(v[57]), "+r"(v[58]), "+r"(v[59]), "+r"(v[60]), "+r"(v[61]), "+r"(v[62]), "+r"(v[63])
        :
        : "memory");
}

// UMMA shared-memory descriptor: K-major operand, 128-byte swizzle, rows of 128 B (64 bf16),
// 8-row swizzle atoms 1024 B apart (stride byte offset), version 1 (Blackwell).
__device__ __forceinline__ u64 umma_desc_sw128(uint32_t smem_addr)
{
    u64 d = (u64)((smem_addr & 0x3FFFFu) >> 4);  // start address, 16-byte units, bits [0,14)
    d |= (u64)1 << 16;                            // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (u64)(1024 >> 4) << 32;                  // stride byte offset, bits [32,46)
    d |= (u64)1 << 46;                            // descriptor version, bits [46,48)
    d |= (u64)2 << 61;                            // layout type SWIZZLE_128B, bits [61,64)
    return d;
}
// UMMA descriptor for the extra K = 16 step: K-major, no swizzle ("interleave"): 8-row x 16-byte
// core matrices; the two 16-byte K chunks of a row are `rows * 16` bytes apart (leading byte
// offset), consecutive 8-row groups 128 bytes apart (stride byte offset) -> layout [chunk][row][16 B]
__device__ __forceinline__ u64 umma_desc_interleave(uint32_t smem_addr, uint32_t rows)
{
    u64 d = (u64)((smem_addr & 0x3FFFFu) >> 4);
    d |= (u64)((rows * 16u) >> 4) << 16;  // leading byte offset
    d |= (u64)(128 >> 4) << 32;           // stride byte offset
    d |= (u64)1 << 46;                    // descriptor version
    return d;                             // layout type 0 = no swizzle
}

// order-preserving float <-> uint (for atomicMin on possibly negative scores)
__device__ __forceinline__ unsigned f2ord(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// mbarrier waits on the MMA <-> epilogue critical path.  NNS_T_SPIN bit 0: the epilogue warps poll
// (try_wait without a suspend hint) instead of sleeping on the barrier; bit 1: the MMA issuer polls.
__device__ __forceinline__ void mbar_wait_poll(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
template <int SPIN = NNS_T_SPIN>
__device__ __forceinline__ void mbar_wait_hot(uint32_t bar, uint32_t parity)  // epilogue
{
    if (SPIN & 1) mbar_wait_poll(bar, parity); else mbar_wait(bar, parity);
}
template <int SPIN = NNS_T_SPIN>
__device__ __forceinline__ void mbar_wait_mma(uint32_t bar, uint32_t parity)  // MMA issuer
{
    if (SPIN & 2) mbar_wait_poll(bar, parity); else mbar_wait(bar, parity);
}

// byte offset of element (row, t) inside an operand image with `rows` rows:
// [t/64][rows][128 B], 16-byte chunks XOR-swizzled with the row (Swizzle<3,4,3>)
__device__ __forceinline__ size_t image_chunk_offset(int rows, int row, int kb, int chunk)
{
    return (size_t)kb * rows * 128 + (size_t)row * 128 + (size_t)((chunk ^ (row & 7)) * 16);
}

// Operand image geometry.  The contraction dimension is laid out as KB blocks of 64 columns in the
// K-major 128-byte-swizzle layout ([KB][rows][128 B]) followed by KS steps of 16 columns in the
// K-major no-swizzle ("interleave") layout ([2*KS][rows][16 B]); `norm_col` is the first of the
// three columns that carry |r'|^2 (references) / 1 (queries).  Small k needs no swizzled block at
// all: for k <= 4 the 3k split-precision columns and the norm fit ONE K = 16 MMA step.
struct TensorGeom {
    int KB, KS, norm_col, ndata, split;
    int en;   // F16 mode, k = 62..64 / 126..128: |r'|^2 is added by the epilogue instead of a sixteen-column MMA step of its own
};
// plain = single BF16 column per dimension even where the split-precision layout exists (see "Precision mode")
__host__ __device__ inline TensorGeom tensor_geom(int k, bool plain = false)
{
    TensorGeom g;
    g.en = 0;
    g.split = (k <= TENSOR_SPLIT_MAX_K && !plain) ? 1 : 0;
    g.ndata = g.split ? 3 * k : k;
    // F16 mode (`plain` = the F16 layout): where the data columns fill the 64-column blocks exactly, the three norm
    // columns would cost a whole extra K = 16 step (k = 128: 9 instead of 8).  The 16-bit epilogue has the slack to add
    // t * s^2 |r'|^2 itself (one HFMA2 per two columns, on the otherwise idle FMA pipe): no norm columns at all.
    if (plain && NNS_T_EPI_NORM && ((g.ndata > 61 && g.ndata <= 64) || (g.ndata > 125 && g.ndata <= 128))) {
        g.KB = g.ndata <= 64 ? 1 : 2; g.KS = 0; g.norm_col = -100; g.en = 1;
        return g;
    }
    // the three norm columns ride in the last block whenever it has room for them
    if (g.ndata + 3 <= 16) { g.KB = 0; g.KS = 1; g.norm_col = g.ndata; }
    else if (g.ndata + 3 <= 32) { g.KB = 0; g.KS = 2; g.norm_col = g.ndata; }
    else if (g.ndata + 3 <= 64) { g.KB = 1; g.KS = 0; g.norm_col = g.ndata; }
    else if (g.ndata <= 64) { g.KB = 1; g.KS = 1; g.norm_col = 64; }
    else if (g.ndata + 3 <= 128) { g.KB = 2; g.KS = 0; g.norm_col = g.ndata; }
    else if (g.ndata <= 128) { g.KB = 2; g.KS = 1; g.norm_col = 128; }
    else { g.KB = (g.ndata + 3 + 63) / 64; g.KS = 0; g.norm_col = g.ndata; }  // K-loop kernel (tensor_longk.cu): whole 64-column blocks
    return g;
}
__host__ __device__ constexpr size_t image_bytes(int rows, int KB, int KS) { return (size_t)rows * (KB * 128 + KS * 32); }
// byte offset of the 16-byte chunk holding columns [8*chunk, 8*chunk + 8) of `row`
__device__ __forceinline__ size_t image_chunk_at(int rows, int KB, int row, int chunk)
{
    if (chunk < KB * 8) return image_chunk_offset(rows, row, chunk >> 3, chunk & 7);
    return (size_t)KB * rows * 128 + (size_t)(chunk - KB * 8) * rows * 16 + (size_t)row * 16;
}

__device__ __forceinline__ bool tensor_mode_mismatch(const unsigned* __restrict__ mode_word, const unsigned my_mode)
{
    return mode_word != nullptr && *reinterpret_cast<const volatile unsigned*>(mode_word) != my_mode;
}

// E(q) >= |S~ - S| + |d_V0 - D'| for every reference (see the file header), a = |q'|, rmax = max |r'|.
// BF16 keeps 8 significant bits: unit roundoff u = 2^-8 (round to nearest).
//   operand rounding, plain: each product (-2 q'_i)(r'_i) is off by <= (2u + u^2) of itself, the sum by
//   <= (2u + u^2) * 2 a rmax = 2^-6 (1 + 2^-9) a rmax;  split precision: x = xh + xl + ex with |xl| <= u (1 + u) |x|,
//   |ex| <= u^2 |x|; the MMA accumulates xh.yh + xh.yl + xl.yh, i.e. drops xl.yl + ex.y + (x - ex).ey
//   <= 3 u^2 (1 + 2u) |x_i| |y_i|, summed <= 3.03 * 2^-16 * 2 a rmax.
//   The MMA's FP32 accumulation is charged 2^-21 per term (truncating adders); FP32 |r'|^2 and its 3-term
//   split (KP + 5) 2^-24 rmax^2; centring and V0's own rounding (KP + 8) 2^-24 (a + rmax)^2; 5 % on top.
__host__ __device__ inline float tensor_error_bound(bool split, int KP, float a, float rmax)
{
    const float u24 = 5.9604645e-8f;
    const float c_round = split ? 6.06f * 1.5258789e-5f : 0.015625f * 1.002f;
    const float E = (c_round + (float)KP * 2.04f * 4.7683716e-7f) * a * rmax + (KP + 5) * u24 * rmax * rmax +
                    (KP + 8) * u24 * (a + rmax) * (a + rmax);
    return E * 1.05f;
}

// ---- F16 mode (THDR_MODE = TMODE_F16): F16 operands AND F16 accumulators ----
// The epilogue of the short contractions is bound by the ALU pipe: one FMNMX3 retires two new FP32 values at
// 2 warp-instructions/clk/SM.  HMNMX2 retires two new F16 values at 4 warp-instructions/clk/SM
// (tools/ubench_f16acc.cu, profiles/r2d_ubench_f16acc.txt), and tcgen05.ld ... pack::16b delivers an F16
// accumulator tile two columns per register -- twice the reduction rate and half the registers.  F16 has 11
// significant bits (u = 2^-11) but only 5 exponent bits, so the operands are scaled by powers of two:
//   references  s r'        |r'|^2 column: s^2 |r'|^2 (three F16 terms)        s: per index (THDR_SCALE),
//   queries    -2 t s q'    norm columns:  t                                   t: per query
// and the accumulator holds u_q S~ with u_q = t s^2 (kept per query; every comparison happens in unscaled FP32
// units).  s brings the SAMPLED radius of the reference cloud to [4, 8) (the exact maximum is only known once
// the images exist); an index whose true radius R = s rmax exceeds F16_R_MAX is flagged unusable (band = INF ->
// FP32 fallback), i.e. the sample may under-estimate the radius 22x.  t = the largest power of two <= 2^8 with
// t R (R + 2A) <= 2^15 and 2 t A <= 2^15 (A = s a): no operand and no partial sum can overflow; t < 2^-14
// (a query 2^20 cloud radii away) makes that query unusable.
// Error: operand rounding (2u + u^2) * 2 a rmax = 2^-9 (1 + 2^-12) a rmax; every MMA instruction rounds the running
// sum to F16 (measured: round to nearest; charged u (1 + 2^-6) of the largest partial sum rmax^2 + 2 a rmax);
// subnormal operands / sums: absolute 2^-25 in scaled units.  The FP32 terms are those of the BF16 modes.
constexpr float F16_R_MAX = 176.0f;
__host__ __device__ inline float tensor_f16_ref_scale(float rmax_sampled)  // power of two, sampled radius -> [4, 8)
{
    if (!(rmax_sampled > 1e-30f) || !(rmax_sampled < 1e30f)) return 1.0f;
    int e;
    frexpf(rmax_sampled, &e);  // rmax = f * 2^e, f in [0.5, 1)
    return ldexpf(1.0f, 3 - e);
}
__host__ __device__ inline float tensor_f16_query_scale(float A, float R)  // 0 = unusable
{
    if (!(A <= 1e30f) || !(R <= F16_R_MAX)) return 0.0f;
    float lim = 256.0f;
    const float p = R * (R + 2.0f * A);
    if (p > 0.0f) lim = fminf(lim, 32768.0f / p);
    if (A > 0.0f) lim = fminf(lim, 16384.0f / A);
    if (!(lim >= 6.1035156e-5f)) return 0.0f;
    int e;
    const float f = frexpf(lim, &e);  // lim = f * 2^e
    (void)f;
    return ldexpf(1.0f, e - 1);       // largest power of two <= lim
}
// en: the norm is a single F16 number (u11 rmax^2) added by one HFMA2 in the epilogue (one more rounding of the full sum)
__host__ __device__ inline float tensor_error_bound_f16(int KP, int k, float a, float rmax, float s, float t, bool en = false)
{
    const float u24 = 5.9604645e-8f, u11 = 4.8828125e-4f;
    const int steps = KP / 16 + (en ? 1 : 0);
    const float big = rmax * rmax + 2.0f * a * rmax;
    const float sub = (sqrtf((float)k) * (2.0f * t * s * a + s * rmax) + (float)(steps + 3)) * 2.9802322e-8f / (t * s * s);
    const float E = (2.0f * u11 * 1.001f + (float)KP * 2.04f * 4.7683716e-7f) * a * rmax + (float)steps * u11 * 1.016f * 1.002f * big + sub +
                    (en ? u11 * 1.001f * rmax * rmax : 0.0f) + (KP + 5) * u24 * rmax * rmax + (KP + 8) * u24 * (a + rmax) * (a + rmax);
    return E * 1.05f;
}

// ---------------------------------------------------------------------------------------------
// candidates
// ---------------------------------------------------------------------------------------------
struct TensorCand { int q; int unit; float smin; };  // unit = 32 consecutive references (tile * 4 + chunk)

// Candidate buffer.  Every CTA of the screen owns a private region of `region_cap` records and
// allocates slots from a SHARED-memory counter, so that emitting a candidate never waits for a
// global atomic round trip (~600 clk, measured with tools/tensor_trace.py: the epilogue warp that
// waited stalled the MMA issuer through the accumulator hand-off).  A CTA whose region is full
// spills into a common region through a global counter; only if that overflows too does the FP32
// wide kernel redo the search.
struct CandBuf {
    TensorCand* rec;       // [n_ctas * region_cap] CTA regions, then [common_cap] common records
    unsigned* cta_count;   // [n_ctas] records used in each CTA region
    unsigned* common_count;  // common records requested (per query batch)
    unsigned* status;      // per search: [0] total records emitted, [1] overflow flag, [2] record capacity
    unsigned region_cap;   // multiple of 32
    unsigned common_cap;
    unsigned n_ctas;
    unsigned fixed_threshold;  // K-nearest search: approx_min[q] is a FIXED threshold, never lowered by the screen
};

__device__ __forceinline__ void cand_emit(const CandBuf& cb, unsigned* s_count, const unsigned cta, const TensorCand& c)
{
    const unsigned slot = atomicAdd(s_count, 1u);  // shared memory; warp-aggregated by the compiler
    if (slot < cb.region_cap) {
        cb.rec[(size_t)cta * cb.region_cap + slot] = c;
    } else {
        const unsigned g = atomicAdd(cb.common_count, 1u);
        if (g < cb.common_cap) cb.rec[(size_t)cb.n_ctas * cb.region_cap + g] = c;
        else cb.status[1] = 1u;  // out of space: CTAs that have not started yet give up at once
    }
}


// tensor_longk.cu
int tensor_longk_rows(int KB);
cudaError_t tensor_longk_launch(int KB, dim3 grid, cudaStream_t st, const unsigned char* qimage, int m, const unsigned char* rimage,
                                int ntiles, int tps, const float* band, unsigned* amin, const CandBuf& cb);

}  // namespace nns
