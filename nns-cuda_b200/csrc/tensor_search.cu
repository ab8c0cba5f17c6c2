// tensor_search.cu -- the k > 32 path (BASELINE config C4, k = 128): the dense contraction
// -2 Q.R^T runs on the 5th-generation tensor cores (tcgen05.mma, BF16 operands, FP32 accumulators
// in TMEM), a fused epilogue reduces every 128-reference tile to a per-query minimum of the
// approximate score S~ = |r'|^2 - 2 q'.r' and emits (query, tile) candidates, and an exact FP32
// re-score of the candidate tiles in V0's subtract-square-accumulate form (core.cu:38-43) decides
// the answer.  The m x n score matrix never leaves TMEM.
//
// Exactness.  q' = fl(q - c), r' = fl(r - c) are the inputs centred on the reference mean c
// (distances are translation invariant; centring shrinks the operands ~4x for data in [0,1]).
// The tensor cores see bf16(-2q') and bf16(r'); E(q) bounds |S~ - S| for every reference plus the
// gap between V0's FP32 distance and the real one (tensor_band_kernel).  A tile is a candidate
// when its minimum S~ is within 2E of the running minimum, and is re-scored when it is within 2E
// of the final minimum: the tile holding V0's answer always qualifies, as does every tile holding
// an exactly tied reference, and the re-score keeps the lowest index through the packed-key
// atomicMin.  If the candidate buffer overflows (adversarial data: e.g. all points identical) a
// device flag makes the FP32 wide kernel redo the search -- there is no host round trip.
//
// Kernel structure (one CTA per 256-query strip x reference range, 10 warps):
//   warp 0   producer: 1-D bulk copies (TMA) of the pre-swizzled BF16 images: the strip's A tile
//            once (64 KiB), then the B tiles (32 KiB per 128 references) through a 4-stage ring
//   warp 1   MMA issuer: one elected lane issues tcgen05.mma.kind::f16 M=128 N=128 K=16, two
//            accumulator halves (query rows 0-127 / 128-255) x double-buffered = all 512 TMEM columns;
//            tcgen05.commit releases the B stage and publishes the accumulator
//   warps 2-9 epilogue: one thread per query row; tcgen05.ld 32 columns at a time (double
//            buffered), FMNMX3 running minimum, candidate test.  |r'|^2 is folded into the
//            contraction as one extra K = 16 step (A carries 1,1,1; B carries |r'|^2 split into three
//            BF16 terms), so the epilogue touches neither shared memory nor the FP32 pipe: shared
//            memory bandwidth is what the MMA operand fetch needs (M = N = 128: 128 B/clk).
// Operand images are K-major with the 128-byte swizzle (Swizzle<3,4,3>), written by the prep
// kernels exactly as the UMMA shared-memory descriptors expect them, so plain bulk copies suffice
// (no tensor maps).  SASS: UTCHMMA / LDTM / UBLKCP.
#include <cuda_bf16.h>

#include <algorithm>

#include "nns_internal.h"

namespace nns {

constexpr int T_BM = 256;     // query rows per CTA (two M = 128 accumulator halves)
constexpr int T_BN = 128;     // references per tile == one index block
constexpr int T_MAX_STAGES = 24;  // B ring depth is chosen per geometry: small tiles need a deep ring to cover the TMA latency
constexpr int T_THREADS = 320;
constexpr int T_EPI_WARPS = 8;

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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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
// instruction descriptor: D = F32, A = B = BF16, both K-major, N = 128, M = 128
constexpr uint32_t T_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(T_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

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
};
__host__ __device__ inline TensorGeom tensor_geom(int k)
{
    TensorGeom g;
    g.split = k <= TENSOR_SPLIT_MAX_K ? 1 : 0;
    g.ndata = g.split ? 3 * k : k;
    if (g.ndata + 3 <= 16) { g.KB = 0; g.KS = 1; g.norm_col = g.ndata; }
    else if (g.ndata + 3 <= 32) { g.KB = 0; g.KS = 2; g.norm_col = g.ndata; }
    else if (g.ndata <= 64) { g.KB = 1; g.KS = 1; g.norm_col = 64; }
    else { g.KB = 2; g.KS = 1; g.norm_col = 128; }
    return g;
}
__host__ __device__ constexpr size_t image_bytes(int rows, int KB, int KS) { return (size_t)rows * (KB * 128 + KS * 32); }
// byte offset of the 16-byte chunk holding columns [8*chunk, 8*chunk + 8) of `row`
__device__ __forceinline__ size_t image_chunk_at(int rows, int KB, int row, int chunk)
{
    if (chunk < KB * 8) return image_chunk_offset(rows, row, chunk >> 3, chunk & 7);
    return (size_t)KB * rows * 128 + (size_t)(chunk - KB * 8) * rows * 16 + (size_t)row * 16;
}

// Split precision (k <= TENSOR_SPLIT_MAX_K).  BF16 keeps 8 significant bits; splitting each centred
// coordinate into hi = bf16(x) and lo = bf16(x - hi) and laying the contraction dimension out as
//     queries:    [ qh (k) | qh (k) | ql (k) ]        references: [ rh (k) | rl (k) | rh (k) ]
// makes one BF16 MMA pass accumulate qh.rh + qh.rl + ql.rh, i.e. q'.r' up to terms of relative size
// ~3 * 2^-18: the screen's error bound E shrinks ~250x, so that it is selective even for k = 3 with
// millions of references (nearest-neighbour distances ~1e-5 of the data extent).
__host__ __device__ constexpr bool tensor_split(int k) { return k <= TENSOR_SPLIT_MAX_K; }
// source dimension and part (0 = hi, 1 = lo) of data column `col` (< ndata); dimension -1 = none
__device__ __forceinline__ void image_column(int k, int ndata, int col, bool query, int& dim, int& part)
{
    if (col >= ndata) { dim = -1; part = 0; return; }
    if (!tensor_split(k)) { dim = col; part = 0; return; }
    const int seg = col / k;
    dim = col - seg * k;
    part = query ? (seg == 2 ? 1 : 0) : (seg == 1 ? 1 : 0);
}
__device__ __forceinline__ __nv_bfloat16 bf16_part(float x, int part)
{
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    if (part == 0) return hi;
    const float rem = x - __bfloat162float(hi);  // exact in FP32
    return __float2bfloat16_rn((fabsf(x) < inf_f()) ? rem : 0.0f);
}

// ---------------------------------------------------------------------------------------------
// reference-side preparation (part of index_build for 32 < k <= 128)
// ---------------------------------------------------------------------------------------------
// section header (floats): [0..127] centre, [128] max |r'|^2 (bits), [129] flags (bit 0: unusable)
__global__ void tensor_colsum_kernel(const float* __restrict__ aos, const int n, const int k, float* __restrict__ sums)
{
    // grid.x blocks of 256 points; thread t < k sums its dimension over the block's points
    const long long j0 = (long long)blockIdx.x * 256;
    const int jn = (int)min((long long)256, n - j0);
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        float s = 0.0f;
        for (int j = 0; j < jn; ++j) {
            const float x = __ldg(aos + (j0 + j) * k + t);
            if (fabsf(x) <= 1e15f) s += x;  // NaN / INF / huge coordinates do not steer the centre
        }
        atomicAdd(sums + t, s);
    }
}

__global__ void tensor_centre_kernel(float* __restrict__ hdr, const int n, const int k)
{
    const int t = threadIdx.x;
    if (t < 128) {
        float c = (t < k && n > 0) ? hdr[t] / (float)n : 0.0f;
        if (!(fabsf(c) <= 1e15f)) {  // NaN / INF / huge input: the tensor path is disabled for this index
            c = 0.0f;
            atomicOr(reinterpret_cast<unsigned*>(hdr) + 129, 1u);
        }
        hdr[t] = c;
    }
}

// one CTA per 128-reference block: BF16 image of r' = fl(r - c) (TensorGeom layout) with |r'|^2
// (FP32, split into three BF16 terms; +INF for padded lanes) in columns norm_col .. norm_col + 2
__global__ void __launch_bounds__(128)
tensor_ref_image_kernel(const float* __restrict__ aos, const int n, const int k, const TensorGeom g,
                        float* __restrict__ hdr, unsigned char* __restrict__ image)
{
    const long long b = blockIdx.x;
    const int row = threadIdx.x;  // one thread per reference
    const long long j = b * T_BN + row;
    const bool valid = j < n;
    unsigned char* img = image + (size_t)b * image_bytes(T_BN, g.KB, g.KS);
    // |r'|^2 first (ascending dimensions), because its columns may share a chunk with data columns
    float rn = 0.0f;
    bool bad = false;
    for (int t = 0; t < k; ++t) {
        float x = 0.0f;
        if (valid) {
            x = __fsub_rn(__ldg(aos + j * k + t), hdr[t]);
            // a NaN / INF coordinate only poisons its own column (that reference cannot win in
            // V0 either); finite but huge values would overflow the error-bound arithmetic
            if (fabsf(x) > 1e15f && fabsf(x) < inf_f()) bad = true;
        }
        rn = __fmaf_rn(x, x, rn);
    }
    const float rv = valid ? rn : inf_f();  // padded lanes can never be a tile minimum
    const __nv_bfloat16 n_hi = __float2bfloat16_rn(rv);
    const float rem1 = (rv < inf_f()) ? rv - __bfloat162float(n_hi) : 0.0f;
    const __nv_bfloat16 n_mid = __float2bfloat16_rn(rem1);
    const __nv_bfloat16 n_lo = __float2bfloat16_rn(rem1 - __bfloat162float(n_mid));
    const int chunks = g.KB * 8 + g.KS * 2;
    for (int ch = 0; ch < chunks; ++ch) {
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int col = ch * 8 + e;
            int dim, part;
            image_column(k, g.ndata, col, false, dim, part);
            float x = 0.0f;
            if (valid && dim >= 0) x = __fsub_rn(__ldg(aos + j * k + dim), hdr[dim]);
            __nv_bfloat16 o = bf16_part(x, part);
            if (col == g.norm_col) o = n_hi;
            if (col == g.norm_col + 1) o = n_mid;
            if (col == g.norm_col + 2) o = n_lo;
            v[e] = o;
        }
        *reinterpret_cast<uint4*>(img + image_chunk_at(T_BN, g.KB, row, ch)) = *reinterpret_cast<const uint4*>(v);
    }
    unsigned bits = (valid && rn < inf_f()) ? __float_as_uint(rn) : 0u;  // NaN / INF norms excluded
    bits = __reduce_max_sync(0xffffffffu, bits);
    if ((threadIdx.x & 31) == 0 && bits) atomicMax(reinterpret_cast<unsigned*>(hdr) + 128, bits);
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(reinterpret_cast<unsigned*>(hdr) + 129, 1u);
}

// ---------------------------------------------------------------------------------------------
// query-side preparation (per search call)
// ---------------------------------------------------------------------------------------------
// one CTA per 256-query strip: BF16 image of -2 q' (TensorGeom layout, 1 in the norm columns),
// band[q] = 2 E(q), approx_min[q] = +INF (ordered encoding)
__global__ void __launch_bounds__(256)
tensor_query_image_kernel(const float* __restrict__ queries, const int m, const int k, const TensorGeom g,
                          const float* __restrict__ hdr, unsigned char* __restrict__ image,
                          float* __restrict__ band, unsigned* __restrict__ approx_min)
{
    const int row = threadIdx.x;
    const long long q = (long long)blockIdx.x * T_BM + row;
    const bool valid = q < m;
    const int KP = g.KB * 64 + g.KS * 16;  // contraction length seen by the MMA
    unsigned char* img = image + (size_t)blockIdx.x * image_bytes(T_BM, g.KB, g.KS);
    float qn = 0.0f;
    for (int t = 0; t < k; ++t) {
        const float x = valid ? __fsub_rn(__ldg(queries + q * k + t), hdr[t]) : 0.0f;
        qn = __fmaf_rn(x, x, qn);
    }
    const int chunks = g.KB * 8 + g.KS * 2;
    for (int ch = 0; ch < chunks; ++ch) {
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int col = ch * 8 + e;
            int dim, part;
            image_column(k, g.ndata, col, true, dim, part);
            float x = 0.0f;
            if (valid && dim >= 0) x = __fsub_rn(__ldg(queries + q * k + dim), hdr[dim]);
            __nv_bfloat16 o = bf16_part(-2.0f * x, part);
            if (col >= g.norm_col && col < g.norm_col + 3) o = __float2bfloat16_rn(1.0f);
            v[e] = o;
        }
        *reinterpret_cast<uint4*>(img + image_chunk_at(T_BM, g.KB, row, ch)) = *reinterpret_cast<const uint4*>(v);
    }
    if (valid) {
        // E(q) >= |S~ - S| + |d_V0 - D'| for every reference (see the file header):
        //   bf16 rounding of both operands   2^-7 (1 + 2^-9) |q'| |r'|
        //   FP32 accumulation in the MMA     K 2^-23 * 2.02 |q'| |r'|
        //   FP32 |r'|^2 and its 3-term split (K+1) 2^-24 |r'|^2 + 2^-22 |r'|^2
        //   centring + V0 rounding           (K+8) 2^-24 (|q'| + |r'|)^2
        const float r2 = __uint_as_float(reinterpret_cast<const unsigned*>(hdr)[128]);
        const float a = sqrtf(qn), rmax = sqrtf(r2);
        const float u24 = 5.9604645e-8f;
        // operand rounding: plain BF16 2^-7 (1 + 2^-9) |q'||r'|; split precision drops only
        // ql.rl and the second-order remainders: 2 * 3.1 * 2^-18 |q'||r'|.  The MMA's FP32
        // accumulation is charged 2^-21 per term (truncating adders).
        const float c_round = tensor_split(k) ? 6.2f * 3.8146973e-6f : 0.0078125f * 1.002f;
        float E = (c_round + (float)KP * 2.04f * 4.7683716e-7f) * a * rmax + (KP + 5) * u24 * r2 +
                  (KP + 8) * u24 * (a + rmax) * (a + rmax);
        E *= 1.05f;
        const bool flagged = (reinterpret_cast<const unsigned*>(hdr)[129] & 1u) != 0;  // NaN / INF / huge references
        const bool usable = !flagged && (qn <= 1e30f) && (r2 <= 1e30f);  // false for NaN too
        band[q] = usable ? 2.0f * E : inf_f();
        approx_min[q] = f2ord(inf_f());
    }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
struct TensorCand { int q; int unit; float smin; };  // unit = 32 consecutive references (tile * 4 + chunk)

template <int KB, int KS, int T_STAGES>
__global__ void __launch_bounds__(T_THREADS, 1)
tensor_screen_kernel(const unsigned char* __restrict__ qimage, const int m, const unsigned char* __restrict__ rimage,
                     const int ntiles, const int tiles_per_split,
                     const float* __restrict__ band, unsigned* __restrict__ approx_min,
                     TensorCand* __restrict__ cand, unsigned* __restrict__ cand_count, const unsigned cand_cap)
{
    // KB 64-column swizzled blocks (one 128-byte swizzle row each), then KS interleaved 16-column steps
    constexpr uint32_t A_MAIN = KB * T_BM * 128, B_MAIN = KB * T_BN * 128;
    constexpr uint32_t A_BYTES = (uint32_t)image_bytes(T_BM, KB, KS);  // 72 KiB at KB = 2, KS = 1
    constexpr uint32_t B_BYTES = (uint32_t)image_bytes(T_BN, KB, KS);  // 36 KiB at KB = 2, KS = 1
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_smem = smem;
    unsigned char* b_smem = smem + A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + T_STAGES * B_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * T_MAX_STAGES + 8);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t b_full = bar0, b_empty = bar0 + 8 * T_STAGES;
    const uint32_t acc_full = bar0 + 8 * 2 * T_STAGES, acc_empty = acc_full + 16, a_full = acc_empty + 16;

    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const int t0 = (int)blockIdx.y * tiles_per_split;
    const int nt = min(ntiles, t0 + tiles_per_split) - t0;
    if (nt <= 0) return;

    if (threadIdx.x == 0) {
        for (int s = 0; s < T_STAGES; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, T_EPI_WARPS); }
        mbar_init(a_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc512(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- producer ----------------
        if (lane == 0) {
            mbar_arrive_expect_tx(a_full, A_BYTES);
            bulk_g2s(smem_u32(a_smem), qimage + (size_t)blockIdx.x * A_BYTES, A_BYTES, a_full);
            for (int t = 0; t < nt; ++t) {
                const int s = t % T_STAGES;
                mbar_wait(b_empty + 8 * s, (uint32_t)(((t / T_STAGES) & 1) ^ 1));
                mbar_arrive_expect_tx(b_full + 8 * s, B_BYTES);
                bulk_g2s(smem_u32(b_smem + (size_t)s * B_BYTES), rimage + (size_t)(t0 + t) * B_BYTES, B_BYTES, b_full + 8 * s);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            mbar_wait(a_full, 0);
            for (int t = 0; t < nt; ++t) {
                const int s = t % T_STAGES, buf = t & 1;
                mbar_wait(acc_empty + 8 * buf, (uint32_t)(((t >> 1) & 1) ^ 1));  // epilogue drained this buffer
                mbar_wait(b_full + 8 * s, (uint32_t)((t / T_STAGES) & 1));       // TMA landed this stage
                tc_fence_after();
                const uint32_t a_addr = smem_u32(a_smem), b_addr = smem_u32(b_smem + (size_t)s * B_BYTES);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const u64 bdesc = umma_desc_sw128(b_addr + kb * (T_BN * 128) + ks * 32);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const u64 adesc = umma_desc_sw128(a_addr + kb * (T_BM * 128) + h * (128 * 128) + ks * 32);
                            tc_mma_bf16(tmem_base + (uint32_t)((buf * 2 + h) * T_BN), adesc, bdesc, T_IDESC, (uint32_t)((kb | ks) != 0));
                        }
                    }
                }
#pragma unroll
                for (int x = 0; x < KS; ++x) {  // interleaved steps (the last columns carry |r'|^2)
                    const u64 bdesc = umma_desc_interleave(b_addr + B_MAIN + x * (2 * T_BN * 16), T_BN);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const u64 adesc = umma_desc_interleave(a_addr + A_MAIN + x * (2 * T_BM * 16) + h * (128 * 16), T_BM);
                        tc_mma_bf16(tmem_base + (uint32_t)((buf * 2 + h) * T_BN), adesc, bdesc, T_IDESC, (uint32_t)((KB | x) != 0));
                    }
                }
                tc_commit(b_empty + 8 * s);      // stage free once these MMAs have read it
                tc_commit(acc_full + 8 * buf);   // accumulator complete
            }
        }
    } else {
        // ---------------- epilogue: thread = query row ----------------
        const int e = warp - 2;                 // 0..7
        const int lq = warp & 3;                // TMEM lane quarter this warp may access
        const int half = e >> 2;                // accumulator half (rows 0-127 / 128-255)
        const int row = half * 128 + lq * 32 + lane;
        const long long q = (long long)blockIdx.x * T_BM + row;
        const float my_band = (q < m) ? band[q] : -inf_f();  // rows past m never qualify
        // other CTAs (reference splits, earlier waves) may already have lowered this query's minimum
        float run_min = (q < m) ? ord2f(approx_min[q]) : inf_f();
        for (int t = 0; t < nt; ++t) {
            const int buf = t & 1;
            mbar_wait(acc_full + 8 * buf, (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)((buf * 2 + half) * T_BN);
            // all four 32-column loads are issued back to back and waited for once: the epilogue of a
            // tile costs one TMEM-load latency instead of four
            uint32_t v[T_BN / 32][32];
            float cmin[T_BN / 32];
#pragma unroll
            for (int c = 0; c < T_BN / 32; ++c) tmem_ld32(taddr + c * 32, v[c]);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < T_BN / 32; ++c) {
                const uint32_t (&cur)[32] = v[c];
                // four independent FMNMX3 chains (depth 4) + a 2-level combine instead of one chain of 16
                float c0 = min3(__uint_as_float(cur[0]), __uint_as_float(cur[1]), __uint_as_float(cur[2]));
                float c1 = min3(__uint_as_float(cur[8]), __uint_as_float(cur[9]), __uint_as_float(cur[10]));
                float c2 = min3(__uint_as_float(cur[16]), __uint_as_float(cur[17]), __uint_as_float(cur[18]));
                float c3 = min3(__uint_as_float(cur[24]), __uint_as_float(cur[25]), __uint_as_float(cur[26]));
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    c0 = min3(c0, __uint_as_float(cur[3 + 2 * j]), __uint_as_float(cur[4 + 2 * j]));
                    c1 = min3(c1, __uint_as_float(cur[11 + 2 * j]), __uint_as_float(cur[12 + 2 * j]));
                    c2 = min3(c2, __uint_as_float(cur[19 + 2 * j]), __uint_as_float(cur[20 + 2 * j]));
                    c3 = min3(c3, __uint_as_float(cur[27 + 2 * j]), __uint_as_float(cur[28 + 2 * j]));
                }
                c0 = fminf(c0, __uint_as_float(cur[7]));
                c1 = fminf(c1, __uint_as_float(cur[15]));
                c2 = fminf(c2, __uint_as_float(cur[23]));
                c3 = fminf(c3, __uint_as_float(cur[31]));
                cmin[c] = fminf(min3(c0, c1, c2), c3);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + 8 * buf);
            // candidates at 32-reference granularity (one TMEM chunk): 4x less to re-score than a
            // tile.  One test per tile on the fast path; the per-unit tests only when it fires.
            if (fminf(fminf(cmin[0], cmin[1]), fminf(cmin[2], cmin[3])) <= run_min + my_band) {
#pragma unroll
                for (int c = 0; c < T_BN / 32; ++c) {
                    if (cmin[c] <= run_min + my_band) {
                        const unsigned slot = atomicAdd(cand_count, 1u);
                        if (slot < cand_cap) {
                            TensorCand cnd;
                            cnd.q = (int)q; cnd.unit = (t0 + t) * (T_BN / 32) + c; cnd.smin = cmin[c];
                            cand[slot] = cnd;
                        }
                        if (cmin[c] < run_min) {
                            run_min = cmin[c];
                            atomicMin(approx_min + q, f2ord(run_min));
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc512(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// exact re-score: one warp per candidate (query, 128-reference block)
// ---------------------------------------------------------------------------------------------
template <bool EXACT>
__global__ void __launch_bounds__(256)
tensor_rescore_kernel(const float* __restrict__ queries, const int k, const float* __restrict__ blocks,
                      const int index_base, const TensorCand* __restrict__ cand, const unsigned* __restrict__ cand_count,
                      const unsigned cand_cap, const float* __restrict__ band, const unsigned* __restrict__ approx_min,
                      u64* __restrict__ keys, int* __restrict__ overflow)
{
    const unsigned total = *cand_count;
    if (total > cand_cap) {  // the wide kernel takes over (launched right after with this flag)
        if (blockIdx.x == 0 && threadIdx.x == 0) *overflow = 1;
        return;
    }
    const int lane = (int)(threadIdx.x & 31);
    const unsigned wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    // each lane tests one candidate against the final minimum; survivors are then re-scored one at
    // a time by the whole warp, one lane per reference of the 32-reference unit
    for (unsigned base = wid * 32; base < total; base += nw * 32) {
        const unsigned ci = base + lane;
        TensorCand mine;
        mine.q = 0; mine.unit = 0; mine.smin = 0.0f;
        bool live = false;
        if (ci < total) {
            mine = cand[ci];
            live = mine.smin <= ord2f(approx_min[mine.q]) + band[mine.q];
        }
        unsigned mask = __ballot_sync(0xffffffffu, live);
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const int cq = __shfl_sync(0xffffffffu, mine.q, src);
            const int unit = __shfl_sync(0xffffffffu, mine.unit, src);
            const int jl = unit * 32 + lane;  // reference index within this index
            const float* col = blocks + (size_t)(jl >> 7) * (k + 1) * LB + (jl & (LB - 1));
            const float* qp = queries + (size_t)cq * k;
            float d = 0.0f;
            for (int tb = 0; tb < k; tb += 32) {
                const float qv = (tb + lane < k) ? __ldg(qp + tb + lane) : 0.0f;
                const int te = min(32, k - tb);
#pragma unroll 8
                for (int tt = 0; tt < te; ++tt) {
                    const float qt = __shfl_sync(0xffffffffu, qv, tt);
                    const float e = qt - __ldg(col + (size_t)(tb + tt) * LB);
                    d = EXACT ? __fadd_rn(d, __fmul_rn(e, e)) : __fmaf_rn(e, e, d);
                }
            }
            u64 key = (d < inf_f()) ? pack_key(d, index_base + jl) : KEY_INIT;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const u64 o = __shfl_xor_sync(0xffffffffu, key, off);
                key = o < key ? o : key;
            }
            if (lane == 0 && key < KEY_INIT) atomicMin(keys + cq, key);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int tensor_kp(int k)
{
    const TensorGeom g = tensor_geom(k);
    return g.KB * 64 + g.KS * 16;
}

size_t tensor_section_floats(int k, int n)
{
    if (k < 1 || k > TENSOR_MAX_K || n <= 0) return 0;
    const size_t nblocks = (size_t)((n + LB - 1) / LB);
    const TensorGeom g = tensor_geom(k);
    return (size_t)TENSOR_HDR_FLOATS + nblocks * image_bytes(T_BN, g.KB, g.KS) / 4;
}

cudaError_t tensor_index_build(int k, int n, const float* d_refs_aos, float* d_section, cudaStream_t st)
{
    if (tensor_section_floats(k, n) == 0) return cudaSuccess;
    const TensorGeom g = tensor_geom(k);
    const int nblocks = (n + LB - 1) / LB;
    float* hdr = d_section;
    unsigned char* image = reinterpret_cast<unsigned char*>(d_section + TENSOR_HDR_FLOATS);
    cudaError_t e = cudaMemsetAsync(hdr, 0, TENSOR_HDR_FLOATS * sizeof(float), st);
    if (e != cudaSuccess) return e;
    tensor_colsum_kernel<<<(n + 255) / 256, 128, 0, st>>>(d_refs_aos, n, k, hdr);
    tensor_centre_kernel<<<1, 128, 0, st>>>(hdr, n, k);
    tensor_ref_image_kernel<<<nblocks, 128, 0, st>>>(d_refs_aos, n, k, g, hdr, image);
    return cudaGetLastError();
}

int tensor_stages(const TensorGeom& g) { return g.KB == 2 ? 4 : g.KB == 1 ? 6 : g.KS == 2 ? 16 : 24; }

size_t tensor_smem_bytes(const TensorGeom& g)
{
    return image_bytes(T_BM, g.KB, g.KS) + (size_t)tensor_stages(g) * image_bytes(T_BN, g.KB, g.KS) +
           (2 * T_MAX_STAGES + 8) * 8 + 16;
}

template <int KB, int KS, int STAGES>
static cudaError_t tensor_screen_launch(dim3 grid, size_t smem, cudaStream_t st, const unsigned char* qimage, int m,
                                        const unsigned char* rimage, int ntiles, int tps, const float* band, unsigned* amin,
                                        TensorCand* cand, unsigned* cnt, unsigned cand_cap)
{
    cudaError_t e = cudaFuncSetAttribute(tensor_screen_kernel<KB, KS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tensor_screen_kernel<KB, KS, STAGES><<<grid, T_THREADS, smem, st>>>(qimage, m, rimage, ntiles, tps, band, amin, cand, cnt, cand_cap);
    return cudaGetLastError();
}

// Search m queries against the n references of the index section; accumulates into keys.
// Returns the number of kernels launched through *launches.
cudaError_t tensor_search(int k, int m, int n, const float* d_queries, const float* d_blocks, const float* d_section,
                          int index_base, u64* d_keys, bool exact, int num_sms, cudaStream_t st, int* launches,
                          unsigned* d_stats, bool tiny_candidate_buffer)
{
    const TensorGeom g = tensor_geom(k);
    const int nblocks = (n + LB - 1) / LB;
    const int strips = (m + T_BM - 1) / T_BM;
    const float* hdr = d_section;
    const unsigned char* rimage = reinterpret_cast<const unsigned char*>(d_section + TENSOR_HDR_FLOATS);

    // reference splits: minimise (waves of one CTA per SM) x (tiles per CTA); every extra split
    // costs each query one more seed candidate, so ties go to fewer splits
    int splits = 1;
    {
        double best = 1e300;
        const int smax = std::min(nblocks, 64);
        for (int sp = 1; sp <= smax; ++sp) {
            const int t = (nblocks + sp - 1) / sp;
            const int se = (nblocks + t - 1) / t;
            const double waves = (double)(((long long)strips * se + num_sms - 1) / num_sms);
            const double cost = waves * ((double)t + 24.0);  // + per-CTA prologue (A tile, TMEM alloc) in tile units
            if (cost < best * 0.97) { best = cost; splits = se; }
        }
    }
    const int tps = (nblocks + splits - 1) / splits;
    splits = (nblocks + tps - 1) / tps;

    // stream-ordered scratch: query image, band, approx_min, candidates, counters.  Every split of
    // a strip emits at least its first tile per query, then running-minimum records + the band.
    const size_t qimg_bytes = (size_t)strips * image_bytes(T_BM, g.KB, g.KS);
    const unsigned cand_cap = tiny_candidate_buffer
                                  ? 64u  // test hook: forces the overflow -> wide-kernel fallback
                                  : (unsigned)std::min<size_t>((size_t)m * (64 + 6 * (size_t)splits) + 65536, (size_t)1 << 30);
    const size_t off_band = (qimg_bytes + 255) & ~(size_t)255;
    const size_t off_amin = off_band + (((size_t)m * 4 + 255) & ~(size_t)255);
    const size_t off_cnt = off_amin + (((size_t)m * 4 + 255) & ~(size_t)255);
    const size_t off_cand = off_cnt + 256;
    const size_t total = off_cand + (size_t)cand_cap * sizeof(TensorCand);
    unsigned char* scratch = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&scratch, total, st);
    if (e != cudaSuccess) return e;
    float* band = reinterpret_cast<float*>(scratch + off_band);
    unsigned* amin = reinterpret_cast<unsigned*>(scratch + off_amin);
    unsigned* cnt = reinterpret_cast<unsigned*>(scratch + off_cnt);
    int* overflow = reinterpret_cast<int*>(cnt + 1);
    TensorCand* cand = reinterpret_cast<TensorCand*>(scratch + off_cand);

    e = cudaMemsetAsync(cnt, 0, 256, st);
    if (e == cudaSuccess) {
        tensor_query_image_kernel<<<strips, 256, 0, st>>>(d_queries, m, k, g, hdr, scratch, band, amin);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        // every CTA allocates all 512 TMEM columns: ask for enough shared memory that only one
        // CTA is resident per SM even at KP = 64
        const size_t smem = std::max(tensor_smem_bytes(g), (size_t)120 * 1024);
        dim3 grid((unsigned)strips, (unsigned)splits);
        if (g.KB == 0 && g.KS == 1) e = tensor_screen_launch<0, 1, 24>(grid, smem, st, scratch, m, rimage, nblocks, tps, band, amin, cand, cnt, cand_cap);
        else if (g.KB == 0) e = tensor_screen_launch<0, 2, 16>(grid, smem, st, scratch, m, rimage, nblocks, tps, band, amin, cand, cnt, cand_cap);
        else if (g.KB == 1) e = tensor_screen_launch<1, 1, 6>(grid, smem, st, scratch, m, rimage, nblocks, tps, band, amin, cand, cnt, cand_cap);
        else e = tensor_screen_launch<2, 1, 4>(grid, smem, st, scratch, m, rimage, nblocks, tps, band, amin, cand, cnt, cand_cap);
    }
    if (e == cudaSuccess) {
        const int rgrid = num_sms * 8;
        if (exact)
            tensor_rescore_kernel<true><<<rgrid, 256, 0, st>>>(d_queries, k, d_blocks, index_base, cand, cnt, cand_cap, band, amin, d_keys, overflow);
        else
            tensor_rescore_kernel<false><<<rgrid, 256, 0, st>>>(d_queries, k, d_blocks, index_base, cand, cnt, cand_cap, band, amin, d_keys, overflow);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        // overflow fallback: the wide kernel runs only if the device flag is set
        WideArgs a{};
        a.queries = d_queries; a.m = m; a.k = k; a.blocks = d_blocks; a.nblocks = nblocks;
        a.nqg = (m + WIDE_QT - 1) / WIDE_QT;
        int s = 1;
        if (a.nqg < 4 * num_sms * 8) s = std::max(1, std::min((nblocks + 1) / 2, (4 * num_sms * 8 + a.nqg - 1) / a.nqg));
        if (s > 65535) s = 65535;
        a.blocks_per_split = (nblocks + s - 1) / s;
        a.splits = (nblocks + a.blocks_per_split - 1) / a.blocks_per_split;
        a.index_base = index_base; a.keys = d_keys; a.stream = st; a.enable = overflow;
        e = wide_launch(exact, a);
    }
    if (launches) *launches = 4;
    // diagnostics: [0] candidates emitted, [1] overflow flag, [2] candidate capacity
    if (e == cudaSuccess && d_stats) {
        e = cudaMemcpyAsync(d_stats, cnt, 2 * sizeof(unsigned), cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_stats + 2, &cand_cap, sizeof(unsigned), cudaMemcpyHostToDevice, st);
    }
    cudaError_t e2 = cudaFreeAsync(scratch, st);
    return e != cudaSuccess ? e : e2;
}

// 1 if the index section says the tensor path must not be used (non-finite / huge inputs)
__global__ void tensor_flag_kernel(const float* hdr, int* out) { *out = (int)(reinterpret_cast<const unsigned*>(hdr)[129] & 1u); }

}  // namespace nns
