// lowk_search.cuh -- the hot path for k <= 32: register-blocked FP32 distance evaluation fused
// with a running (dist, idx) minimum; the m x n distance matrix is never materialised.
//
// Replaces v3/v4/v7/v8/v9::cudaCallKernel (reference core.cu:215-257, 307-349, 589-633, 716-760,
// 872-964).  Result semantics are V0's (core.cu:31-52): ascending-t FP32 accumulation, first
// minimum wins.
//
// Mapping.  A CTA owns 32*W*Q queries (W consumer warps, Q queries per thread held in registers
// for the whole kernel) and a contiguous range of reference blocks.  One extra producer warp
// streams reference tiles [TB][K][128] from the tiled-SoA index in HBM into a multi-stage
// shared-memory ring with 1-D bulk async copies (TMA, UBLKCP) completing on mbarriers.  Every
// consumer lane reads the same reference quad with one broadcast LDS.128 per dimension and
// evaluates 4 references x Q queries with packed FADD2/FMUL2/FFMA2 (two references per
// instruction, so the FP32 pipe is fed with half the issue slots and the argmin bookkeeping
// co-issues on the ALU pipe).
//
// Two-phase argmin.  Per reference quad only min(d0..d3) is formed (on the distance bit patterns: VIMNMX + VIMNMX3) and compared
// with the running best; the rare quad that improves a lane's best takes a divergent slow path
// that rescans the four distances in ascending index order with a strict '<', so the first
// (lowest-index) minimum is kept exactly as V0 does.  Each lane scans its whole reference range
// in ascending order; partial results of different CTAs (reference splits, other GPUs) are
// merged with an integer atomicMin on the packed (dist, idx) key, which breaks ties toward the
// lowest index as well.
#pragma once
#include "nns_common.cuh"
#include "nns_plan.h"

namespace nns {

struct LowkArgs {
    const float* queries;  // device AoS [m][k]
    int m;
    const float* index;    // device tiled SoA [nblocks][k][128]
    int nblocks;           // reference blocks in the index
    int blocks_per_split;  // reference blocks handled by one CTA (grid.y = splits)
    int index_base;        // global index of reference 0 of this index
    u64* keys;             // device [m]
    int warps;             // consumer warps per CTA (1..8)
    int stages;            // ring depth (2..LOWK_MAX_STAGES)
    int nqb;               // query blocks (grid.x)
    int splits;            // reference splits (grid.y)
    cudaStream_t stream;
};

// Distances of 4 references (one broadcast LDS.128 per dimension) x Q queries, two references per
// packed instruction: a01[i] = (d(q_i, r_j0), d(q_i, r_j0+1)), a23[i] = (.., r_j0+2), (.., r_j0+3).
// Ascending-t accumulation from the first product, like V0 (core.cu:38-43).
template <int K, int Q, bool EXACT>
__device__ __forceinline__ void lowk_quad_dist(const float* __restrict__ g, const u64 (&qq)[Q][K],
                                               u64 (&a01)[Q], u64 (&a23)[Q])
{
#pragma unroll
    for (int t = 0; t < K; ++t) {
        const ulonglong2 rv = lds_v2u64(g + t * LB);
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            const u64 d0 = sub2(rv.x, qq[i][t]);
            const u64 d1 = sub2(rv.y, qq[i][t]);
            if (t == 0) {
                a01[i] = mul2(d0, d0);
                a23[i] = mul2(d1, d1);
            } else if (EXACT) {
                a01[i] = add2(a01[i], mul2(d0, d0));
                a23[i] = add2(a23[i], mul2(d1, d1));
            } else {
                a01[i] = fma2(d0, d0, a01[i]);
                a23[i] = fma2(d1, d1, a23[i]);
            }
        }
    }
}

// Two-phase argmin over one quad.  Distances are >= +0 (or NaN = 0x7fffffff, or +INF), so their
// IEEE bit patterns order like unsigned integers and NaN sorts last: the bookkeeping runs entirely
// on the integer pipe (VIMNMX / VIMNMX3 / ISETP), which co-issues with the packed FP32 pipe.
// Fast path: min of the four patterns vs the running best.  Rare slow path: ascending rescan with
// a strict '<' so the first (lowest-index) minimum is kept exactly as V0 does (core.cu:44).
template <int Q>
__device__ __forceinline__ void lowk_quad_argmin(const u64 (&a01)[Q], const u64 (&a23)[Q], unsigned (&best)[Q],
                                                 int (&bidx)[Q], const int j0)
{
    unsigned mq[Q];
    bool any = false;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        unsigned x0, x1, x2, x3;
        upk2u(a01[i], x0, x1);
        upk2u(a23[i], x2, x3);
        mq[i] = min(min(x0, x1), min(x2, x3));
        any |= (mq[i] < best[i]);
    }
    if (any) {
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            if (mq[i] < best[i]) {
                unsigned x0, x1, x2, x3;
                upk2u(a01[i], x0, x1);
                upk2u(a23[i], x2, x3);
                if (x0 < best[i]) { best[i] = x0; bidx[i] = j0; }
                if (x1 < best[i]) { best[i] = x1; bidx[i] = j0 + 1; }
                if (x2 < best[i]) { best[i] = x2; bidx[i] = j0 + 2; }
                if (x3 < best[i]) { best[i] = x3; bidx[i] = j0 + 3; }
            }
        }
    }
}

// MINB = resident CTAs per SM the register allocation is held to, UNROLL = quads per loop body,
// PIPE = software-pipeline the argmin of quad g-1 under the distances of quad g (see below).
template <int K, int Q, bool EXACT, int MINB, int UNROLL, bool PIPE>
__global__ void __launch_bounds__(288, MINB)
lowk_search_kernel(const float* __restrict__ queries, const int m, const float* __restrict__ index,
                   const int nblocks, const int blocks_per_split, const int index_base,
                   const int stages, u64* __restrict__ keys)
{
    constexpr int TB = lowk_tb(K);
    constexpr int BLOCK_FLOATS = K * LB;
    constexpr int TILE_FLOATS = TB * BLOCK_FLOATS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw + LOWK_BAR_BYTES);

    const int W = (int)(blockDim.x >> 5) - 1;  // consumer warps; warp W is the producer
    const int warp = (int)(threadIdx.x >> 5);
    const int lane = (int)(threadIdx.x & 31);
    const int b0 = (int)blockIdx.y * blocks_per_split;
    const int b1 = min(nblocks, b0 + blocks_per_split);
    const int nb = b1 - b0;
    if (nb <= 0) return;  // uniform over the CTA
    const int ntiles = (nb + TB - 1) / TB;

    const uint32_t full0 = smem_u32(smem_raw);
    const uint32_t empty0 = full0 + 8u * LOWK_MAX_STAGES;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full0 + 8u * s, 1u);
            mbar_init(empty0 + 8u * s, (uint32_t)W);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == W) {
        // ---- producer warp: one lane streams tiles through the ring ----
        if (lane == 0) {
            const float* src = index + (size_t)b0 * BLOCK_FLOATS;
            int s = 0, round = 0;
            for (int tile = 0; tile < ntiles; ++tile) {
                if (round > 0) mbar_wait(empty0 + 8u * s, (uint32_t)((round - 1) & 1));
                const int tb = min(TB, nb - tile * TB);
                const uint32_t bytes = (uint32_t)tb * (uint32_t)(BLOCK_FLOATS * 4);
                mbar_arrive_expect_tx(full0 + 8u * s, bytes);
                bulk_g2s(smem_u32(tiles + (size_t)s * TILE_FLOATS),
                         src + (size_t)tile * TILE_FLOATS, bytes, full0 + 8u * s);
                if (++s == stages) { s = 0; ++round; }
            }
        }
        return;
    }

    // ---- consumer warps ----
    const int nct = W * 32;
    const int qbase = (int)blockIdx.x * (nct * Q) + (int)threadIdx.x;
    u64 qq[Q][K];
    unsigned best[Q];  // FP32 bit pattern of the running minimum
    int bidx[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const int q = qbase + i * nct;
        best[i] = 0x7f800000u;  // +INF
        bidx[i] = 0;
#pragma unroll
        for (int t = 0; t < K; ++t) {
            const float v = (q < m) ? __ldg(queries + (size_t)q * K + t) : nan_f();
            qq[i][t] = pk2(v, v);
        }
    }

    // Software pipeline (PIPE): the argmin bookkeeping of quad g-1 (integer pipe) is issued in the
    // same basic block as the distance evaluation of quad g (FP32 pipe), so the two pipes overlap
    // inside every warp instead of alternating.  `pa*` carry the pending quad; they start as
    // (+INF, +INF), which can never beat a running best.
    u64 pa01[Q], pa23[Q];
    int pj = 0;
#pragma unroll
    for (int i = 0; i < Q; ++i) { pa01[i] = 0x7f8000007f800000ull; pa23[i] = 0x7f8000007f800000ull; }

    int s = 0, round = 0;
    for (int tile = 0; tile < ntiles; ++tile) {
        mbar_wait(full0 + 8u * s, (uint32_t)(round & 1));
        const float* tsm = tiles + (size_t)s * TILE_FLOATS;
        const int tb = min(TB, nb - tile * TB);
        int j0 = index_base + (b0 + tile * TB) * LB;
        for (int b = 0; b < tb; ++b) {
            const float* blk = tsm + b * BLOCK_FLOATS;
#pragma unroll UNROLL
            for (int g = 0; g < LB / 4; ++g) {
                u64 a01[Q], a23[Q];
                lowk_quad_dist<K, Q, EXACT>(blk + 4 * g, qq, a01, a23);
                if (PIPE) {
                    lowk_quad_argmin<Q>(pa01, pa23, best, bidx, pj);
#pragma unroll
                    for (int i = 0; i < Q; ++i) { pa01[i] = a01[i]; pa23[i] = a23[i]; }
                    pj = j0 + 4 * g;
                } else {
                    lowk_quad_argmin<Q>(a01, a23, best, bidx, j0 + 4 * g);
                }
            }
            j0 += LB;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8u * s);
        if (++s == stages) { s = 0; ++round; }
    }
    if (PIPE) lowk_quad_argmin<Q>(pa01, pa23, best, bidx, pj);

#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const int q = qbase + i * nct;
        if (q < m && best[i] < 0x7f800000u) atomicMin(keys + q, ((u64)best[i] << 32) | (u64)(unsigned)bidx[i]);
    }
}

template <int K, int Q, bool EXACT, int MINB = lowk_minb(K, Q), int UNROLL = lowk_unroll(K), bool PIPE = lowk_pipe(K)>
cudaError_t lowk_launch_t(const LowkArgs& a, int* occupancy_out)
{
    auto kern = lowk_search_kernel<K, Q, EXACT, MINB, UNROLL, PIPE>;
    const size_t smem = (size_t)LOWK_BAR_BYTES + (size_t)a.stages * lowk_tile_bytes(K);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int threads = (a.warps + 1) * 32;
    if (occupancy_out) {
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occupancy_out, kern, threads, smem);
    }
    dim3 grid((unsigned)a.nqb, (unsigned)a.splits);
    kern<<<grid, threads, smem, a.stream>>>(a.queries, a.m, a.index, a.nblocks, a.blocks_per_split,
                                           a.index_base, a.stages, a.keys);
    return cudaGetLastError();
}

// dispatch over (k, q, exact); q must be lowk_q_default(k) or lowk_q_alt(k) (alt: FMA mode only).
// occupancy_out != NULL: do not launch, report resident CTAs per SM for that configuration.
cudaError_t lowk_launch(int k, int q, bool exact, const LowkArgs& a, int* occupancy_out);

}  // namespace nns
