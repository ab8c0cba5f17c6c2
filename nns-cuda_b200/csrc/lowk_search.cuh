// lowk_search.cuh -- the hot path for k <= 32: register-blocked FP32 distance evaluation fused
// with a running (dist, idx) minimum; the m x n distance matrix is never materialised.
//
// Replaces v3/v4/v7/v8/v9::cudaCallKernel (reference core.cu:215-257, 307-349, 589-633, 716-760,
// 872-964).  Result semantics are V0's (core.cu:31-52): ascending-t FP32 accumulation, first
// minimum wins.
//
// Mapping.  A CTA owns 32*W*Q queries (W consumer warps, Q queries per thread held in registers
// for the whole kernel) and a contiguous range of reference blocks.  One extra producer warp
// streams reference tiles [TB][K+1][128] from the tiled-SoA index in HBM into a multi-stage
// shared-memory ring with 1-D bulk async copies (TMA, UBLKCP) completing on mbarriers.  Every
// consumer lane reads the same reference quad with one broadcast LDS.128 per dimension and
// evaluates 4 references x Q queries with packed FADD2/FMUL2/FFMA2 (two references per
// instruction).
//
// Two kernels share that skeleton:
//
//  * lowk_exact_kernel  -- V0's formulation for every pair: k packed subtractions + k packed
//    multiply-adds (2k FP32 lane-slots per pair), then a two-phase argmin on the integer pipe:
//    per quad only min(d0..d3) of the distance bit patterns is formed (VIMNMX/VIMNMX3) and
//    compared with the running best; the rare quad that improves a lane's best takes a divergent
//    slow path that rescans the four distances in ascending index order with a strict '<'.
//    With EXACT it rounds mul and add separately (bit-identical to V0's distances).
//
//  * lowk_filter_kernel -- screens every pair with the norm expansion
//        s_j = |r_j|^2 - 2 q.r_j            (k packed FMAs = k lane-slots per pair)
//    against a per-query threshold tau = best + E - |q|^2, where E is a rigorous bound on the
//    rounding error of s (see lowk_filter_threshold).  Only pairs that pass -- a superset of the
//    pairs whose V0-form distance is below the running best -- are evaluated exactly, in V0's
//    subtract-square-accumulate form, by the slow path, which alone updates (best, idx).  The
//    returned indices are therefore the same as lowk_exact_kernel's on every input; the filter
//    halves the FP32 work per pair.  NaN passes the filter ('not greater' comparison), and
//    magnitudes large enough to overflow disable it (tau = NaN), so the exact path decides.
//
// Each lane scans its whole reference range in ascending order; partial results of different
// CTAs (reference splits, other GPUs) are merged with an integer atomicMin on the packed
// (dist, idx) key, which breaks ties toward the lowest index as well.
#pragma once
#include "nns_common.cuh"
#include "nns_plan.h"

namespace nns {

struct LowkArgs {
    const float* queries;  // device AoS [m][k]
    int m;
    const float* header;   // index header (INDEX_HEADER_FLOATS floats)
    const float* blocks;   // first reference block of the range to search: [nblocks][k+1][128]
    int nblocks;           // reference blocks in the range
    int blocks_per_split;  // reference blocks handled by one CTA (grid.y = splits)
    int index_base;        // global index of the first reference of the range
    u64* keys;             // device [m]
    int warps;             // consumer warps per CTA (1..8)
    int stages;            // ring depth (2..LOWK_MAX_STAGES)
    int nqb;               // query blocks (grid.x)
    int splits;            // reference splits (grid.y)
    cudaStream_t stream;
    const int* enable = nullptr;  // optional device flag: the kernel exits immediately when it is 0
};

// ---------------------------------------------------------------------------------------------
// shared skeleton: ring of reference tiles
// ---------------------------------------------------------------------------------------------
template <int K>
struct TileRing {
    static constexpr int TB = lowk_tb(K);
    static constexpr int BLOCK_FLOATS = index_block_floats(K);
    static constexpr int TILE_FLOATS = TB * BLOCK_FLOATS;
    uint32_t full0, empty0;
    float* tiles;
    int stages;

    __device__ __forceinline__ void init(unsigned char* smem_raw, int stages_, int consumer_warps)
    {
        stages = stages_;
        tiles = reinterpret_cast<float*>(smem_raw + LOWK_BAR_BYTES);
        full0 = smem_u32(smem_raw);
        empty0 = full0 + 8u * LOWK_MAX_STAGES;
        if (threadIdx.x == 0) {
            for (int s = 0; s < stages; ++s) {
                mbar_init(full0 + 8u * s, 1u);
                mbar_init(empty0 + 8u * s, (uint32_t)consumer_warps);
            }
            mbar_fence_init();
        }
        __syncthreads();
    }
    // producer: one lane streams `nb` blocks starting at `src` through the ring
    __device__ __forceinline__ void produce(const float* __restrict__ src, int nb) const
    {
        const int ntiles = (nb + TB - 1) / TB;
        int s = 0, round = 0;
        for (int tile = 0; tile < ntiles; ++tile) {
            if (round > 0) mbar_wait(empty0 + 8u * s, (uint32_t)((round - 1) & 1));
            const int tb = min(TB, nb - tile * TB);
            const uint32_t bytes = (uint32_t)tb * (uint32_t)(BLOCK_FLOATS * 4);
            mbar_arrive_expect_tx(full0 + 8u * s, bytes);
            bulk_g2s(smem_u32(tiles + (size_t)s * TILE_FLOATS), src + (size_t)tile * TILE_FLOATS, bytes,
                     full0 + 8u * s);
            if (++s == stages) { s = 0; ++round; }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// exact-form kernel
// ---------------------------------------------------------------------------------------------
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
// IEEE bit patterns order like unsigned integers and NaN sorts last: the bookkeeping runs on the
// integer pipe (VIMNMX / VIMNMX3 / ISETP).  Fast path: min of the four patterns vs the running
// best.  Rare slow path: ascending rescan with a strict '<' so the first (lowest-index) minimum
// is kept exactly as V0 does (core.cu:44).
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

// MINB = resident CTAs per SM the register allocation is held to, UNROLL = quads per loop body.
template <int K, int Q, bool EXACT, int MINB, int UNROLL>
__global__ void __launch_bounds__(288, MINB)
lowk_exact_kernel(const float* __restrict__ queries, const int m, const float* __restrict__ blocks,
                  const int nblocks, const int blocks_per_split, const int index_base, const int stages,
                  u64* __restrict__ keys, const int* __restrict__ enable)
{
    using Ring = TileRing<K>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (enable && *enable == 0) return;        // fallback launch that turned out not to be needed
    const int W = (int)(blockDim.x >> 5) - 1;  // consumer warps; warp W is the producer
    const int warp = (int)(threadIdx.x >> 5);
    const int lane = (int)(threadIdx.x & 31);
    const int b0 = (int)blockIdx.y * blocks_per_split;
    const int nb = min(nblocks, b0 + blocks_per_split) - b0;
    if (nb <= 0) return;  // uniform over the CTA
    Ring ring;
    ring.init(smem_raw, stages, W);
    if (warp == W) {
        if (lane == 0) ring.produce(blocks + (size_t)b0 * Ring::BLOCK_FLOATS, nb);
        return;
    }

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

    const int ntiles = (nb + Ring::TB - 1) / Ring::TB;
    int s = 0, round = 0;
    for (int tile = 0; tile < ntiles; ++tile) {
        mbar_wait(ring.full0 + 8u * s, (uint32_t)(round & 1));
        const float* tsm = ring.tiles + (size_t)s * Ring::TILE_FLOATS;
        const int tb = min(Ring::TB, nb - tile * Ring::TB);
        int j0 = index_base + (b0 + tile * Ring::TB) * LB;
        for (int b = 0; b < tb; ++b) {
            const float* blk = tsm + b * Ring::BLOCK_FLOATS;
#pragma unroll UNROLL
            for (int g = 0; g < LB / 4; ++g) {
                u64 a01[Q], a23[Q];
                lowk_quad_dist<K, Q, EXACT>(blk + 4 * g, qq, a01, a23);
                lowk_quad_argmin<Q>(a01, a23, best, bidx, j0 + 4 * g);
            }
            j0 += LB;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(ring.empty0 + 8u * s);
        if (++s == stages) { s = 0; ++round; }
    }

#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const int q = qbase + i * nct;
        if (q < m && best[i] < 0x7f800000u) atomicMin(keys + q, ((u64)best[i] << 32) | (u64)(unsigned)bidx[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// filter kernel
// ---------------------------------------------------------------------------------------------
// Threshold of the screen for one query.  With u = 2^-24, D the real squared distance, d the
// V0-form FP32 distance, qn = fl(|q|^2), rn_j = fl(|r_j|^2) (row k of the index) and
// s_j = fl(rn_j + sum_t (-2 q_t) r_jt) (k FMAs, ascending t):
//     d >= D (1 - (k+2) u)                         (k subtractions, k FMAs)
//     |s_j - (|r_j|^2 - 2 q.r_j)| <= 2.02 k u (|q| + |r_j|)^2
//     |q|^2 >= qn (1 - (k+1) u)
// so  d < best  implies  s_j <= best + E - qn  for any  E >= (4.02 k + 4) u (|q| + Rmax)^2 plus
// three roundings of the threshold arithmetic itself.  E = 6 (k+4) u (sqrt(qn) + sqrt(R2max))^2
// covers that with margin.  The bound needs every intermediate to stay finite: if qn or R2max
// exceed 1e18 (or are NaN) the threshold is NaN, which every pair passes.
template <int K>
__device__ __forceinline__ float lowk_filter_slack(float qn, float r2max)
{
    if (!(qn <= 1e18f) || !(r2max <= 1e18f)) return nan_f();
    const float a = sqrtf(qn) + sqrtf(r2max);
    const float E = (6.0f * (K + 4) * 5.9604645e-8f) * a * a;
    return E - qn;  // tau = best + (E - qn)
}

// Exact V0-form distances of one query to the 4 references of quad `g` (slow path; scalar ops,
// bit-identical to the packed arithmetic of lowk_quad_dist<.., false>).  For small K the query
// coordinates are recovered from the registers holding -2q (exact: a power-of-two scaling) so the
// slow path never waits on global memory; `from_regs` is false when magnitudes are outside the
// screened range (slack is NaN), where -2q may have overflowed, and for large K.
template <int K>
__device__ __forceinline__ void lowk_exact4(const float* __restrict__ g, const float* __restrict__ qp,
                                            const u64 (&mq)[K], const bool from_regs, float (&d)[4])
{
    if (K <= 4 && from_regs) {
#pragma unroll
        for (int t = 0; t < K; ++t) {
            const float4 r = *reinterpret_cast<const float4*>(g + t * LB);
            float lo, hi;
            upk2(mq[t], lo, hi);
            const float q = -0.5f * lo;
            const float e0 = __fsub_rn(r.x, q), e1 = __fsub_rn(r.y, q), e2 = __fsub_rn(r.z, q), e3 = __fsub_rn(r.w, q);
            if (t == 0) {
                d[0] = __fmul_rn(e0, e0); d[1] = __fmul_rn(e1, e1); d[2] = __fmul_rn(e2, e2); d[3] = __fmul_rn(e3, e3);
            } else {
                d[0] = __fmaf_rn(e0, e0, d[0]); d[1] = __fmaf_rn(e1, e1, d[1]);
                d[2] = __fmaf_rn(e2, e2, d[2]); d[3] = __fmaf_rn(e3, e3, d[3]);
            }
        }
        return;
    }
#pragma unroll 1
    for (int t = 0; t < K; ++t) {
        const float4 r = *reinterpret_cast<const float4*>(g + t * LB);
        const float q = __ldg(qp + t);
        const float e0 = __fsub_rn(r.x, q), e1 = __fsub_rn(r.y, q), e2 = __fsub_rn(r.z, q), e3 = __fsub_rn(r.w, q);
        if (t == 0) {
            d[0] = __fmul_rn(e0, e0); d[1] = __fmul_rn(e1, e1); d[2] = __fmul_rn(e2, e2); d[3] = __fmul_rn(e3, e3);
        } else {
            d[0] = __fmaf_rn(e0, e0, d[0]); d[1] = __fmaf_rn(e1, e1, d[1]);
            d[2] = __fmaf_rn(e2, e2, d[2]); d[3] = __fmaf_rn(e3, e3, d[3]);
        }
    }
}

// G = quads (of 4 references) screened per check.
template <int K, int Q, int MINB, int UNROLL, int G>
__global__ void __launch_bounds__(288, MINB)
lowk_filter_kernel(const float* __restrict__ queries, const int m, const float* __restrict__ header,
                   const float* __restrict__ blocks, const int nblocks, const int blocks_per_split,
                   const int index_base, const int stages, u64* __restrict__ keys, const int* __restrict__ enable)
{
    using Ring = TileRing<K>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (enable && *enable == 0) return;  // fallback launch that turned out not to be needed
    const int W = (int)(blockDim.x >> 5) - 1;
    const int warp = (int)(threadIdx.x >> 5);
    const int lane = (int)(threadIdx.x & 31);
    const int b0 = (int)blockIdx.y * blocks_per_split;
    const int nb = min(nblocks, b0 + blocks_per_split) - b0;
    if (nb <= 0) return;
    Ring ring;
    ring.init(smem_raw, stages, W);
    if (warp == W) {
        if (lane == 0) ring.produce(blocks + (size_t)b0 * Ring::BLOCK_FLOATS, nb);
        return;
    }

    const int nct = W * 32;
    const int qbase = (int)blockIdx.x * (nct * Q) + (int)threadIdx.x;
    const float r2max = __ldg(header);
    u64 mqq[Q][K];   // (-2 q_t, -2 q_t)
    float slack[Q];  // E - |q|^2
    float tau[Q];    // best + slack: pairs with s <= tau (or unordered) go to the exact path
    float best[Q];
    int bidx[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const int q = qbase + i * nct;
        best[i] = inf_f();
        bidx[i] = 0;
        float qn = 0.0f;
#pragma unroll
        for (int t = 0; t < K; ++t) {
            const float v = (q < m) ? __ldg(queries + (size_t)q * K + t) : 0.0f;
            qn = __fmaf_rn(v, v, qn);
            const float w = -2.0f * v;
            mqq[i][t] = pk2(w, w);
        }
        slack[i] = lowk_filter_slack<K>(qn, r2max);
        // lanes past m: tau = -INF passes nothing but NaN padding
        tau[i] = (q < m) ? best[i] + slack[i] : -inf_f();
    }

    const int ntiles = (nb + Ring::TB - 1) / Ring::TB;
    int s = 0, round = 0;
    for (int tile = 0; tile < ntiles; ++tile) {
        mbar_wait(ring.full0 + 8u * s, (uint32_t)(round & 1));
        const float* tsm = ring.tiles + (size_t)s * Ring::TILE_FLOATS;
        const int tb = min(Ring::TB, nb - tile * Ring::TB);
        int j0 = index_base + (b0 + tile * Ring::TB) * LB;
        for (int b = 0; b < tb; ++b) {
            const float* blk = tsm + b * Ring::BLOCK_FLOATS;
#pragma unroll UNROLL
            for (int g = 0; g < LB / (4 * G); ++g) {
                const float* gp = blk + 4 * G * g;
                // screen: s = |r|^2 - 2 q.r for G quads x Q queries; only the minimum per query is
                // compared (a min tree keeps the dependency chains short; fmin drops NaN operands,
                // which is safe: NaN arises only from NaN references, whose distance never wins,
                // or beyond the magnitude guard, where tau is NaN and everything passes)
                // all 2*G*Q accumulator chains advance together, one dimension at a time, so that
                // dependent FFMA2s are 2*G*Q instructions apart (no fixed-latency stalls)
                ulonglong2 rn[G];
                u64 s01[G][Q], s23[G][Q];
#pragma unroll
                for (int h = 0; h < G; ++h) rn[h] = lds_v2u64(gp + 4 * h + K * LB);
#pragma unroll
                for (int t = 0; t < K; ++t) {
#pragma unroll
                    for (int h = 0; h < G; ++h) {
                        const ulonglong2 rv = lds_v2u64(gp + 4 * h + t * LB);
#pragma unroll
                        for (int i = 0; i < Q; ++i) {
                            s01[h][i] = fma2(rv.x, mqq[i][t], t == 0 ? rn[h].x : s01[h][i]);
                            s23[h][i] = fma2(rv.y, mqq[i][t], t == 0 ? rn[h].y : s23[h][i]);
                        }
                    }
                }
                float mn[Q];
#pragma unroll
                for (int h = 0; h < G; ++h) {
#pragma unroll
                    for (int i = 0; i < Q; ++i) {
                        float x0, x1, x2, x3;
                        upk2(s01[h][i], x0, x1);
                        upk2(s23[h][i], x2, x3);
                        mn[i] = (h == 0) ? fminf(min3(x0, x1, x2), x3) : min3(mn[i], min3(x0, x1, x2), x3);
                    }
                }
                bool any = false;
#pragma unroll
                for (int i = 0; i < Q; ++i) any |= !(mn[i] > tau[i]);
                if (any) {
#pragma unroll
                    for (int i = 0; i < Q; ++i) {
                        const int q = qbase + i * nct;
                        if (q < m && !(mn[i] > tau[i])) {
#pragma unroll 1
                            for (int h = 0; h < G; ++h) {
                                float d[4];
                                lowk_exact4<K>(gp + 4 * h, queries + (size_t)q * K, mqq[i], slack[i] == slack[i], d);
                                const int jq = j0 + 4 * (G * g + h);
                                if (d[0] < best[i]) { best[i] = d[0]; bidx[i] = jq; }
                                if (d[1] < best[i]) { best[i] = d[1]; bidx[i] = jq + 1; }
                                if (d[2] < best[i]) { best[i] = d[2]; bidx[i] = jq + 2; }
                                if (d[3] < best[i]) { best[i] = d[3]; bidx[i] = jq + 3; }
                            }
                            tau[i] = best[i] + slack[i];
                        }
                    }
                }
            }
            j0 += LB;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(ring.empty0 + 8u * s);
        if (++s == stages) { s = 0; ++round; }
    }

#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const int q = qbase + i * nct;
        if (q < m && best[i] < inf_f()) atomicMin(keys + q, pack_key(best[i], bidx[i]));
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
enum LowkMode { LOWK_FILTER = 0, LOWK_EXACT_FMA = 1, LOWK_EXACT_V0 = 2 };

template <typename Kern>
cudaError_t lowk_launch_kernel(Kern kern, int K, const LowkArgs& a, int* occupancy_out, bool with_header)
{
    const size_t smem = (size_t)LOWK_BAR_BYTES + (size_t)a.stages * lowk_tile_bytes(K);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int threads = (a.warps + 1) * 32;
    if (occupancy_out) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occupancy_out, kern, threads, smem);
    dim3 grid((unsigned)a.nqb, (unsigned)a.splits);
    void* args_h[] = {(void*)&a.queries, (void*)&a.m, (void*)&a.header, (void*)&a.blocks, (void*)&a.nblocks,
                      (void*)&a.blocks_per_split, (void*)&a.index_base, (void*)&a.stages, (void*)&a.keys, (void*)&a.enable};
    void* args_n[] = {(void*)&a.queries, (void*)&a.m, (void*)&a.blocks, (void*)&a.nblocks,
                      (void*)&a.blocks_per_split, (void*)&a.index_base, (void*)&a.stages, (void*)&a.keys, (void*)&a.enable};
    return cudaLaunchKernel((const void*)kern, grid, dim3(threads), with_header ? args_h : args_n, smem, a.stream);
}

template <int K, int Q>
cudaError_t lowk_launch_t(int mode, const LowkArgs& a, int* occ)
{
    constexpr int MINB = lowk_minb(K, Q), UNR = lowk_unroll(K);
    if (mode == LOWK_FILTER) return lowk_launch_kernel(lowk_filter_kernel<K, Q, MINB, UNR, lowk_screen_quads(K)>, K, a, occ, true);
    if (mode == LOWK_EXACT_FMA) return lowk_launch_kernel(lowk_exact_kernel<K, Q, false, MINB, UNR>, K, a, occ, false);
    return lowk_launch_kernel(lowk_exact_kernel<K, Q, true, MINB, UNR>, K, a, occ, false);
}

// dispatch over (k, q, mode); q must be lowk_q_default(k) or lowk_q_alt(k) (alt: not for V0 rounding).
// occupancy_out != NULL: do not launch, report resident CTAs per SM for that configuration.
cudaError_t lowk_launch(int k, int q, int mode, const LowkArgs& a, int* occupancy_out);

}  // namespace nns
