// wide_search.cu -- reference-parallel fused distance + argmin for the shapes the register-
// blocked kernel does not cover: any k (k > 32 until the tensor path takes over), and very few
// queries (the reference's m = 1 benchmark shapes, main.cu:39-42, which are HBM-bound).
//
// A CTA takes WIDE_QT queries (coordinates broadcast from shared memory) and a range of
// reference blocks; each WARP owns one 128-point block at a time and each lane four consecutive
// points of it, read with one 16-byte load per dimension (a warp reads a whole 512-byte row of the
// tiled-SoA index per instruction: the m = 1 shapes are HBM-bound and need the bytes in flight).
// A thread keeps a running (dist, idx) per query over its points in ascending index order (strict
// '<' => first minimum, as V0 core.cu:44), and the CTA merges through the packed key: warp shuffle
// min -> shared memory -> one atomicMin per query.  Same result semantics and the same FP32 operation
// order (ascending t, FMA or V0 rounding) as the low-k kernel.
// Replaces the structure of v7::cudaCallKernel + host merge (core.cu:589-633, 669-696).
#include "nns_internal.h"

namespace nns {

static_assert(WIDE_QT == 4, "the query tile is read as float4");

template <bool EXACT>
__global__ void __launch_bounds__(WIDE_THREADS)
wide_search_kernel(const float* __restrict__ queries, const int m, const int k,
                   const float* __restrict__ blocks, const int nblocks, const int blocks_per_split,
                   const int index_base, u64* __restrict__ keys, const int* __restrict__ enable)
{
    if (enable != nullptr && *enable == 0) return;  // conditional fallback launch (tensor path overflow)
    extern __shared__ __align__(16) float qs[];  // [k][WIDE_QT]
    __shared__ u64 red[WIDE_THREADS / 32][WIDE_QT];
    const int q0 = (int)blockIdx.x * WIDE_QT;
    for (int e = threadIdx.x; e < k * WIDE_QT; e += WIDE_THREADS) {
        const int t = e / WIDE_QT, i = e - t * WIDE_QT;
        const int q = q0 + i;
        qs[e] = (q < m) ? __ldg(queries + (size_t)q * k + t) : nan_f();
    }
    __syncthreads();

    const int b0 = (int)blockIdx.y * blocks_per_split;
    const int b1 = min(nblocks, b0 + blocks_per_split);
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);

    float best[WIDE_QT];
    int bidx[WIDE_QT];
#pragma unroll
    for (int i = 0; i < WIDE_QT; ++i) { best[i] = inf_f(); bidx[i] = 0; }

    for (int b = b0 + warp; b < b1; b += WIDE_THREADS / 32) {
        // rows 0..k-1 of block b (row k = |r|^2 is not used here); lane = points 4*lane .. 4*lane+3
        const float4* blk = reinterpret_cast<const float4*>(blocks + (size_t)b * (k + 1) * LB) + lane;
        float acc[4][WIDE_QT];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int i = 0; i < WIDE_QT; ++i) acc[e][i] = 0.0f;
#pragma unroll 4
        for (int t = 0; t < k; ++t) {
            const float4 r4 = __ldg(blk + (size_t)t * (LB / 4));
            const float4 q4 = *reinterpret_cast<const float4*>(qs + t * WIDE_QT);
            const float r[4] = {r4.x, r4.y, r4.z, r4.w};
            const float q[WIDE_QT] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int i = 0; i < WIDE_QT; ++i) {
                    const float d = q[i] - r[e];
                    acc[e][i] = EXACT ? __fadd_rn(acc[e][i], __fmul_rn(d, d)) : __fmaf_rn(d, d, acc[e][i]);
                }
        }
        const int j = index_base + b * LB + 4 * lane;
#pragma unroll
        for (int e = 0; e < 4; ++e)  // ascending index: the first minimum wins
#pragma unroll
            for (int i = 0; i < WIDE_QT; ++i) {
                if (acc[e][i] < best[i]) { best[i] = acc[e][i]; bidx[i] = j + e; }
            }
    }

#pragma unroll
    for (int i = 0; i < WIDE_QT; ++i) {
        u64 key = pack_key(best[i], bidx[i]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const u64 o = __shfl_xor_sync(0xffffffffu, key, off);
            key = o < key ? o : key;
        }
        if (lane == 0) red[warp][i] = key;
    }
    __syncthreads();
    if (threadIdx.x < WIDE_QT) {
        u64 key = red[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < WIDE_THREADS / 32; ++w) {
            const u64 o = red[w][threadIdx.x];
            key = o < key ? o : key;
        }
        const int q = q0 + (int)threadIdx.x;
        if (q < m && key < KEY_INIT) atomicMin(keys + q, key);
    }
}

cudaError_t wide_launch(bool exact, const WideArgs& a)
{
    const size_t smem = (size_t)a.k * WIDE_QT * sizeof(float);
    dim3 grid((unsigned)a.nqg, (unsigned)a.splits);
    cudaError_t e;
    if (exact) {
        e = cudaFuncSetAttribute(wide_search_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        wide_search_kernel<true><<<grid, WIDE_THREADS, smem, a.stream>>>(
            a.queries, a.m, a.k, a.blocks, a.nblocks, a.blocks_per_split, a.index_base, a.keys, a.enable);
    } else {
        e = cudaFuncSetAttribute(wide_search_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        wide_search_kernel<false><<<grid, WIDE_THREADS, smem, a.stream>>>(
            a.queries, a.m, a.k, a.blocks, a.nblocks, a.blocks_per_split, a.index_base, a.keys, a.enable);
    }
    return cudaGetLastError();
}

}  // namespace nns
