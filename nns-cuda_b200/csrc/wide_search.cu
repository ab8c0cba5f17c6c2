// wide_search.cu -- reference-parallel fused distance + argmin for the shapes the register-
// blocked kernel does not cover: any k (k > 32 until the tensor path takes over), and very few
// queries (the reference's m = 1 benchmark shapes, main.cu:39-42, which are HBM-bound).
//
// A CTA takes WIDE_QT queries (coordinates broadcast from shared memory) and a range of
// reference blocks; each thread owns one lane of a 128-point block, reads its coordinates with
// coalesced loads from the tiled-SoA index, keeps a running (dist, idx) per query (ascending j,
// strict '<' => first minimum, as V0 core.cu:44), and the CTA merges through the packed key:
// warp shuffle min -> shared memory -> one atomicMin per query.  Same result semantics and the
// same FP32 operation order (ascending t, FMA or V0 rounding) as the low-k kernel.
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
    const int half = (int)(threadIdx.x >> 7);  // WIDE_THREADS / LB = 2 blocks in flight
    const int l = (int)(threadIdx.x & (LB - 1));

    float best[WIDE_QT];
    int bidx[WIDE_QT];
#pragma unroll
    for (int i = 0; i < WIDE_QT; ++i) { best[i] = inf_f(); bidx[i] = 0; }

    for (int b = b0 + half; b < b1; b += WIDE_THREADS / LB) {
        const float* blk = blocks + (size_t)b * (k + 1) * LB + l;  // rows 0..k-1 (row k = |r|^2 is not used here)
        float acc[WIDE_QT];
#pragma unroll
        for (int i = 0; i < WIDE_QT; ++i) acc[i] = 0.0f;
#pragma unroll 4
        for (int t = 0; t < k; ++t) {
            const float r = __ldg(blk + (size_t)t * LB);
            const float4 q4 = *reinterpret_cast<const float4*>(qs + t * WIDE_QT);
            const float d0 = q4.x - r, d1 = q4.y - r, d2 = q4.z - r, d3 = q4.w - r;
            if (EXACT) {
                acc[0] = __fadd_rn(acc[0], __fmul_rn(d0, d0));
                acc[1] = __fadd_rn(acc[1], __fmul_rn(d1, d1));
                acc[2] = __fadd_rn(acc[2], __fmul_rn(d2, d2));
                acc[3] = __fadd_rn(acc[3], __fmul_rn(d3, d3));
            } else {
                acc[0] = __fmaf_rn(d0, d0, acc[0]);
                acc[1] = __fmaf_rn(d1, d1, acc[1]);
                acc[2] = __fmaf_rn(d2, d2, acc[2]);
                acc[3] = __fmaf_rn(d3, d3, acc[3]);
            }
        }
        const int j = index_base + b * LB + l;
#pragma unroll
        for (int i = 0; i < WIDE_QT; ++i) {
            if (acc[i] < best[i]) { best[i] = acc[i]; bidx[i] = j; }
        }
    }

    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
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
