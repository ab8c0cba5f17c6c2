// index_build.cu -- AoS -> tiled-SoA reference index, and the packed-key helpers.
//
// index layout (nns_plan.h): float[32] header, then float[nblocks][k+1][128]; reference j lives in
// block j/128, lane j%128; rows 0..k-1 are its coordinates, row k is |r_j|^2 (FP32, ascending-t
// FMA chain); the tail of the last block is NaN so that padded points can never win a comparison.
// header[0] accumulates max_j |r_j|^2 (atomicMax on the bit pattern; NaN norms are skipped).
// Replaces the reference's naive transpose v4::mat_inv_kernel (core.cu:293-306): here both the
// global read (a contiguous 128*k-float AoS chunk per block) and the global write are coalesced,
// staged through a padded shared-memory tile.
#include "nns_internal.h"

namespace nns {

constexpr int IB_KC = 32;       // dimensions staged per pass
constexpr int IB_THREADS = 256;

__global__ void __launch_bounds__(IB_THREADS)
index_build_kernel(const float* __restrict__ aos, const int n, const int k, float* __restrict__ header,
                   float* __restrict__ blocks)
{
    __shared__ float tile[LB][IB_KC + 1];
    const long long b = blockIdx.x;
    const long long j0 = b * LB;
    float* out_block = blocks + b * (long long)(k + 1) * LB;
    float rn = 0.0f;
    for (int c0 = 0; c0 < k; c0 += IB_KC) {
        const int kc = min(IB_KC, k - c0);
        // coalesced read: consecutive threads walk the AoS rows of this block
        for (int e = threadIdx.x; e < LB * kc; e += IB_THREADS) {
            const int r = e / kc, t = e - r * kc;
            const long long j = j0 + r;
            tile[r][t] = (j < n) ? __ldg(aos + j * k + c0 + t) : nan_f();
        }
        __syncthreads();
        // coalesced write: consecutive threads walk one dimension row of the block
        float* out = out_block + (long long)c0 * LB;
        for (int e = threadIdx.x; e < LB * kc; e += IB_THREADS) {
            const int t = e / LB, r = e - t * LB;
            out[(long long)t * LB + r] = tile[r][t];
        }
        if (threadIdx.x < LB) {
            for (int t = 0; t < kc; ++t) {
                const float v = tile[threadIdx.x][t];
                rn = __fmaf_rn(v, v, rn);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < LB) {
        out_block[(long long)k * LB + threadIdx.x] = rn;  // NaN for padded lanes
        unsigned bits = (rn == rn) ? __float_as_uint(rn) : 0u;
        bits = __reduce_max_sync(0xffffffffu, bits);
        if ((threadIdx.x & 31) == 0 && bits != 0u) atomicMax(reinterpret_cast<unsigned*>(header), bits);
    }
}

__global__ void keys_init_kernel(u64* __restrict__ keys, const int m)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) keys[i] = KEY_INIT;
}

__global__ void keys_unpack_kernel(const u64* __restrict__ keys, const int m, int* __restrict__ idx,
                                   float* __restrict__ dist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        const u64 key = keys[i];
        idx[i] = (int)(unsigned)(key & 0xffffffffull);
        if (dist) dist[i] = __uint_as_float((unsigned)(key >> 32));
    }
}

// dst[i] = min(dst[i], src[i]); dst may live in a peer GPU's memory (NVLink P2P atomics)
__global__ void keys_merge_kernel(u64* __restrict__ dst, const u64* __restrict__ src, const int m)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        const u64 key = src[i];
        if (key < KEY_INIT) atomicMin(dst + i, key);
    }
}

cudaError_t launch_keys_merge(u64* d_dst, const u64* d_src, int m, cudaStream_t st)
{
    if (m == 0) return cudaSuccess;
    keys_merge_kernel<<<(m + 255) / 256, 256, 0, st>>>(d_dst, d_src, m);
    return cudaGetLastError();
}

cudaError_t launch_index_build(int k, int n, const float* d_refs_aos, float* d_header, float* d_blocks,
                               bool reset_header, cudaStream_t st)
{
    if (reset_header) {
        cudaError_t e = cudaMemsetAsync(d_header, 0, INDEX_HEADER_FLOATS * sizeof(float), st);
        if (e != cudaSuccess) return e;
    }
    const int nblocks = (n + LB - 1) / LB;
    if (nblocks == 0) return cudaSuccess;
    index_build_kernel<<<nblocks, IB_THREADS, 0, st>>>(d_refs_aos, n, k, d_header, d_blocks);
    return cudaGetLastError();
}

cudaError_t launch_keys_init(u64* d_keys, int m, cudaStream_t st)
{
    if (m == 0) return cudaSuccess;
    keys_init_kernel<<<(m + 255) / 256, 256, 0, st>>>(d_keys, m);
    return cudaGetLastError();
}

cudaError_t launch_keys_unpack(const u64* d_keys, int m, int* d_idx, float* d_dist, cudaStream_t st)
{
    if (m == 0) return cudaSuccess;
    keys_unpack_kernel<<<(m + 255) / 256, 256, 0, st>>>(d_keys, m, d_idx, d_dist);
    return cudaGetLastError();
}

}  // namespace nns
