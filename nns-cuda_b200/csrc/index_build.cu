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

// `dst` lists the destinations of the built blocks: the index of this GPU and, in the single-process
// multi-GPU ingest (capi.cu, search_multi), the same slice of every peer GPU's index -- the kernel
// stores each row to all of them, so the NVLink all-gather of the built index is fused into the build
// (posted peer stores; nothing re-reads the slice).
__global__ void __launch_bounds__(IB_THREADS)
index_build_kernel(const float* __restrict__ aos, const int n, const int k, unsigned* __restrict__ hmax,
                   const BlockDsts dst)
{
    __shared__ float tile[LB][IB_KC + 1];
    const long long b = blockIdx.x;
    const long long j0 = b * LB;
    const long long block_off = b * (long long)(k + 1) * LB;
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
        for (int e = threadIdx.x; e < LB * kc; e += IB_THREADS) {
            const int t = e / LB, r = e - t * LB;
            const float v = tile[r][t];
            const long long o = block_off + (long long)(c0 + t) * LB + r;
            for (int d = 0; d < dst.count; ++d) dst.p[d][o] = v;
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
        for (int d = 0; d < dst.count; ++d) dst.p[d][block_off + (long long)k * LB + threadIdx.x] = rn;  // NaN for padded lanes
        unsigned bits = (rn == rn) ? __float_as_uint(rn) : 0u;
        bits = __reduce_max_sync(0xffffffffu, bits);
        if ((threadIdx.x & 31) == 0 && bits != 0u) atomicMax(hmax, bits);
    }
}

// Multi-GPU ingest: every GPU accumulates the maxima / flags of ITS slice in its own slot of the two
// headers (index header word HDR_PART_MAX + g; tensor section words THDR_PART_MAX + g, THDR_PART_FLAGS + g),
// publishes the slot to every peer with plain peer stores, and after the cross-GPU event wait each GPU
// folds the slots into the words the search kernels read.
__global__ void header_publish_kernel(const HeaderPeers hp, const int g)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned a = reinterpret_cast<const unsigned*>(hp.header[hp.self])[HDR_PART_MAX + g];
    unsigned b = 0, f = 0;
    if (hp.section[hp.self]) {
        b = reinterpret_cast<const unsigned*>(hp.section[hp.self])[THDR_PART_MAX + g];
        f = reinterpret_cast<const unsigned*>(hp.section[hp.self])[THDR_PART_FLAGS + g];
    }
    for (int d = 0; d < hp.count; ++d) {
        if (d == hp.self) continue;
        reinterpret_cast<unsigned*>(hp.header[d])[HDR_PART_MAX + g] = a;
        if (hp.section[d]) {
            reinterpret_cast<unsigned*>(hp.section[d])[THDR_PART_MAX + g] = b;
            reinterpret_cast<unsigned*>(hp.section[d])[THDR_PART_FLAGS + g] = f;
        }
    }
}

__global__ void header_fold_kernel(float* __restrict__ header, float* __restrict__ section, const int parts)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned a = 0, b = 0, f = 0;
    for (int g = 0; g < parts; ++g) {
        a = max(a, reinterpret_cast<const unsigned*>(header)[HDR_PART_MAX + g]);
        if (section) {
            b = max(b, reinterpret_cast<const unsigned*>(section)[THDR_PART_MAX + g]);
            f |= reinterpret_cast<const unsigned*>(section)[THDR_PART_FLAGS + g];
        }
    }
    reinterpret_cast<unsigned*>(header)[0] = a;
    if (section) {
        reinterpret_cast<unsigned*>(section)[THDR_MAX] = b;
        reinterpret_cast<unsigned*>(section)[THDR_FLAGS] = f;
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

// dst[i] = min(dst[i], src[i]); dst may live in a peer GPU's memory and is updated by several GPUs at
// once, so the atomic is SYSTEM scope (a device-scope atomic is only atomic against threads of the
// issuing GPU): red.global.min.u64 over NVLink, the cross-GPU (dist, idx) reduction of the
// reference-sharded search.
__global__ void keys_merge_kernel(u64* __restrict__ dst, const u64* __restrict__ src, const int m)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        const u64 key = src[i];
        if (key < KEY_INIT) atomicMin_system(dst + i, key);
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
    BlockDsts dst{};
    dst.p[0] = d_blocks;
    dst.count = 1;
    return launch_index_build_to(k, n, d_refs_aos, d_header, 0, dst, reset_header, st);
}

cudaError_t launch_index_build_to(int k, int n, const float* d_refs_aos, float* d_header, int hmax_word,
                                  const BlockDsts& dst, bool reset_header, cudaStream_t st, int write_blocks)
{
    if (reset_header) {
        cudaError_t e = cudaMemsetAsync(d_header, 0, INDEX_HEADER_FLOATS * sizeof(float), st);
        if (e != cudaSuccess) return e;
    }
    // write_blocks > ceil(n / 128): the extra blocks are all padding (NaN), e.g. the equal-sized last slice of an all-gather
    const int nblocks = write_blocks > 0 ? write_blocks : (n + LB - 1) / LB;
    if (nblocks == 0) return cudaSuccess;
    index_build_kernel<<<nblocks, IB_THREADS, 0, st>>>(d_refs_aos, n, k, reinterpret_cast<unsigned*>(d_header) + hmax_word, dst);
    return cudaGetLastError();
}

cudaError_t launch_header_publish(const HeaderPeers& hp, int g, cudaStream_t st)
{
    header_publish_kernel<<<1, 32, 0, st>>>(hp, g);
    return cudaGetLastError();
}

cudaError_t launch_header_fold(float* d_header, float* d_section, int parts, cudaStream_t st)
{
    header_fold_kernel<<<1, 32, 0, st>>>(d_header, d_section, parts);
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
