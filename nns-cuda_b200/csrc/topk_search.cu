// topk_search.cu -- K nearest neighbours (K <= 32), an extension behind the extended ABI: the
// reference returns one index per query (core.cu:52).  Same distance semantics as the 1-NN kernels:
// V0's subtract-square-accumulate form in FP32 over ascending dimensions (core.cu:38-43; FMA by
// default, separately rounded mul/add with NNS_B200_FLAG_V0_ROUNDING), neighbours ordered by the
// packed (dist, idx) key, i.e. by distance and then by the lower index -- for K = 1 this is V0's
// first-minimum rule.  References whose distance is NaN or +INF are never reported; missing
// neighbours (n < K) are KEY_INIT (index -1, distance +INF after unpacking).
//
// Mapping.  A warp owns TOPK_WQ queries (coordinates broadcast from shared memory) and streams the
// reference blocks of its CTA's split; each lane evaluates four consecutive references of a block
// (one 16-byte load per dimension from the tiled-SoA index) against the warp's queries.  The K best
// keys of a query live in REGISTERS, one per lane (lane j = entry j, unsorted); tau = the worst kept
// distance is warp-uniform, so the per-pair cost on top of the distance is one compare, and the rare
// pair that passes (K ln(n/K) per query) is inserted with a warp max-reduction.  Splits of one query
// write their lists to scratch and topk_merge_kernel -- which also folds in the keys the caller
// already holds, so shards / GPUs accumulate like the 1-NN keys do -- selects and sorts the K best.
#include "nns_internal.h"

namespace nns {

constexpr int TOPK_WQ = 4;        // queries per warp
constexpr int TOPK_WARPS = 8;     // warps per CTA, each with its own queries
constexpr int TOPK_THREADS = 32 * TOPK_WARPS;
constexpr int TOPK_CTA_Q = TOPK_WQ * TOPK_WARPS;

__device__ __forceinline__ u64 warp_max_u64(u64 v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const u64 o = __shfl_xor_sync(0xffffffffu, v, off);
        v = o > v ? o : v;
    }
    return v;
}

// insert `key` into the list held one entry per lane (in_list = lane < K; the other lanes hold 0 and are never the
// maximum); returns the new worst key
__device__ __forceinline__ u64 topk_insert(u64& mine, const u64 key, const u64 worst, const int lane, const bool in_list)
{
    if (key < worst) {
        // a key that is already in the list (the same reference offered twice: overlapping shards, or the FP32 pass
        // that finishes a tensor-screened search whose first batches were merged already) is not inserted again
        if (__any_sync(0xffffffffu, in_list && mine == key)) return worst;
        const unsigned holders = __ballot_sync(0xffffffffu, mine == worst);
        if (lane == __ffs(holders) - 1) mine = key;
        return warp_max_u64(mine);
    }
    return worst;
}

template <bool EXACT>
__global__ void __launch_bounds__(TOPK_THREADS)
topk_search_kernel(const float* __restrict__ queries, const int m, const int k, const int K,
                   const float* __restrict__ blocks, const int nblocks, const int blocks_per_split,
                   const int index_base, u64* __restrict__ lists /* [splits][m][K] */,
                   const int stride, const int only_sampled, const int* __restrict__ enable)
{
    // stride > 1: the reference blocks are partitioned into the SAMPLED ones (b % stride == 0) and the rest;
    // only_sampled = 1 scans the sample, 0 the rest (the tensor-screened K-nearest search: tensor_topk_search)
    extern __shared__ __align__(16) float qs[];  // [TOPK_WARPS][k][TOPK_WQ]
    if (enable != nullptr && *enable == 0) return;  // conditional fallback launch
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const int q0 = (int)blockIdx.x * TOPK_CTA_Q + warp * TOPK_WQ;
    float* wq = qs + (size_t)warp * k * TOPK_WQ;
    for (int e = lane; e < k * TOPK_WQ; e += 32) {
        const int t = e / TOPK_WQ, i = e - t * TOPK_WQ;
        wq[e] = (q0 + i < m) ? __ldg(queries + (size_t)(q0 + i) * k + t) : nan_f();
    }
    __syncwarp();
    if (q0 >= m) return;

    u64 mine[TOPK_WQ], worst[TOPK_WQ];
    float tau[TOPK_WQ];
#pragma unroll
    for (int i = 0; i < TOPK_WQ; ++i) {
        mine[i] = lane < K ? KEY_INIT : 0ull;
        worst[i] = KEY_INIT;
        tau[i] = inf_f();
    }
    const int b0 = (int)blockIdx.y * blocks_per_split;
    const int b1 = min(nblocks, b0 + blocks_per_split);
    for (int b = b0; b < b1; ++b) {
        if (stride > 1 && ((b % stride == 0) != (only_sampled != 0))) continue;
        const float4* blk = reinterpret_cast<const float4*>(blocks + (size_t)b * (k + 1) * LB) + lane;
        float acc[4][TOPK_WQ];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int i = 0; i < TOPK_WQ; ++i) acc[e][i] = 0.0f;
#pragma unroll 4
        for (int t = 0; t < k; ++t) {
            const float4 r4 = __ldg(blk + (size_t)t * (LB / 4));
            const float4 q4 = *reinterpret_cast<const float4*>(wq + t * TOPK_WQ);
            const float r[4] = {r4.x, r4.y, r4.z, r4.w};
            const float q[TOPK_WQ] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int i = 0; i < TOPK_WQ; ++i) {
                    const float d = q[i] - r[e];
                    acc[e][i] = EXACT ? __fadd_rn(acc[e][i], __fmul_rn(d, d)) : __fmaf_rn(d, d, acc[e][i]);
                }
        }
#pragma unroll
        for (int i = 0; i < TOPK_WQ; ++i) {
            // distances <= tau (the worst kept one; equal distances may still win on the index)
            unsigned hit = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) hit |= (acc[e][i] <= tau[i]) ? (1u << e) : 0u;
            unsigned lanes = __ballot_sync(0xffffffffu, hit != 0u);
            while (lanes) {  // warp-uniform loop over the lanes that hold a passing pair
                const int src = __ffs(lanes) - 1;
                lanes &= lanes - 1;
                const unsigned h = __shfl_sync(0xffffffffu, hit, src);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float d = __shfl_sync(0xffffffffu, acc[e][i], src);
                    if ((h >> e) & 1u) {  // uniform
                        const u64 key = (d < inf_f()) ? pack_key(d, index_base + b * LB + 4 * src + e) : KEY_INIT;
                        worst[i] = topk_insert(mine[i], key, worst[i], lane, lane < K);
                    }
                }
                tau[i] = __uint_as_float((unsigned)(worst[i] >> 32));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < TOPK_WQ; ++i)
        if (q0 + i < m && lane < K) lists[((size_t)blockIdx.y * m + (q0 + i)) * K + lane] = mine[i];
}

// merge of variable-length candidate lists (tensor-screened K-nearest search): exact[q][0 .. count[q]) into keys[q]
__global__ void __launch_bounds__(256)
topk_merge_var_kernel(u64* __restrict__ keys, const u64* __restrict__ exact, const unsigned* __restrict__ count, const int m,
                      const int K, const unsigned cap, const unsigned* __restrict__ skip_if_set)
{
    if (skip_if_set != nullptr && *skip_if_set != 0u) return;  // the screen overflowed: the FP32 pass launched next decides
    const int lane = (int)(threadIdx.x & 31);
    const int q = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (q >= m) return;
    u64 mine = lane < K ? keys[(size_t)q * K + lane] : 0ull;
    u64 worst = warp_max_u64(mine);
    const unsigned cnt = min(count[q], cap);
    for (unsigned i0 = 0; i0 < cnt; i0 += 32) {
        const u64 cand = (i0 + lane < cnt) ? exact[(size_t)q * cap + i0 + lane] : KEY_INIT;
        unsigned live = __ballot_sync(0xffffffffu, cand < worst);
        while (live) {
            const int src = __ffs(live) - 1;
            live &= live - 1;
            const u64 key = __shfl_sync(0xffffffffu, cand, src);
            worst = topk_insert(mine, key, worst, lane, lane < K);
        }
    }
    int rank = 0;
    for (int j = 0; j < K; ++j) {
        const u64 o = __shfl_sync(0xffffffffu, mine, j);
        rank += (o < mine || (o == mine && j < lane)) ? 1 : 0;
    }
    if (lane < K) keys[(size_t)q * K + rank] = mine;
}

cudaError_t topk_merge_var_launch(u64* d_keys, const u64* d_exact, const unsigned* d_count, int m, int K, unsigned cap,
                                  const unsigned* skip_if_set, cudaStream_t st)
{
    if (m == 0) return cudaSuccess;
    topk_merge_var_kernel<<<(unsigned)(((long long)m * 32 + 255) / 256), 256, 0, st>>>(d_keys, d_exact, d_count, m, K, cap, skip_if_set);
    return cudaGetLastError();
}

// one warp per query: the K smallest keys of (keys[q][0..K) as given) U (lists[s][q][0..K) for every split),
// written back to keys[q] in ascending order
__global__ void __launch_bounds__(256)
topk_merge_kernel(u64* __restrict__ keys, const u64* __restrict__ lists, const int m, const int K, const int splits,
                  const int* __restrict__ enable)
{
    if (enable != nullptr && *enable == 0) return;  // the search behind this merge did not run
    const int lane = (int)(threadIdx.x & 31);
    const int q = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (q >= m) return;
    u64 mine = lane < K ? keys[(size_t)q * K + lane] : 0ull;
    u64 worst = warp_max_u64(mine);
    for (int s = 0; s < splits; ++s) {
        const u64 cand = lane < K ? lists[((size_t)s * m + q) * K + lane] : KEY_INIT;
        // candidates that cannot enter are skipped warp-wide with one ballot
        unsigned live = __ballot_sync(0xffffffffu, cand < worst);
        while (live) {
            const int src = __ffs(live) - 1;
            live &= live - 1;
            const u64 key = __shfl_sync(0xffffffffu, cand, src);
            worst = topk_insert(mine, key, worst, lane, lane < K);
        }
    }
    // rank = number of entries that sort before mine (KEY_INIT pads tie: break by lane)
    int rank = 0;
    for (int j = 0; j < K; ++j) {
        const u64 o = __shfl_sync(0xffffffffu, mine, j);
        rank += (o < mine || (o == mine && j < lane)) ? 1 : 0;
    }
    if (lane < K) keys[(size_t)q * K + rank] = mine;
}

__global__ void topk_unpack_kernel(const u64* __restrict__ keys, const long long count, int* __restrict__ idx, float* __restrict__ dist)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        const u64 key = keys[i];
        idx[i] = key == KEY_INIT ? -1 : (int)(unsigned)(key & 0xffffffffull);
        if (dist) dist[i] = __uint_as_float((unsigned)(key >> 32));
    }
}

cudaError_t topk_unpack_launch(const u64* d_keys, int m, int K, int* d_idx, float* d_dist, cudaStream_t st)
{
    const long long count = (long long)m * K;
    if (count == 0) return cudaSuccess;
    topk_unpack_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(d_keys, count, d_idx, d_dist);
    return cudaGetLastError();
}

size_t topk_scratch_bytes(int m, int K, int splits) { return (size_t)splits * (size_t)m * (size_t)K * sizeof(u64); }

int topk_choose_splits(int m, int n, int num_sms)
{
    const int nblocks = (n + LB - 1) / LB;
    const int groups = (m + TOPK_CTA_Q - 1) / TOPK_CTA_Q;
    const long long want = 8LL * num_sms;  // CTAs
    int s = groups >= want ? 1 : (int)((want + groups - 1) / groups);
    if (s > nblocks) s = nblocks > 0 ? nblocks : 1;
    if (s > 1024) s = 1024;
    return s < 1 ? 1 : s;
}

cudaError_t topk_search_launch(int k, int m, int n, int K, const float* d_queries, const float* d_blocks, int index_base,
                               u64* d_keys, u64* d_scratch, int splits, bool exact, cudaStream_t st, int* launches,
                               int stride, int only_sampled, const int* enable)
{
    if (launches) *launches = 0;
    if (m == 0 || n == 0) return cudaSuccess;
    const int nblocks = (n + LB - 1) / LB;
    const int bps = (nblocks + splits - 1) / splits;
    splits = (nblocks + bps - 1) / bps;
    const size_t smem = (size_t)TOPK_WARPS * k * TOPK_WQ * sizeof(float);
    dim3 grid((unsigned)((m + TOPK_CTA_Q - 1) / TOPK_CTA_Q), (unsigned)splits);
    cudaError_t e;
    if (exact) {
        e = cudaFuncSetAttribute(topk_search_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        topk_search_kernel<true><<<grid, TOPK_THREADS, smem, st>>>(d_queries, m, k, K, d_blocks, nblocks, bps, index_base, d_scratch,
                                                                   stride, only_sampled, enable);
    } else {
        e = cudaFuncSetAttribute(topk_search_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        topk_search_kernel<false><<<grid, TOPK_THREADS, smem, st>>>(d_queries, m, k, K, d_blocks, nblocks, bps, index_base, d_scratch,
                                                                    stride, only_sampled, enable);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    topk_merge_kernel<<<(unsigned)(((long long)m * 32 + 255) / 256), 256, 0, st>>>(d_keys, d_scratch, m, K, splits, enable);
    if (launches) *launches = 2;
    return cudaGetLastError();
}

}  // namespace nns
