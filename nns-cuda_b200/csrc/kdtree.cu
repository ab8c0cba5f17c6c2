// kdtree.cu -- exact nearest-neighbour search through a bucketed KD-tree: what
// the reference's V10/V11 were meant to be (core.cu:1059-1163 CPU KD-tree; core.cu:1289-1451 GPU KD-tree whose
// kernel body is commented out, so it returns zeros) -- here functional, on the GPU, and returning exactly
// V0's answer (same FP32 distance arithmetic as the brute-force kernels, lowest index on exact ties).
//
// Shape.  An implicit heap of bounding boxes like the reference's implicit-heap KD-tree (core.cu:1072-1114,
// 1374-1399), but BUCKETED: the leaves are the engine's 128-point reference blocks.  Two builds fill the same
// structure: on the GPU (default) the points are sorted along a Morton curve and cut into runs of 128; on the
// host (NNS_B200_TREE_HOST_BUILD=1), as the reference does it, by median splits:
//   * L = 2^ceil(log2(n / 128)) leaves; every inner node splits its points in halves (std::nth_element) along
//     the dimension of largest extent, so every leaf holds floor/ceil(n / L) <= 128 points;
//   * a leaf is stored exactly like a block of the brute-force index -- float[k + 1][128], coordinate rows
//     (lane = point), padding lanes NaN -- plus int[128] original indices;
//   * every node of the heap (2L - 1) carries its bounding box float[2][k].
// Search.  One WARP per query, lane t holds coordinate t (k <= 32).  An explicit stack in shared memory holds
// (node, lower bound); an inner node computes the box distances of both children lane-parallel over the
// dimensions and descends into the nearer one first; a leaf is scanned like the brute-force kernels scan a
// block (four points per lane, one 16-byte load per dimension, V0's subtract-square-accumulate form) and the
// packed (dist, original index) keys are min-reduced across the warp.  A subtree is skipped only when its lower
// bound, shrunk by 4 (k + 2) ulp to cover the FP32 rounding of both the bound and the distances, is STRICTLY
// greater than the best distance so far -- so every point that could equal the minimum is visited and the
// packed key resolves exact ties to the lowest original index, as V0 does (core.cu:44).
// NaN coordinates never win (V0: NaN > x is false) and are left out of the boxes; a query whose distances
// are all NaN / +INF gets index 0 (V0's initial value, core.cu:34-35).
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <mutex>
#include <vector>

#include "host_state.h"

namespace nns {

constexpr int KD_MAX_K = 32;
constexpr int KD_STACK = 48;          // depth of a tree over 2^31 points is 24; two pushes per level
constexpr int KD_WARPS = 8;           // queries (warps) per CTA

struct KdNodeBox { float lo, hi; };

template <bool EXACT>
__global__ void __launch_bounds__(32 * KD_WARPS)
kdtree_search_kernel(const float* __restrict__ queries, const int m, const int k, const float* __restrict__ blocks,
                     const int* __restrict__ perm, const float* __restrict__ boxes, const int leaves,
                     int* __restrict__ out_idx, float* __restrict__ out_dist)
{
    __shared__ int s_node[KD_WARPS][KD_STACK];
    __shared__ float s_lb[KD_WARPS][KD_STACK];
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const int q = (int)blockIdx.x * KD_WARPS + warp;
    if (q >= m) return;
    const float qt = lane < k ? __ldg(queries + (size_t)q * k + lane) : 0.0f;
    const float shrink = 1.0f - 4.0f * (float)(k + 2) * 5.9604645e-8f;
    u64 best = KEY_INIT;
    float best_d = inf_f();
    int sp = 0;
    if (lane == 0) { s_node[warp][0] = 0; s_lb[warp][0] = 0.0f; }
    sp = 1;
    __syncwarp();
    while (sp > 0) {
        --sp;
        const int node = s_node[warp][sp];
        const float lb = s_lb[warp][sp];
        __syncwarp();
        if (lb * shrink > best_d) continue;  // cannot hold a point that beats or ties the best one
        if (node >= leaves - 1) {
            // ---- leaf: one 128-point block, four points per lane ----
            const int leaf = node - (leaves - 1);
            const float4* blk = reinterpret_cast<const float4*>(blocks + (size_t)leaf * (k + 1) * LB) + lane;
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int t = 0; t < k; ++t) {
                const float4 r4 = __ldg(blk + (size_t)t * (LB / 4));
                const float c = __shfl_sync(0xffffffffu, qt, t);
                const float r[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float d = c - r[e];
                    acc[e] = EXACT ? __fadd_rn(acc[e], __fmul_rn(d, d)) : __fmaf_rn(d, d, acc[e]);
                }
            }
            const int4 id4 = __ldg(reinterpret_cast<const int4*>(perm + (size_t)leaf * LB) + lane);
            const int ids[4] = {id4.x, id4.y, id4.z, id4.w};
            u64 key = KEY_INIT;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (ids[e] >= 0 && acc[e] < inf_f()) {  // padding lanes and NaN / INF distances never win
                    const u64 cand = pack_key(acc[e], ids[e]);
                    key = cand < key ? cand : key;
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const u64 o = __shfl_xor_sync(0xffffffffu, key, off);
                key = o < key ? o : key;
            }
            if (key < best) {
                best = key;
                best_d = __uint_as_float((unsigned)(best >> 32));
            }
        } else {
            // ---- inner node: box distance of both children, lane = dimension ----
            float d2[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int child = 2 * node + 1 + c;
                float x = 0.0f;
                if (lane < k) {
                    const float lo = __ldg(boxes + ((size_t)child * 2 + 0) * k + lane);
                    const float hi = __ldg(boxes + ((size_t)child * 2 + 1) * k + lane);
                    const float e = fmaxf(fmaxf(lo - qt, qt - hi), 0.0f);  // NaN query coordinate -> 0 (never prunes)
                    x = e * e;
                    if (!(lo <= hi)) x = inf_f();  // empty box (only NaN points below): nothing to find there
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
                d2[c] = x;
            }
            const int near = d2[1] < d2[0] ? 1 : 0;
            // farther child first, so that the nearer one is popped next
            if (lane == 0) {
                int p = sp;
                if (d2[1 - near] * shrink <= best_d) { s_node[warp][p] = 2 * node + 1 + (1 - near); s_lb[warp][p] = d2[1 - near]; ++p; }
                if (d2[near] * shrink <= best_d) { s_node[warp][p] = 2 * node + 1 + near; s_lb[warp][p] = d2[near]; ++p; }
            }
            sp += (d2[1 - near] * shrink <= best_d ? 1 : 0) + (d2[near] * shrink <= best_d ? 1 : 0);
            __syncwarp();
        }
    }
    if (lane == 0) {
        out_idx[q] = (int)(unsigned)(best & 0xffffffffull);  // KEY_INIT -> index 0, V0's answer when nothing is < +INF
        if (out_dist) out_dist[q] = best_d;
    }
}

// ---------------------------------------------------------------------------------------------
// GPU build (default): Morton order instead of median splits
// ---------------------------------------------------------------------------------------------
// The search only needs spatially coherent 128-point leaves under a heap of bounding boxes; it does not care how
// the leaves were cut.  On the GPU the points are sorted along a Morton curve (floor(63 / k) bits per dimension
// over the finite bounding box; points with a NaN coordinate sort last), consecutive runs of 128 sorted points
// become the leaves, and the boxes are computed bottom-up.  The radix sort of the 64-bit codes is CUB's (library
// code, like the reference's use of Thrust in its host build); everything else is the kernels below.  An index
// of 4.2 M points is built in a few milliseconds instead of 0.5 s on the host, which makes the tree worth
// building even for a single batch of queries at low k.
__device__ __forceinline__ unsigned f2ord_u(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f_u(unsigned o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

// bbox[t] = ordered-uint min, bbox[k + t] = ordered-uint max over the finite coordinates
__global__ void kd_bbox_kernel(const float* __restrict__ aos, const int n, const int k, unsigned* __restrict__ bbox)
{
    const long long total = (long long)n * k;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float x = aos[i];
        if (fabsf(x) < inf_f()) {  // finite
            const int t = (int)(i % k);
            atomicMin(bbox + t, f2ord_u(x));
            atomicMax(bbox + k + t, f2ord_u(x));
        }
    }
}

__global__ void kd_morton_kernel(const float* __restrict__ aos, const int n, const int k, const unsigned* __restrict__ bbox,
                                 u64* __restrict__ codes, int* __restrict__ ids)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int bits = 63 / k;  // k <= 32 -> at least one bit per dimension
    u64 code = 0;
    bool nan = false;
    for (int t = 0; t < k; ++t) {
        const float x = aos[(size_t)j * k + t];
        const float lo = ord2f_u(bbox[t]), hi = ord2f_u(bbox[k + t]);
        nan |= (x != x);
        float u = (hi > lo) ? (x - lo) / (hi - lo) : 0.0f;  // +-INF clamp to the ends
        u = fminf(fmaxf(u, 0.0f), 1.0f);
        const unsigned cell = (unsigned)fminf(u * (float)(1u << bits), (float)((1u << bits) - 1u));
        for (int b = 0; b < bits; ++b)  // bit b of dimension t -> position b * k + (k - 1 - t)
            code |= (u64)((cell >> b) & 1u) << (b * k + (k - 1 - t));
    }
    codes[j] = nan ? ~0ull : code;
    ids[j] = j;
}

// leaf block `leaf` = sorted points [leaf * 128, leaf * 128 + 128): SoA rows, original indices, bounding box
__global__ void __launch_bounds__(LB)
kd_leaf_kernel(const float* __restrict__ aos, const int n, const int k, const int* __restrict__ sorted_ids, const int leaves,
               float* __restrict__ blocks, int* __restrict__ perm, float* __restrict__ boxes)
{
    const int leaf = blockIdx.x, lane = threadIdx.x;
    const long long pos = (long long)leaf * LB + lane;
    const int j = pos < n ? sorted_ids[pos] : -1;
    perm[(size_t)leaf * LB + lane] = j;
    bool nan = false;
    for (int t = 0; t < k; ++t) nan |= (j >= 0) && (aos[(size_t)j * k + t] != aos[(size_t)j * k + t]);
    __shared__ unsigned s_lo[4], s_hi[4];
    const size_t node = (size_t)leaves - 1 + leaf;
    for (int t = 0; t < k; ++t) {
        const float x = j >= 0 ? aos[(size_t)j * k + t] : nan_f();
        blocks[(size_t)leaf * (k + 1) * LB + (size_t)t * LB + lane] = x;
        // a point with a NaN coordinate can never win: it does not stretch the box
        unsigned lo = (j >= 0 && !nan) ? f2ord_u(x) : 0xffffffffu, hi = (j >= 0 && !nan) ? f2ord_u(x) : 0u;
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if ((lane & 31) == 0) { s_lo[lane >> 5] = lo; s_hi[lane >> 5] = hi; }
        __syncthreads();
        if (lane == 0) {
            const unsigned l = min(min(s_lo[0], s_lo[1]), min(s_lo[2], s_lo[3])), h = max(max(s_hi[0], s_hi[1]), max(s_hi[2], s_hi[3]));
            const bool empty = l > h;  // no point contributed: (+INF, -INF) marks an empty box
            boxes[(node * 2 + 0) * k + t] = empty ? inf_f() : ord2f_u(l);
            boxes[(node * 2 + 1) * k + t] = empty ? -inf_f() : ord2f_u(h);
        }
        __syncthreads();
    }
    if (lane == 0) blocks[(size_t)leaf * (k + 1) * LB + (size_t)k * LB] = 0.0f;  // row k (the brute-force index's norms) is unused here
}

// one heap level: nodes [first, first + count) = union of their children
__global__ void kd_inner_boxes_kernel(float* __restrict__ boxes, const int first, const int count, const int k)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count * k) return;
    const size_t node = (size_t)first + i / k;
    const int t = i % k;
    const size_t a = 2 * node + 1, b = 2 * node + 2;
    boxes[(node * 2 + 0) * k + t] = fminf(boxes[(a * 2 + 0) * k + t], boxes[(b * 2 + 0) * k + t]);
    boxes[(node * 2 + 1) * k + t] = fmaxf(boxes[(a * 2 + 1) * k + t], boxes[(b * 2 + 1) * k + t]);
}

// d_aos: the n points on the device; fills d_blocks [leaves][k+1][128], d_perm [leaves][128], d_boxes [2 leaves - 1][2][k]
static int kd_build_device(DeviceCtx* c, int k, int n, int leaves, const float* d_aos, float* d_blocks, int* d_perm, float* d_boxes,
                           cudaStream_t st)
{
    unsigned* bbox = nullptr;
    u64 *codes = nullptr, *codes2 = nullptr;
    int *ids = nullptr, *ids2 = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, codes, codes2, ids, ids2, n, 0, 64, st));
    const size_t o_codes = 256, o_codes2 = o_codes + (((size_t)n * 8 + 255) & ~(size_t)255), o_ids = o_codes2 + (((size_t)n * 8 + 255) & ~(size_t)255),
                 o_ids2 = o_ids + (((size_t)n * 4 + 255) & ~(size_t)255), o_tmp = o_ids2 + (((size_t)n * 4 + 255) & ~(size_t)255);
    unsigned char* scratch = nullptr;
    CU_TRY(cudaMallocFromPoolAsync((void**)&scratch, o_tmp + tmp_bytes, c->pool, st));
    bbox = reinterpret_cast<unsigned*>(scratch);
    codes = reinterpret_cast<u64*>(scratch + o_codes);
    codes2 = reinterpret_cast<u64*>(scratch + o_codes2);
    ids = reinterpret_cast<int*>(scratch + o_ids);
    ids2 = reinterpret_cast<int*>(scratch + o_ids2);
    tmp = scratch + o_tmp;
    cudaError_t e = cudaMemsetAsync(bbox, 0xff, (size_t)k * 4, st);               // minima: 0xffffffff
    if (e == cudaSuccess) e = cudaMemsetAsync(bbox + k, 0, (size_t)k * 4, st);    // maxima: 0
    if (e == cudaSuccess) {
        kd_bbox_kernel<<<c->num_sms * 8, 256, 0, st>>>(d_aos, n, k, bbox);
        kd_morton_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_aos, n, k, bbox, codes, ids);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, codes, codes2, ids, ids2, n, 0, 64, st);
    if (e == cudaSuccess) {
        kd_leaf_kernel<<<leaves, LB, 0, st>>>(d_aos, n, k, ids2, leaves, d_blocks, d_perm, d_boxes);
        for (int count = leaves / 2; count >= 1; count /= 2)  // heap level with `count` nodes starts at node count - 1
            kd_inner_boxes_kernel<<<(count * k + 255) / 256, 256, 0, st>>>(d_boxes, count - 1, count, k);
        e = cudaGetLastError();
    }
    count_launches(4);
    const cudaError_t fe = cudaFreeAsync(scratch, st);
    CU_TRY(e);
    CU_TRY(fe);
    return NNS_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// host build (NNS_B200_TREE_HOST_BUILD=1): median splits, the reference's way
// ---------------------------------------------------------------------------------------------
struct KdBuild {
    int k, n, leaves;
    const float* r;
    std::vector<int> order;    // point indices, leaf after leaf
    std::vector<int> leaf_lo;  // [leaves + 1] offsets into order
};

static inline float kd_key(float x) { return x != x ? std::numeric_limits<float>::infinity() : x; }  // NaN sorts last

// heap node `node` owns order[lo, hi) and covers leaves [leaf0, leaf0 + nleaves)
static void kd_split(KdBuild& b, int lo, int hi, int leaf0, int nleaves, int par_depth)
{
    if (nleaves == 1) {
        b.leaf_lo[leaf0] = lo;
        return;
    }
    const int k = b.k;
    // dimension of largest extent over the finite coordinates of this node
    int dim = 0;
    float best = -1.0f;
    for (int t = 0; t < k; ++t) {
        float mn = std::numeric_limits<float>::infinity(), mx = -mn;
        for (int i = lo; i < hi; ++i) {
            const float x = b.r[(size_t)b.order[i] * k + t];
            if (std::isfinite(x)) { mn = x < mn ? x : mn; mx = x > mx ? x : mx; }
        }
        const float ext = mx >= mn ? mx - mn : -1.0f;
        if (ext > best) { best = ext; dim = t; }
    }
    const int mid = lo + (hi - lo + 1) / 2;  // left half gets the extra point
    if (mid < hi)
        std::nth_element(b.order.begin() + lo, b.order.begin() + mid, b.order.begin() + hi, [&](int a, int c) {
            const float xa = kd_key(b.r[(size_t)a * k + dim]), xc = kd_key(b.r[(size_t)c * k + dim]);
            return xa < xc || (xa == xc && a < c);  // total order: deterministic trees
        });
    if (par_depth > 0 && hi - lo > (1 << 16)) {
        auto left = std::async(std::launch::async, [&]() { kd_split(b, lo, mid, leaf0, nleaves / 2, par_depth - 1); });
        kd_split(b, mid, hi, leaf0 + nleaves / 2, nleaves / 2, par_depth - 1);
        left.get();
    } else {
        kd_split(b, lo, mid, leaf0, nleaves / 2, 0);
        kd_split(b, mid, hi, leaf0 + nleaves / 2, nleaves / 2, 0);
    }
}

}  // namespace nns

using namespace nns;

struct nns_b200_tree {
    int k = 0, n = 0, device = 0, leaves = 0;
    DeviceCtx* ctx = nullptr;
    float* d_blocks = nullptr;
    int* d_perm = nullptr;
    float* d_boxes = nullptr;
    std::mutex mu;
};

extern "C" {

int nns_b200_tree_create(int k, int n, const float* r_points, int device, nns_b200_tree_t** out)
{
    if (!out) return fail(NNS_B200_ERR_INVALID, "NULL out");
    *out = nullptr;
    if (k <= 0 || n < 0 || (n > 0 && !r_points)) return fail(NNS_B200_ERR_INVALID, "invalid tree k=%d n=%d", k, n);
    if (k > KD_MAX_K) return fail(NNS_B200_ERR_UNSUPPORTED, "the tree search covers k <= %d", KD_MAX_K);
    DeviceCtx* c;
    ST_TRY(ctx_get(device, &c));
    int leaves = 1;
    while ((long long)leaves * LB < n) leaves *= 2;
    const size_t bf = (size_t)(k + 1) * LB;
    const size_t nodes = (size_t)2 * leaves - 1;
    const char* env = getenv("NNS_B200_TREE_HOST_BUILD");
    const bool host_build = env && atoi(env) != 0;
    std::vector<float> blocks, boxes;
    std::vector<int> perm;
    if (host_build) {
        // ---- host: median splits down to 128-point leaves ----
        KdBuild b;
        b.k = k; b.n = n; b.r = r_points;
        b.leaves = leaves;
        b.order.resize(n);
        for (int i = 0; i < n; ++i) b.order[i] = i;
        b.leaf_lo.assign(leaves + 1, n);
        if (n > 0) kd_split(b, 0, n, 0, leaves, 3);
        b.leaf_lo[leaves] = n;
        blocks.assign((size_t)leaves * bf, std::numeric_limits<float>::quiet_NaN());
        perm.assign((size_t)leaves * LB, -1);
        boxes.resize(nodes * 2 * k);
        const float inf = std::numeric_limits<float>::infinity();
        for (size_t nd = 0; nd < nodes; ++nd)
            for (int t = 0; t < k; ++t) { boxes[(nd * 2 + 0) * k + t] = inf; boxes[(nd * 2 + 1) * k + t] = -inf; }  // empty
        for (int lf = 0; lf < leaves; ++lf) {
            const int lo = b.leaf_lo[lf], hi = b.leaf_lo[lf + 1];
            // ascending original index inside a leaf (not needed for correctness -- keys decide -- but deterministic)
            std::sort(b.order.begin() + lo, b.order.begin() + hi);
            const size_t nd = (size_t)leaves - 1 + lf;
            for (int i = lo; i < hi; ++i) {
                const int j = b.order[i], lane = i - lo;
                perm[(size_t)lf * LB + lane] = j;
                bool has_nan = false;
                for (int t = 0; t < k; ++t) has_nan |= (r_points[(size_t)j * k + t] != r_points[(size_t)j * k + t]);
                for (int t = 0; t < k; ++t) {
                    const float x = r_points[(size_t)j * k + t];
                    blocks[(size_t)lf * bf + (size_t)t * LB + lane] = x;
                    if (!has_nan) {  // a point with a NaN coordinate can never win: it does not stretch the box
                        float& blo = boxes[(nd * 2 + 0) * k + t];
                        float& bhi = boxes[(nd * 2 + 1) * k + t];
                        blo = x < blo ? x : blo;
                        bhi = x > bhi ? x : bhi;
                    }
                }
            }
        }
        for (long long nd = (long long)leaves - 2; nd >= 0; --nd)  // inner nodes bottom-up: union of the children
            for (int t = 0; t < k; ++t) {
                boxes[((size_t)nd * 2 + 0) * k + t] = std::min(boxes[((size_t)(2 * nd + 1) * 2 + 0) * k + t], boxes[((size_t)(2 * nd + 2) * 2 + 0) * k + t]);
                boxes[((size_t)nd * 2 + 1) * k + t] = std::max(boxes[((size_t)(2 * nd + 1) * 2 + 1) * k + t], boxes[((size_t)(2 * nd + 2) * 2 + 1) * k + t]);
            }
    }
    // ---- device ----
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard guard;
    ST_TRY(guard.enter(c->device));
    nns_b200_tree* h = new nns_b200_tree();
    h->k = k; h->n = n; h->device = c->device; h->leaves = leaves; h->ctx = c;
    cudaError_t e = cudaMalloc((void**)&h->d_blocks, (size_t)leaves * bf * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_perm, (size_t)leaves * LB * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_boxes, nodes * 2 * k * sizeof(float));
    int st = e == cudaSuccess ? NNS_B200_OK : fail_cuda(e, __FILE__, __LINE__);
    if (host_build) {
        if (st == NNS_B200_OK) st = h2d_async(c, h->d_blocks, blocks.data(), blocks.size() * sizeof(float), c->copy);
        if (st == NNS_B200_OK) st = h2d_async(c, h->d_perm, perm.data(), perm.size() * sizeof(int), c->copy);
        if (st == NNS_B200_OK) st = h2d_async(c, h->d_boxes, boxes.data(), boxes.size() * sizeof(float), c->copy);
    } else if (n > 0) {
        // ---- GPU build: upload the points (staged when pageable), Morton sort, leaves, boxes ----
        if (st == NNS_B200_OK) st = buf_reserve(&c->r, (size_t)n * k * sizeof(float));
        if (st == NNS_B200_OK) st = ctx_events(c, 1);
        if (st == NNS_B200_OK) st = h2d_async(c, c->r.p, r_points, (size_t)n * k * sizeof(float), c->copy);
        if (st == NNS_B200_OK) {
            e = cudaEventRecord(c->events[0], c->copy);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(c->compute, c->events[0], 0);
            if (e != cudaSuccess) st = fail_cuda(e, __FILE__, __LINE__);
        }
        if (st == NNS_B200_OK) st = kd_build_device(c, k, n, leaves, (const float*)c->r.p, h->d_blocks, h->d_perm, h->d_boxes, c->compute);
    }
    if (st == NNS_B200_OK) {
        e = cudaStreamSynchronize(c->copy);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->compute);
        if (e != cudaSuccess) st = fail_cuda(e, __FILE__, __LINE__);
    }
    if (st != NNS_B200_OK) {
        cudaFree(h->d_blocks); cudaFree(h->d_perm); cudaFree(h->d_boxes);
        delete h;
        return st;
    }
    *out = h;
    return NNS_B200_OK;
}

int nns_b200_tree_search(nns_b200_tree_t* h, int m, const float* s_points, int* results, float* distances)
{
    if (!h) return fail(NNS_B200_ERR_INVALID, "NULL tree");
    if (m < 0 || (m > 0 && (!s_points || !results))) return fail(NNS_B200_ERR_INVALID, "invalid queries");
    if (m == 0) return NNS_B200_OK;
    DeviceCtx* c = h->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    std::lock_guard<std::mutex> lh(h->mu);
    DeviceGuard guard;
    ST_TRY(guard.enter(c->device));
    const int k = h->k;
    ST_TRY(buf_reserve(&c->q, (size_t)m * k * sizeof(float)));
    ST_TRY(buf_reserve(&c->idx, (size_t)m * sizeof(int) * 2));
    int* d_idx = (int*)c->idx.p;
    float* d_dist = (float*)(d_idx + m);
    ST_TRY(ctx_events(c, 1));
    ST_TRY(h2d_async(c, c->q.p, s_points, (size_t)m * k * sizeof(float), c->copy));
    CU_TRY(cudaEventRecord(c->events[0], c->copy));
    CU_TRY(cudaStreamWaitEvent(c->compute, c->events[0], 0));
    if (h->n == 0) {
        CU_TRY(cudaMemsetAsync(d_idx, 0, (size_t)m * sizeof(int), c->compute));
        if (distances) for (int i = 0; i < m; ++i) distances[i] = std::numeric_limits<float>::infinity();
    } else {
        const unsigned grid = (unsigned)((m + KD_WARPS - 1) / KD_WARPS);
        if (host_flags() & NNS_B200_FLAG_V0_ROUNDING)
            kdtree_search_kernel<true><<<grid, 32 * KD_WARPS, 0, c->compute>>>((const float*)c->q.p, m, k, h->d_blocks, h->d_perm, h->d_boxes,
                                                                                h->leaves, d_idx, distances ? d_dist : nullptr);
        else
            kdtree_search_kernel<false><<<grid, 32 * KD_WARPS, 0, c->compute>>>((const float*)c->q.p, m, k, h->d_blocks, h->d_perm, h->d_boxes,
                                                                                 h->leaves, d_idx, distances ? d_dist : nullptr);
        CU_TRY(cudaGetLastError());
        count_launches(1);
        if (distances) CU_TRY(cudaMemcpyAsync(distances, d_dist, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
    }
    CU_TRY(cudaMemcpyAsync(results, d_idx, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, c->compute));
    CU_TRY(stream_drain(c->compute));
    return NNS_B200_OK;
}

int nns_b200_tree_destroy(nns_b200_tree_t* h)
{
    if (!h) return NNS_B200_OK;
    {
        std::lock_guard<std::mutex> lk(h->ctx->mu);
        DeviceGuard guard;
        if (guard.enter(h->device) == NNS_B200_OK) {
            cudaStreamSynchronize(h->ctx->compute);
            cudaFree(h->d_blocks);
            cudaFree(h->d_perm);
            cudaFree(h->d_boxes);
        }
    }
    delete h;
    return NNS_B200_OK;
}

}  // extern "C"
