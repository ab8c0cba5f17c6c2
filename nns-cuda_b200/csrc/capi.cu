// capi.cu -- the C ABI of include/nns_b200.h: planning, per-device state, host-pointer
// ingest, single-process multi-GPU fan-out.  All compute is in the CUDA kernels of this
// library; there is no CPU search path anywhere in this file.
#include "../../include/nns_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "nns_internal.h"

using namespace nns;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local char g_err_file[128] = "";
static thread_local int g_err_line = 0;
static thread_local int g_err_code = 0;

static int fail(int status, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

static int fail_cuda(cudaError_t e, const char* file, int line)
{
    snprintf(g_err_file, sizeof(g_err_file), "%s", file);
    g_err_line = line;
    g_err_code = (int)e;
    snprintf(g_err, sizeof(g_err), "%s:%d, code:%d, reason: %s", file, line, (int)e,
             cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? NNS_B200_ERR_NOMEM : NNS_B200_ERR_CUDA;
}

#define CU_TRY(call)                                                   \
    do {                                                               \
        const cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) return fail_cuda(e__, __FILE__, __LINE__); \
    } while (0)

#define ST_TRY(call)                        \
    do {                                    \
        const int st__ = (call);            \
        if (st__ != NNS_B200_OK) return st__; \
    } while (0)

static std::atomic<unsigned long long> g_launches{0};

// ---------------------------------------------------------------------------------------------
// planning (pure host logic)
// ---------------------------------------------------------------------------------------------
struct Plan {
    int path;    // 0 low-k, 1 wide, 2 tensor
    int q;       // queries per thread (low-k)
    int warps;   // consumer warps per CTA (low-k)
    int stages;  // ring depth (low-k)
    int nqb;     // query blocks / groups (grid.x)
    int splits;  // reference splits (grid.y)
    int bps;     // reference blocks per split
    int smem;    // dynamic shared memory bytes
};

// k <= 32: the split-precision tcgen05 screen (csrc/tensor_search.cu) beats the FP32 screened kernel
// once the problem amortises its fixed cost (query image, four launches, per-CTA TMEM set-up):
// B200, profiles/r1_tensor_*: C2 (k = 3, 2.7e11 pairs) 18.6 ms vs 39.9 ms, C3 (k = 16) 0.49 s vs 2.6 s,
// C1 (6.7e7 pairs) 0.38 ms vs 0.064 ms; the reference's largest shape (k = 16, m = 1024, n = 2^20,
// 1.1e9 pairs) 8.1 ms vs 6.7 ms end to end.  Both paths return identical indices.
static bool lowk_prefers_tensor(int k, int m, int n)
{
    const double pairs = (double)m * (double)n;
    return m >= 1024 && pairs >= (k <= 8 ? 4e9 : 2e9);
}

typedef int (*occ_fn)(void* user, int k, int q, int mode, int warps, int stages);

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// register estimate used when no device is available to ask (tests / nns_b200_plan)
static int est_ctas_per_sm(int k, int q, int warps, int stages)
{
    const int threads = (warps + 1) * 32;
    const int smem = LOWK_BAR_BYTES + stages * lowk_tile_bytes(k);
    int regs = 40 + q * (k + 10);
    if (regs > 224) regs = 224;
    regs = (regs + 7) & ~7;
    int by_regs = 65536 / (regs * threads);
    int by_smem = (227 * 1024) / (smem + 1024);
    int by_thr = 2048 / threads;
    int c = by_regs < by_smem ? by_regs : by_smem;
    c = c < by_thr ? c : by_thr;
    return c < 1 ? 1 : (c > 32 ? 32 : c);
}

// choose reference splits for `nqb` query blocks: minimise waves * (blocks per CTA + overhead)
static void choose_splits(int nqb, int nblocks, int tb, int slots, double overhead_blocks, int* splits,
                          int* bps, double* cost)
{
    const int max_s = ceil_div(nblocks, tb) < 65535 ? ceil_div(nblocks, tb) : 65535;
    double best = 1e300;
    int best_s = 1, best_bps = nblocks;
    int last_bps = -1;
    for (int s = 1; s <= max_s; ++s) {
        int b = ceil_div(nblocks, s);
        b = ceil_div(b, tb) * tb;  // whole tiles per split
        if (b == last_bps) continue;
        last_bps = b;
        const int s_eff = ceil_div(nblocks, b);
        const double ctas = (double)nqb * s_eff;
        const double waves = (double)((long long)((ctas + slots - 1) / slots));
        const double c = waves * ((double)b + overhead_blocks);
        if (c < best * 0.999) { best = c; best_s = s_eff; best_bps = b; }
        if ((long long)nqb * s_eff > 64LL * slots && waves > 16) break;  // deep enough
    }
    *splits = best_s;
    *bps = best_bps;
    if (cost) *cost = best;
}

static int lowk_mode_of(unsigned flags)
{
    if (flags & NNS_B200_FLAG_V0_ROUNDING) return LOWK_EXACT_V0;
    if (flags & NNS_B200_FLAG_EXACT_FORM) return LOWK_EXACT_FMA;
    return LOWK_FILTER;
}

static int make_plan(int k, int m, int n, unsigned flags, int num_sms, occ_fn occ, void* occ_user, Plan* p)
{
    if (k <= 0 || m < 0 || n < 0) return fail(NNS_B200_ERR_INVALID, "invalid shape k=%d m=%d n=%d", k, m, n);
    if (num_sms <= 0) num_sms = 148;
    const int mode = lowk_mode_of(flags);
    const bool exact = mode == LOWK_EXACT_V0;
    const int nblocks = ceil_div(n, LB);
    memset(p, 0, sizeof(*p));
    const bool tensor_ok = k <= TENSOR_MAX_K;
    if ((flags & NNS_B200_FLAG_FORCE_TENSOR) && !tensor_ok)
        return fail(NNS_B200_ERR_UNSUPPORTED, "tensor path needs k <= %d", TENSOR_MAX_K);
    const bool tensor_auto = k > LOWK_MAX_K ? m >= 256 : (lowk_prefers_tensor(k, m, n) && !(flags & NNS_B200_FLAG_EXACT_FORM));
    if (tensor_ok && !(flags & (NNS_B200_FLAG_FORCE_WIDE | NNS_B200_FLAG_FORCE_LOWK)) &&
        ((flags & NNS_B200_FLAG_FORCE_TENSOR) || tensor_auto)) {
        // tcgen05 path: one CTA per 256-query strip x reference range (splits chosen at launch)
        p->path = 2;
        p->q = 1;
        p->warps = 10;
        p->stages = 4;
        p->nqb = ceil_div(m, 256);
        p->splits = 1;
        p->bps = nblocks;
        p->smem = 0;
        return NNS_B200_OK;
    }
    bool lowk = (k <= LOWK_MAX_K) && m >= 16;
    if (flags & NNS_B200_FLAG_FORCE_LOWK) {
        if (k > LOWK_MAX_K) return fail(NNS_B200_ERR_UNSUPPORTED, "low-k path needs k <= %d", LOWK_MAX_K);
        lowk = true;
    }
    if (flags & NNS_B200_FLAG_FORCE_WIDE) lowk = false;

    if (!lowk) {
        const size_t smem = (size_t)k * WIDE_QT * sizeof(float);
        if (smem > 200 * 1024) return fail(NNS_B200_ERR_UNSUPPORTED, "k=%d too large for the wide path", k);
        p->path = 1;
        p->q = WIDE_QT;
        p->warps = WIDE_THREADS / 32;
        p->nqb = ceil_div(m, WIDE_QT);
        const int slots = num_sms * 8;
        int s = 1;
        if (p->nqb < 4 * slots) s = ceil_div(4LL * slots, p->nqb > 0 ? p->nqb : 1);
        const int max_s = ceil_div(nblocks, 2) > 0 ? ceil_div(nblocks, 2) : 1;
        if (s > max_s) s = max_s;
        if (s > 65535) s = 65535;
        p->bps = ceil_div(nblocks, s) > 0 ? ceil_div(nblocks, s) : 1;
        p->bps = ceil_div(p->bps, WIDE_THREADS / 32) * (WIDE_THREADS / 32);  // one block per warp at a time: whole rounds
        p->splits = ceil_div(nblocks, p->bps) > 0 ? ceil_div(nblocks, p->bps) : 1;
        p->smem = (int)smem;
        return NNS_B200_OK;
    }

    // low-k: enumerate (q, warps), model the time, keep the cheapest
    const int q_over = (int)((flags >> 8) & 0xff), w_over = (int)((flags >> 16) & 0xff);
    const int st_over = (int)((flags >> 24) & 0xf);
    const int tb = lowk_tb(k);
    const int stages = st_over ? st_over : 4;
    if (stages < 2 || stages > LOWK_MAX_STAGES) return fail(NNS_B200_ERR_INVALID, "stages override %d", stages);
    double best_cost = 1e300;
    const int qs[2] = {lowk_q_default(k), lowk_q_alt(k)};
    for (int qi = 0; qi < 2; ++qi) {
        const int q = qs[qi];
        if (q_over ? q != q_over : qi != 0) continue;  // the alternative blocking only on request
        if (exact && q != lowk_q_default(k)) continue;
        for (int w = 8; w >= 1; w >>= 1) {
            if (w_over && w != w_over) continue;
            const int qb = 32 * w * q;
            const int nqb = ceil_div(m, qb);
            int cps = occ ? occ(occ_user, k, q, mode, w, stages) : est_ctas_per_sm(k, q, w, stages);
            if (cps < 1) cps = 1;
            const int slots = num_sms * cps;
            // work of one CTA per reference block, in SM cycles: 128 refs * q queries * 2k FP32
            // lane-slots per lane; W*cps warps share 4 schedulers
            const double share = (double)(w * cps) / 4.0;
            const double lane_slots = (mode == LOWK_FILTER ? 1.0 : mode == LOWK_EXACT_FMA ? 2.0 : 3.0) * k + 1.5;
            const double block_cycles = 128.0 * q * lane_slots * (share > 1.0 ? share : 1.0) * (1.0 + 0.15 / q);
            const double overhead_blocks = 6000.0 / block_cycles;  // launch/prologue/atomics
            int s, bps;
            double c;
            choose_splits(nqb, nblocks > 0 ? nblocks : 1, tb, slots, overhead_blocks, &s, &bps, &c);
            const double cost = c * block_cycles;
            if (cost < best_cost * 0.98) {
                best_cost = cost;
                p->path = 0; p->q = q; p->warps = w; p->stages = stages;
                p->nqb = nqb; p->splits = s; p->bps = bps;
                p->smem = LOWK_BAR_BYTES + stages * lowk_tile_bytes(k);
            }
        }
    }
    if (best_cost >= 1e300) return fail(NNS_B200_ERR_INVALID, "no low-k configuration matches the overrides");
    return NNS_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// kernel dispatch
// ---------------------------------------------------------------------------------------------
static cudaError_t lowk_dispatch(int k, int q, int mode, const LowkArgs& a, int* occ)
{
    typedef cudaError_t (*range_fn)(int, int, int, const LowkArgs&, int*);
    static const range_fn table[16] = {
        lowk_launch_range_0,  lowk_launch_range_1,  lowk_launch_range_2,  lowk_launch_range_3,
        lowk_launch_range_4,  lowk_launch_range_5,  lowk_launch_range_6,  lowk_launch_range_7,
        lowk_launch_range_8,  lowk_launch_range_9,  lowk_launch_range_10, lowk_launch_range_11,
        lowk_launch_range_12, lowk_launch_range_13, lowk_launch_range_14, lowk_launch_range_15};
    if (k < 1 || k > LOWK_MAX_K) return cudaErrorInvalidValue;
    return table[(k - 1) / 2](k, q, mode, a, occ);
}

// ---------------------------------------------------------------------------------------------
// per-device state
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct DeviceCtx {
    int device = -1;
    int num_sms = 0;
    bool ready = false;
    std::mutex mu;  // serialises users of the cached buffers/streams of this device
    cudaStream_t compute = nullptr, copy = nullptr;
    std::vector<cudaEvent_t> events;
    DevBuf q, r, index, keys, idx, stats, peer_keys;
    std::map<std::tuple<int, int, int, int, int>, int> occ_cache;
};

static std::mutex g_ctx_mu;
static std::map<int, DeviceCtx*> g_ctx;

static int ctx_get(int device, DeviceCtx** out)
{
    if (device < 0) CU_TRY(cudaGetDevice(&device));
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    auto it = g_ctx.find(device);
    DeviceCtx* c;
    if (it == g_ctx.end()) {
        c = new DeviceCtx();
        c->device = device;
        g_ctx[device] = c;
    } else {
        c = it->second;
    }
    if (!c->ready) {
        int prev = 0;
        CU_TRY(cudaGetDevice(&prev));
        CU_TRY(cudaSetDevice(device));
        cudaDeviceProp prop;
        CU_TRY(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) {
            cudaSetDevice(prev);
            return fail(NNS_B200_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only",
                        device, prop.major, prop.minor);
        }
        c->num_sms = prop.multiProcessorCount;
        // the tensor path takes its scratch (query image, candidate records) from the stream-ordered
        // allocator; keep freed blocks in the pool across synchronisations instead of returning
        // ~100 MB to the driver after every host-pointer call
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        CU_TRY(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
        c->ready = true;
        CU_TRY(cudaSetDevice(prev));
    }
    *out = c;
    return NNS_B200_OK;
}

static int buf_reserve(DevBuf* b, size_t bytes)
{
    if (bytes <= b->cap) return NNS_B200_OK;
    if (b->p) CU_TRY(cudaFree(b->p));
    b->p = nullptr;
    b->cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    CU_TRY(cudaMalloc(&b->p, want));
    b->cap = want;
    return NNS_B200_OK;
}

static int occ_query(void* user, int k, int q, int mode, int warps, int stages)
{
    DeviceCtx* c = (DeviceCtx*)user;
    const auto key = std::make_tuple(k, q, mode, warps, stages);
    auto it = c->occ_cache.find(key);
    if (it != c->occ_cache.end()) return it->second;
    LowkArgs a{};
    a.warps = warps;
    a.stages = stages;
    int occ = 0;
    if (lowk_dispatch(k, q, mode, a, &occ) != cudaSuccess) {
        cudaGetLastError();
        occ = 0;
    }
    c->occ_cache[key] = occ;
    return occ;
}

// The hot path on device-resident data: plan, launch.  `c` supplies num_sms and the occupancy
// cache; the caller must have made c->device current.
static int search_keys_on(DeviceCtx* c, int k, int m, int n, const float* d_queries, const float* d_header,
                          const float* d_blocks, int index_base, u64* d_keys, unsigned flags, cudaStream_t st)
{
    if (m == 0 || n == 0) return NNS_B200_OK;
    Plan p;
    ST_TRY(make_plan(k, m, n, flags, c->num_sms, occ_query, c, &p));
    const int mode = lowk_mode_of(flags);
    const int nblocks = ceil_div(n, LB);
    if (p.path == 2) {
        // the tensor section (centre, |r'|^2, BF16 image) follows the FP32 blocks of the whole index
        const float* d_section = d_blocks + (size_t)nblocks * index_block_floats(k);
        int launches = 0;
        ST_TRY(buf_reserve(&c->stats, 64));
        // per-call status words {candidates, overflow flag, capacity} from the stream-ordered pool: the
        // flag gates the fallback kernel below, so it must not be shared with a search that another
        // stream of this device has in flight
        unsigned* d_status = nullptr;
        CU_TRY(cudaMallocAsync((void**)&d_status, 64, st));
        cudaError_t te = tensor_search(k, m, n, d_queries, d_blocks, d_section, index_base, d_keys, mode == LOWK_EXACT_V0,
                                       c->num_sms, st, &launches, d_status, (flags & NNS_B200_FLAG_TEST_TINY_CANDIDATES) != 0);
        g_launches.fetch_add((unsigned long long)launches + 1, std::memory_order_relaxed);
        // Fallback for data whose near-ties overflow the candidate buffer (e.g. all points identical, or
        // clusters far denser than the BF16 screen resolves): the FP32 kernel of this shape, launched
        // unconditionally behind the device-side flag the re-score kernel leaves in d_status[1]; it
        // exits at once when the flag is clear.  No host round trip.
        int fst = NNS_B200_OK;
        if (te == cudaSuccess) {
            const int* enable = (const int*)d_status + 1;
            unsigned fb = (flags & ~(NNS_B200_FLAG_FORCE_TENSOR | NNS_B200_FLAG_TEST_TINY_CANDIDATES)) |
                          ((k <= LOWK_MAX_K && m >= 16) ? NNS_B200_FLAG_FORCE_LOWK : NNS_B200_FLAG_FORCE_WIDE);
            fst = make_plan(k, m, n, fb, c->num_sms, occ_query, c, &p);
            if (fst == NNS_B200_OK && p.path == 0) {
                LowkArgs a{};
                a.queries = d_queries; a.m = m; a.header = d_header; a.blocks = d_blocks; a.nblocks = nblocks;
                a.blocks_per_split = p.bps; a.index_base = index_base; a.keys = d_keys;
                a.warps = p.warps; a.stages = p.stages; a.nqb = p.nqb; a.splits = p.splits; a.stream = st;
                a.enable = enable;
                te = lowk_dispatch(k, p.q, mode, a, nullptr);
            } else if (fst == NNS_B200_OK) {
                WideArgs a{};
                a.queries = d_queries; a.m = m; a.k = k; a.blocks = d_blocks; a.nblocks = nblocks;
                a.blocks_per_split = p.bps; a.index_base = index_base; a.keys = d_keys;
                a.nqg = p.nqb; a.splits = p.splits; a.stream = st; a.enable = enable;
                te = wide_launch(mode == LOWK_EXACT_V0, a);
            }
            // keep the last search's status words for nns_b200_tensor_stats()
            if (te == cudaSuccess) te = cudaMemcpyAsync(c->stats.p, d_status, 3 * sizeof(unsigned), cudaMemcpyDeviceToDevice, st);
        }
        const cudaError_t fe = cudaFreeAsync(d_status, st);
        ST_TRY(fst);
        CU_TRY(te);
        CU_TRY(fe);
        return NNS_B200_OK;
    }
    if (p.path == 0) {
        LowkArgs a{};
        a.queries = d_queries; a.m = m; a.header = d_header; a.blocks = d_blocks; a.nblocks = nblocks;
        a.blocks_per_split = p.bps; a.index_base = index_base; a.keys = d_keys;
        a.warps = p.warps; a.stages = p.stages; a.nqb = p.nqb; a.splits = p.splits; a.stream = st;
        CU_TRY(lowk_dispatch(k, p.q, mode, a, nullptr));
    } else {
        WideArgs a{};
        a.queries = d_queries; a.m = m; a.k = k; a.blocks = d_blocks; a.nblocks = nblocks;
        a.blocks_per_split = p.bps; a.index_base = index_base; a.keys = d_keys;
        a.nqg = p.nqb; a.splits = p.splits; a.stream = st;
        CU_TRY(wide_launch(mode == LOWK_EXACT_V0, a));
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return NNS_B200_OK;
}

struct DeviceGuard {
    int prev = -1;
    bool active = false;
    int enter(int device)
    {
        CU_TRY(cudaGetDevice(&prev));
        if (prev != device) CU_TRY(cudaSetDevice(device));
        active = true;
        return NNS_B200_OK;
    }
    ~DeviceGuard()
    {
        if (active && prev >= 0) cudaSetDevice(prev);
    }
};

static unsigned host_flags();

// Host arrays -> device -> keys (h_keys != NULL) or indices (h_idx != NULL) on the host.
// References are ingested in chunks: the H2D copy of chunk c+1 (copy stream) overlaps the
// index build + search of chunk c (compute stream); every chunk accumulates into the same
// packed keys with its own index base.
static int search_host_on(DeviceCtx* c, int k, int m, int n, const float* s, const float* r, int index_base,
                          u64* h_keys, int* h_idx, u64* ext_keys = nullptr, float* h_dist = nullptr)
{
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard guard;
    ST_TRY(guard.enter(c->device));
    const size_t qbytes = (size_t)m * k * sizeof(float);
    const size_t rbytes = (size_t)n * k * sizeof(float);
    const size_t ibytes = nns_b200_index_floats(k, n) * sizeof(float);
    ST_TRY(buf_reserve(&c->q, qbytes));
    ST_TRY(buf_reserve(&c->r, rbytes));
    ST_TRY(buf_reserve(&c->index, ibytes));
    ST_TRY(buf_reserve(&c->keys, (size_t)m * sizeof(u64)));
    ST_TRY(buf_reserve(&c->idx, (size_t)m * sizeof(int) * (h_dist ? 2 : 1)));
    float* d_q = (float*)c->q.p;
    float* d_r = (float*)c->r.p;
    float* d_index = (float*)c->index.p;
    // ext_keys: an already initialised key array, possibly in a PEER GPU's memory (NVLink P2P).
    // The search accumulates into this GPU's own keys; one merge kernel then atomicMin's them
    // into ext_keys, i.e. the cross-GPU (dist, idx) reduction is m device-side atomics over NVLink.
    u64* d_keys = (u64*)c->keys.p;
    int* d_idx = (int*)c->idx.p;

    CU_TRY(cudaMemcpyAsync(d_q, s, qbytes, cudaMemcpyHostToDevice, c->compute));
    CU_TRY(launch_keys_init(d_keys, m, c->compute));
    g_launches.fetch_add(h_idx ? 2 : 1, std::memory_order_relaxed);  // keys init (+ unpack below)

    // chunk = about 32 MiB of AoS reference data, a whole number of reference blocks
    long long chunk = ((32ll << 20) / ((long long)k * 4)) / LB * LB;
    if (chunk < LB) chunk = LB;
    // The tensor section (centre, BF16 operand image) is only built when this call is planned onto the
    // tcgen05 path; the planner's choice is monotone in n, so the chunks of a search that is not
    // never pick it either.  The centre needs the whole reference set: one chunk.
    Plan whole{};
    if (n > 0) ST_TRY(make_plan(k, m, n, host_flags(), c->num_sms, nullptr, nullptr, &whole));
    const bool has_tensor = whole.path == 2 && tensor_section_floats(k, n) != 0;
    if (has_tensor) chunk = ((long long)n + LB - 1) / LB * LB;
    const int nchunks = n > 0 ? (int)((n + chunk - 1) / chunk) : 0;
    while ((int)c->events.size() < nchunks) {
        cudaEvent_t ev;
        CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        c->events.push_back(ev);
    }
    for (int ci = 0; ci < nchunks; ++ci) {
        const long long j0 = (long long)ci * chunk;
        const int cn = (int)((n - j0) < chunk ? (n - j0) : chunk);
        CU_TRY(cudaMemcpyAsync(d_r + j0 * k, r + j0 * k, (size_t)cn * k * sizeof(float),
                               cudaMemcpyHostToDevice, c->copy));
        CU_TRY(cudaEventRecord(c->events[ci], c->copy));
        CU_TRY(cudaStreamWaitEvent(c->compute, c->events[ci], 0));
        float* d_blocks_c = d_index + INDEX_HEADER_FLOATS + (j0 / LB) * (long long)index_block_floats(k);
        CU_TRY(launch_index_build(k, cn, d_r + j0 * k, d_index, d_blocks_c, ci == 0, c->compute));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (has_tensor)
            CU_TRY(tensor_index_build(k, cn, d_r, d_index + INDEX_HEADER_FLOATS + (size_t)ceil_div(cn, LB) * index_block_floats(k),
                                      c->compute));
        ST_TRY(search_keys_on(c, k, m, cn, d_q, d_index, d_blocks_c, index_base + (int)j0, d_keys, host_flags(),
                              c->compute));
    }
    if (ext_keys) {
        CU_TRY(launch_keys_merge(ext_keys, d_keys, m, c->compute));
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    if (h_keys) {
        CU_TRY(cudaMemcpyAsync(h_keys, d_keys, (size_t)m * sizeof(u64), cudaMemcpyDeviceToHost, c->compute));
    }
    if (h_idx) {
        float* d_dist = h_dist ? (float*)(d_idx + m) : nullptr;  // second half of the idx buffer
        CU_TRY(launch_keys_unpack(d_keys, m, d_idx, d_dist, c->compute));
        CU_TRY(cudaMemcpyAsync(h_idx, d_idx, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, c->compute));
        if (h_dist) CU_TRY(cudaMemcpyAsync(h_dist, d_dist, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
    }
    CU_TRY(cudaStreamSynchronize(c->compute));
    CU_TRY(cudaStreamSynchronize(c->copy));
    return NNS_B200_OK;
}

// NNS_B200_FLAGS (environment, a C integer literal) applies the flags word to the host-pointer
// entry points, whose reference signature has no flags argument.
static unsigned host_flags()
{
    static const unsigned f = []() {
        const char* e = getenv("NNS_B200_FLAGS");
        return e ? (unsigned)strtoul(e, nullptr, 0) : 0u;
    }();
    return f;
}

static int check_host_args(int k, int m, int n, const void* s, const void* r, const void* out)
{
    if (k <= 0 || m < 0 || n < 0) return fail(NNS_B200_ERR_INVALID, "invalid shape k=%d m=%d n=%d", k, m, n);
    if ((long long)k * m > 0x7fffffffLL * 4 || (long long)k * n > 0x7fffffffLL * 4)
        return fail(NNS_B200_ERR_INVALID, "shape too large");
    if ((m > 0 && (!s || !out)) || (n > 0 && m > 0 && !r)) return fail(NNS_B200_ERR_INVALID, "NULL array");
    return NNS_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// exported
// ---------------------------------------------------------------------------------------------
extern "C" {

int nns_b200_version(void) { return NNS_B200_VERSION; }
const char* nns_b200_last_error(void) { return g_err; }
unsigned long long nns_b200_launch_count(void) { return g_launches.load(); }

int nns_b200_init(int device)
{
    DeviceCtx* c;
    return ctx_get(device, &c);
}

int nns_b200_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    for (auto& kv : g_ctx) {
        DeviceCtx* c = kv.second;
        std::lock_guard<std::mutex> lk2(c->mu);
        if (c->ready && cudaSetDevice(c->device) == cudaSuccess) {
            cudaStreamSynchronize(c->compute);
            cudaStreamSynchronize(c->copy);
            for (DevBuf* b : {&c->q, &c->r, &c->index, &c->keys, &c->idx, &c->stats, &c->peer_keys}) {
                if (b->p) cudaFree(b->p);
                b->p = nullptr;
                b->cap = 0;
            }
            for (cudaEvent_t ev : c->events) cudaEventDestroy(ev);
            c->events.clear();
            cudaStreamDestroy(c->compute);
            cudaStreamDestroy(c->copy);
            // hand the stream-ordered scratch kept in the device's default pool back to the driver
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
            cudaGetLastError();
            c->ready = false;
            c->occ_cache.clear();
        }
    }
    if (prev >= 0) cudaSetDevice(prev);
    return NNS_B200_OK;
}

size_t nns_b200_index_floats(int k, int n)
{
    if (k <= 0 || n <= 0) return 0;
    return (size_t)INDEX_HEADER_FLOATS + (size_t)ceil_div(n, LB) * (size_t)index_block_floats(k) +
           tensor_section_floats(k, n);
}

int nns_b200_index_build(int k, int n, const float* d_refs_aos, float* d_index, void* stream)
{
    if (k <= 0 || n < 0) return fail(NNS_B200_ERR_INVALID, "invalid shape k=%d n=%d", k, n);
    if (n > 0 && (!d_refs_aos || !d_index)) return fail(NNS_B200_ERR_INVALID, "NULL array");
    if (((uintptr_t)d_index & 15) != 0) return fail(NNS_B200_ERR_INVALID, "index must be 16-byte aligned");
    CU_TRY(launch_index_build(k, n, d_refs_aos, d_index, d_index + INDEX_HEADER_FLOATS, true, (cudaStream_t)stream));
    CU_TRY(tensor_index_build(k, n, d_refs_aos, d_index + INDEX_HEADER_FLOATS + (size_t)ceil_div(n, LB) * index_block_floats(k),
                              (cudaStream_t)stream));
    g_launches.fetch_add(n > 0 ? 1 : 0, std::memory_order_relaxed);
    return NNS_B200_OK;
}

int nns_b200_keys_init(uint64_t* d_keys, int m, void* stream)
{
    if (m < 0 || (m > 0 && !d_keys)) return fail(NNS_B200_ERR_INVALID, "invalid keys");
    CU_TRY(launch_keys_init((u64*)d_keys, m, (cudaStream_t)stream));
    g_launches.fetch_add(m > 0 ? 1 : 0, std::memory_order_relaxed);
    return NNS_B200_OK;
}

int nns_b200_keys_unpack(const uint64_t* d_keys, int m, int* d_idx, float* d_dist, void* stream)
{
    if (m < 0 || (m > 0 && (!d_keys || !d_idx))) return fail(NNS_B200_ERR_INVALID, "invalid keys");
    CU_TRY(launch_keys_unpack((const u64*)d_keys, m, d_idx, d_dist, (cudaStream_t)stream));
    g_launches.fetch_add(m > 0 ? 1 : 0, std::memory_order_relaxed);
    return NNS_B200_OK;
}

int nns_b200_search_keys(int k, int m, int n, const float* d_queries, const float* d_index, int index_base,
                         uint64_t* d_keys, unsigned flags, void* stream)
{
    if (k <= 0 || m < 0 || n < 0) return fail(NNS_B200_ERR_INVALID, "invalid shape k=%d m=%d n=%d", k, m, n);
    if (m > 0 && n > 0 && (!d_queries || !d_index || !d_keys)) return fail(NNS_B200_ERR_INVALID, "NULL array");
    if ((long long)index_base + n > 0x7fffffffLL) return fail(NNS_B200_ERR_INVALID, "index_base + n overflows int32");
    if (((uintptr_t)d_index & 15) != 0) return fail(NNS_B200_ERR_INVALID, "index must be 16-byte aligned");
    DeviceCtx* c;
    ST_TRY(ctx_get(-1, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    return search_keys_on(c, k, m, n, d_queries, d_index, d_index + INDEX_HEADER_FLOATS, index_base, (u64*)d_keys, flags,
                          (cudaStream_t)stream);
}

size_t nns_b200_workspace_bytes(int k, int m, int n)
{
    const size_t ib = (nns_b200_index_floats(k, n) * sizeof(float) + 255) & ~(size_t)255;
    const size_t kb = ((size_t)(m > 0 ? m : 0) * sizeof(u64) + 255) & ~(size_t)255;
    return ib + kb + 256;
}

int nns_b200_search_device(int k, int m, int n, const float* d_queries, const float* d_refs_aos, int* d_idx,
                           void* d_workspace, size_t workspace_bytes, unsigned flags, void* stream)
{
    if (k <= 0 || m < 0 || n < 0) return fail(NNS_B200_ERR_INVALID, "invalid shape k=%d m=%d n=%d", k, m, n);
    if (m == 0) return NNS_B200_OK;
    if (!d_idx || !d_queries || (n > 0 && !d_refs_aos)) return fail(NNS_B200_ERR_INVALID, "NULL array");
    if (workspace_bytes < nns_b200_workspace_bytes(k, m, n) || !d_workspace)
        return fail(NNS_B200_ERR_INVALID, "workspace too small");
    char* w = (char*)(((uintptr_t)d_workspace + 255) & ~(uintptr_t)255);
    float* d_index = (float*)w;
    u64* d_keys = (u64*)(w + ((nns_b200_index_floats(k, n) * sizeof(float) + 255) & ~(size_t)255));
    cudaStream_t st = (cudaStream_t)stream;
    DeviceCtx* c;
    ST_TRY(ctx_get(-1, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    CU_TRY(launch_keys_init(d_keys, m, st));
    CU_TRY(launch_index_build(k, n, d_refs_aos, d_index, d_index + INDEX_HEADER_FLOATS, true, st));
    // the tensor section only when this search is planned onto the tcgen05 path (as in search_host_on)
    Plan whole{};
    if (n > 0) ST_TRY(make_plan(k, m, n, flags, c->num_sms, nullptr, nullptr, &whole));
    if (whole.path == 2)
        CU_TRY(tensor_index_build(k, n, d_refs_aos, d_index + INDEX_HEADER_FLOATS + (size_t)ceil_div(n, LB) * index_block_floats(k), st));
    g_launches.fetch_add(n > 0 ? 3 : 2, std::memory_order_relaxed);  // + the unpack below
    ST_TRY(search_keys_on(c, k, m, n, d_queries, d_index, d_index + INDEX_HEADER_FLOATS, 0, d_keys, flags, st));
    CU_TRY(launch_keys_unpack(d_keys, m, d_idx, nullptr, st));
    return NNS_B200_OK;
}

int nns_b200_search_host(int k, int m, int n, const float* s_points, const float* r_points, int* results)
{
    ST_TRY(check_host_args(k, m, n, s_points, r_points, results));
    if (m == 0) return NNS_B200_OK;
    DeviceCtx* c;
    ST_TRY(ctx_get(-1, &c));
    return search_host_on(c, k, m, n, s_points, r_points, 0, nullptr, results);
}

int nns_b200_search_host_dist(int k, int m, int n, const float* s_points, const float* r_points, int* results,
                              float* distances)
{
    ST_TRY(check_host_args(k, m, n, s_points, r_points, results));
    if (m > 0 && !distances) return fail(NNS_B200_ERR_INVALID, "NULL array");
    if (m == 0) return NNS_B200_OK;
    DeviceCtx* c;
    ST_TRY(ctx_get(-1, &c));
    return search_host_on(c, k, m, n, s_points, r_points, 0, nullptr, results, nullptr, distances);
}

int nns_b200_search_multi(int k, int m, int n, const float* s_points, const float* r_points, int* results,
                          int num_gpus, int shard_mode)
{
    ST_TRY(check_host_args(k, m, n, s_points, r_points, results));
    if (shard_mode != 0 && shard_mode != 1) return fail(NNS_B200_ERR_INVALID, "shard_mode must be 0 or 1");
    if (m == 0) return NNS_B200_OK;
    int visible = 0;
    CU_TRY(cudaGetDeviceCount(&visible));
    if (num_gpus <= 0 || num_gpus > visible) num_gpus = visible;
    if (num_gpus <= 0) return fail(NNS_B200_ERR_CUDA, "no CUDA device");
    const int G = num_gpus;
    std::vector<DeviceCtx*> ctx(G);
    for (int g = 0; g < G; ++g) ST_TRY(ctx_get(g, &ctx[g]));

    std::vector<int> status(G, NNS_B200_OK);
    std::vector<std::string> msgs(G);
    std::vector<std::thread> th;
    if (shard_mode == 0) {
        // query-sharded: GPU g answers queries [g*per, ...) against every reference point
        const int per = ceil_div(m, G);
        for (int g = 0; g < G; ++g) {
            th.emplace_back([&, g]() {
                const int q0 = g * per;
                const int qn = q0 >= m ? 0 : ((m - q0) < per ? (m - q0) : per);
                if (qn > 0)
                    status[g] = search_host_on(ctx[g], k, qn, n, s_points + (size_t)q0 * k, r_points, 0, nullptr,
                                               results + q0);
                if (status[g] != NNS_B200_OK) msgs[g] = g_err;
            });
        }
        for (auto& t : th) t.join();
    } else {
        // reference-sharded: GPU g owns a contiguous slice of whole reference blocks
        // (core.cu:781-791 without the <= 0 tail defect D9); packed keys merged by integer MIN
        const long long blocks = ceil_div(n, LB);
        const long long per_blocks = (blocks + G - 1) / G;
        // Reduction over NVLink peer memory: when every GPU can address GPU 0's memory, each GPU
        // atomicMin's its packed keys straight into ONE key array resident on GPU 0
        // (keys_merge_kernel) -- no NCCL, no host merge.
        bool p2p = G > 1;
        for (int g = 1; g < G && p2p; ++g) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, g, 0) != cudaSuccess || !can) p2p = false;
        }
        if (p2p) {
            int prev = 0;
            CU_TRY(cudaGetDevice(&prev));
            for (int g = 1; g < G; ++g) {
                CU_TRY(cudaSetDevice(g));
                const cudaError_t pe = cudaDeviceEnablePeerAccess(0, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) p2p = false;
                cudaGetLastError();
            }
            CU_TRY(cudaSetDevice(prev));
        }
        if (p2p) {
            DeviceCtx* c0 = ctx[0];
            u64* shared_keys = nullptr;
            {
                std::lock_guard<std::mutex> lk(c0->mu);
                DeviceGuard guard;
                ST_TRY(guard.enter(0));
                ST_TRY(buf_reserve(&c0->peer_keys, (size_t)m * sizeof(u64)));
                shared_keys = (u64*)c0->peer_keys.p;
                CU_TRY(launch_keys_init(shared_keys, m, c0->compute));
                CU_TRY(cudaStreamSynchronize(c0->compute));
                g_launches.fetch_add(1, std::memory_order_relaxed);
            }
            for (int g = 0; g < G; ++g) {
                th.emplace_back([&, g]() {
                    const long long r0 = (long long)g * per_blocks * LB;
                    const long long rn = r0 >= n ? 0 : ((n - r0) < per_blocks * LB ? (n - r0) : per_blocks * LB);
                    if (rn <= 0) return;
                    status[g] = search_host_on(ctx[g], k, m, (int)rn, s_points, r_points + r0 * k, (int)r0, nullptr, nullptr,
                                               shared_keys);
                    if (status[g] != NNS_B200_OK) msgs[g] = g_err;
                });
            }
            for (auto& t : th) t.join();
            for (int g = 0; g < G; ++g)
                if (status[g] != NNS_B200_OK) return fail(status[g], "gpu %d: %s", g, msgs[g].c_str());
            std::lock_guard<std::mutex> lk(c0->mu);
            DeviceGuard guard;
            ST_TRY(guard.enter(0));
            ST_TRY(buf_reserve(&c0->idx, (size_t)m * sizeof(int)));
            CU_TRY(launch_keys_unpack(shared_keys, m, (int*)c0->idx.p, nullptr, c0->compute));
            CU_TRY(cudaMemcpyAsync(results, c0->idx.p, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, c0->compute));
            CU_TRY(cudaStreamSynchronize(c0->compute));
            g_launches.fetch_add(1, std::memory_order_relaxed);
            return NNS_B200_OK;
        }
        // no peer access: per-GPU keys to the host, merged there
        std::vector<std::vector<u64>> keys(G);
        for (int g = 0; g < G; ++g) {
            th.emplace_back([&, g]() {
                const long long r0 = (long long)g * per_blocks * LB;
                const long long rn = r0 >= n ? 0 : ((n - r0) < per_blocks * LB ? (n - r0) : per_blocks * LB);
                if (rn <= 0) return;
                keys[g].resize(m);
                status[g] = search_host_on(ctx[g], k, m, (int)rn, s_points, r_points + r0 * k, (int)r0,
                                           keys[g].data(), nullptr);
                if (status[g] != NNS_B200_OK) msgs[g] = g_err;
            });
        }
        for (auto& t : th) t.join();
        for (int g = 0; g < G; ++g)
            if (status[g] != NNS_B200_OK) return fail(status[g], "gpu %d: %s", g, msgs[g].c_str());
        for (int i = 0; i < m; ++i) {
            u64 best = KEY_INIT;
            for (int g = 0; g < G; ++g)
                if (!keys[g].empty() && keys[g][i] < best) best = keys[g][i];
            results[i] = (int)(unsigned)(best & 0xffffffffull);
        }
    }
    for (int g = 0; g < G; ++g)
        if (status[g] != NNS_B200_OK) return fail(status[g], "gpu %d: %s", g, msgs[g].c_str());
    return NNS_B200_OK;
}

void nns_b200_cudaCall(int k, int m, int n, float* s_points, float* r_points, int** results)
{
    // core.cu:31 -- the callee allocates, the caller frees
    int* out = (int*)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
    const int st = out ? nns_b200_search_host(k, m, n, s_points, r_points, out)
                       : fail(NNS_B200_ERR_NOMEM, "malloc(%zu) failed", sizeof(int) * (size_t)m);
    if (st != NNS_B200_OK) {
        // utils.h:16-26 -- the reference's CHECK prints and exits; there is no status to return
        if (st == NNS_B200_ERR_CUDA || g_err_line)
            printf("Error: %s:%d, code:%d, reason: %s \n", g_err_file, g_err_line, g_err_code,
                   cudaGetErrorString((cudaError_t)g_err_code));
        else
            printf("Error: nns_b200: %s \n", g_err);
        exit(1);
    }
    *results = out;
}

int nns_b200_tensor_stats(unsigned* out3)
{
    if (!out3) return fail(NNS_B200_ERR_INVALID, "NULL out");
    DeviceCtx* c;
    ST_TRY(ctx_get(-1, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    out3[0] = out3[1] = out3[2] = 0;
    if (!c->stats.p) return NNS_B200_OK;
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(out3, c->stats.p, 3 * sizeof(unsigned), cudaMemcpyDeviceToHost));
    return NNS_B200_OK;
}

int nns_b200_plan(int k, int m, int n, unsigned flags, int num_sms, int* plan)
{
    if (!plan) return fail(NNS_B200_ERR_INVALID, "NULL plan");
    Plan p;
    ST_TRY(make_plan(k, m, n, flags, num_sms, nullptr, nullptr, &p));
    plan[0] = p.path; plan[1] = p.q; plan[2] = p.warps; plan[3] = p.stages;
    plan[4] = p.nqb; plan[5] = p.splits; plan[6] = p.bps; plan[7] = p.smem;
    return NNS_B200_OK;
}

}  // extern "C"
