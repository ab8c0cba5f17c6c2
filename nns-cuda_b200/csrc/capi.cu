// capi.cu -- the C ABI of include/nns_b200.h: planning, per-device state, host-pointer
// ingest, single-process multi-GPU fan-out.  All compute is in the CUDA kernels of this
// library; there is no CPU search path anywhere in this file.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "host_state.h"

namespace nns {

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local char g_err_file[128] = "";
static thread_local int g_err_line = 0;
static thread_local int g_err_code = 0;

int fail(int status, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

int fail_cuda(cudaError_t e, const char* file, int line)
{
    snprintf(g_err_file, sizeof(g_err_file), "%s", file);
    g_err_line = line;
    g_err_code = (int)e;
    snprintf(g_err, sizeof(g_err), "%s:%d, code:%d, reason: %s", file, line, (int)e,
             cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? NNS_B200_ERR_NOMEM : NNS_B200_ERR_CUDA;
}

const char* last_error_text() { return g_err; }
void last_cuda_error(const char** file, int* line, int* code)
{
    *file = g_err_file;
    *line = g_err_line;
    *code = g_err_code;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launches(unsigned long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------------
// planning (pure host logic)
// ---------------------------------------------------------------------------------------------
// k <= 32: the split-precision tcgen05 screen (csrc/tensor_search.cu) beats the FP32 screened kernel
// once the problem amortises its fixed cost (query image, four launches, per-CTA TMEM set-up):
// B200, profiles/r1_tensor_*: C2 (k = 3, 2.7e11 pairs) 18.6 ms vs 39.9 ms, C3 (k = 16) 0.49 s vs 2.6 s,
// C1 (6.7e7 pairs) 0.38 ms vs 0.064 ms; the reference's largest shape (k = 16, m = 1024, n = 2^20,
// 1.1e9 pairs) 8.1 ms vs 6.7 ms end to end.  Both paths return identical indices.
// The crossover is a measured property of this build on B200, not of the algorithm: NNS_B200_TENSOR_MIN_PAIRS
// (a floating-point pair count) overrides it for other clocks / drivers without a rebuild.
static bool lowk_prefers_tensor(int k, int m, int n)
{
    static const double forced = []() {
        const char* e = getenv("NNS_B200_TENSOR_MIN_PAIRS");
        return e ? atof(e) : -1.0;
    }();
    const double pairs = (double)m * (double)n;
    return m >= 1024 && pairs >= (forced >= 0.0 ? forced : (k <= 8 ? 4e9 : 2e9));
}

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

int make_plan(int k, int m, int n, unsigned flags, int num_sms, occ_fn occ, void* occ_user, Plan* p)
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
    // EXACT_FORM asks for V0's formulation on every pair: never the screen (k > 32: the wide kernel);
    // k > 32: the screen needs whole 256-query strips and enough pairs to amortise its fixed cost
    const bool tensor_auto = !(flags & NNS_B200_FLAG_EXACT_FORM) &&
                             (k > LOWK_MAX_K ? (m >= 256 && (double)m * (double)n >= 1e7) : lowk_prefers_tensor(k, m, n));
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
static std::mutex g_ctx_mu;
static std::map<int, DeviceCtx*> g_ctx;

int ctx_get(int device, DeviceCtx** out)
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
        DeviceGuard guard;
        ST_TRY(guard.enter(device));
        cudaDeviceProp prop;
        CU_TRY(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            return fail(NNS_B200_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only",
                        device, prop.major, prop.minor);
        c->num_sms = prop.multiProcessorCount;
        // The tensor path takes its scratch (query image, candidate records) from a PRIVATE
        // stream-ordered pool: freed blocks stay in the pool across synchronisations (returning ~1 GB
        // to the driver after every host-pointer call cost 8 ms of a 28 ms call), but only up to
        // NNS_B200_POOL_KEEP_MB (default 4096), and the process-wide default pool -- which torch and
        // other co-tenants use -- is left alone.
        cudaMemPoolProps pp{};
        pp.allocType = cudaMemAllocationTypePinned;
        pp.handleTypes = cudaMemHandleTypeNone;
        pp.location.type = cudaMemLocationTypeDevice;
        pp.location.id = device;
        CU_TRY(cudaMemPoolCreate(&c->pool, &pp));
        const char* e = getenv("NNS_B200_POOL_KEEP_MB");
        unsigned long long keep = (e ? strtoull(e, nullptr, 0) : 4096ull) << 20;
        CU_TRY(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep));
        CU_TRY(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
        c->ready = true;
    }
    *out = c;
    return NNS_B200_OK;
}

int buf_reserve(DevBuf* b, size_t bytes)
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

int ctx_events(DeviceCtx* c, int count)
{
    while ((int)c->events.size() < count) {
        cudaEvent_t ev;
        CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        c->events.push_back(ev);
    }
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

bool plan_wants_tensor(int k, int m, int n, unsigned flags, int num_sms)
{
    if (m <= 0 || n <= 0) return false;
    Plan p{};
    if (make_plan(k, m, n, flags, num_sms, nullptr, nullptr, &p) != NNS_B200_OK) return false;
    return p.path == 2 && tensor_section_floats(k, n) != 0;
}

int search_keys_on(DeviceCtx* c, int k, int m, int n, const float* d_queries, const float* d_header,
                   const float* d_blocks, const float* d_section, int index_base, u64* d_keys, unsigned flags,
                   cudaStream_t st)
{
    if (m == 0 || n == 0) return NNS_B200_OK;
    if (!d_section) {
        if (flags & NNS_B200_FLAG_FORCE_TENSOR) return fail(NNS_B200_ERR_INVALID, "this index has no tensor section");
        flags |= (k <= LOWK_MAX_K && m >= 16) ? NNS_B200_FLAG_FORCE_LOWK : NNS_B200_FLAG_FORCE_WIDE;
    }
    Plan p;
    ST_TRY(make_plan(k, m, n, flags, c->num_sms, occ_query, c, &p));
    const int mode = lowk_mode_of(flags);
    const int nblocks = ceil_div(n, LB);
    if (p.path == 2) {
        int launches = 0;
        ST_TRY(buf_reserve(&c->stats, 64));
        // per-call status words {candidates, overflow flag, capacity} from the stream-ordered pool: the
        // flag gates the fallback kernel below, so it must not be shared with a search that another
        // stream of this device has in flight
        unsigned* d_status = nullptr;
        CU_TRY(cudaMallocFromPoolAsync((void**)&d_status, 64, c->pool, st));
        cudaError_t te = tensor_search(k, m, n, d_queries, d_blocks, d_section, index_base, d_keys, mode == LOWK_EXACT_V0,
                                       c->num_sms, st, c->pool, &launches, d_status,
                                       (flags & NNS_B200_FLAG_TEST_TINY_CANDIDATES) != 0);
        count_launches((unsigned long long)launches + 1);
        // Fallback for data whose near-ties overflow the candidate buffer (e.g. all points identical):
        // the FP32 kernel of this shape, launched unconditionally behind the device-side flag the
        // re-score kernel leaves in d_status[1]; it exits at once when the flag is clear.  No host
        // round trip.
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
            if (te == cudaSuccess) te = cudaMemcpyAsync(c->stats.p, d_status, 4 * sizeof(unsigned), cudaMemcpyDeviceToDevice, st);
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
    count_launches(1);
    return NNS_B200_OK;
}

// K nearest neighbours: the tcgen05 screen (tensor_topk_search) for large problems, else the FP32 kernel.  The
// screen's fixed cost (sample pass, image, lists) pays off from ~4e9 pairs; FORCE_TENSOR / FORCE_LOWK / FORCE_WIDE
// / EXACT_FORM select explicitly (the last three: the FP32 kernel).
bool topk_wants_tensor(int k, int m, int n, unsigned flags)
{
    if (k > TENSOR_MAX_K || n < 64 * LB) return false;
    if (flags & (NNS_B200_FLAG_FORCE_LOWK | NNS_B200_FLAG_FORCE_WIDE | NNS_B200_FLAG_EXACT_FORM)) return false;
    if (flags & NNS_B200_FLAG_FORCE_TENSOR) return true;
    return m >= 1024 && (double)m * (double)n >= 4e9;
}

int topk_keys_on(DeviceCtx* c, int k, int m, int n, int K, const float* d_queries, const float* d_blocks, const float* d_section,
                 int index_base, u64* d_keys, unsigned flags, cudaStream_t st)
{
    const bool exact = (flags & NNS_B200_FLAG_V0_ROUNDING) != 0;
    if (d_section && topk_wants_tensor(k, m, n, flags)) {
        ST_TRY(buf_reserve(&c->stats, 64));
        unsigned* d_status = nullptr;
        CU_TRY(cudaMallocFromPoolAsync((void**)&d_status, 64, c->pool, st));
        int launches = 0;
        cudaError_t e = tensor_topk_search(k, m, n, K, d_queries, d_blocks, d_section, index_base, d_keys, exact, c->num_sms, st, c->pool,
                                           &launches, d_status);
        count_launches(launches);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->stats.p, d_status, 4 * sizeof(unsigned), cudaMemcpyDeviceToDevice, st);
        const cudaError_t fe = cudaFreeAsync(d_status, st);
        CU_TRY(e);
        CU_TRY(fe);
        return NNS_B200_OK;
    }
    const int splits = topk_choose_splits(m, n, c->num_sms);
    u64* scratch = nullptr;
    CU_TRY(cudaMallocFromPoolAsync((void**)&scratch, topk_scratch_bytes(m, K, splits), c->pool, st));
    int launches = 0;
    const cudaError_t e = topk_search_launch(k, m, n, K, d_queries, d_blocks, index_base, d_keys, scratch, splits, exact, st, &launches);
    count_launches(launches);
    const cudaError_t fe = cudaFreeAsync(scratch, st);
    CU_TRY(e);
    CU_TRY(fe);
    return NNS_B200_OK;
}

// NNS_B200_FLAGS (environment, a C integer literal) applies the flags word to the host-pointer
// entry points, whose reference signature has no flags argument.
unsigned host_flags()
{
    static const unsigned f = []() {
        const char* e = getenv("NNS_B200_FLAGS");
        return e ? (unsigned)strtoul(e, nullptr, 0) : 0u;
    }();
    return f;
}

int check_host_args(int k, int m, int n, const void* s, const void* r, const void* out)
{
    if (k <= 0 || m < 0 || n < 0) return fail(NNS_B200_ERR_INVALID, "invalid shape k=%d m=%d n=%d", k, m, n);
    if ((long long)k * m > 0x7fffffffLL * 4 || (long long)k * n > 0x7fffffffLL * 4)
        return fail(NNS_B200_ERR_INVALID, "shape too large");
    if ((m > 0 && (!s || !out)) || (n > 0 && m > 0 && !r)) return fail(NNS_B200_ERR_INVALID, "NULL array");
    return NNS_B200_OK;
}

// used by nns_b200_shutdown
static void ctx_release_all()
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
            for (DevBuf* b : {&c->q, &c->r, &c->index, &c->tsec, &c->keys, &c->idx, &c->stats, &c->peer_keys}) {
                if (b->p) cudaFree(b->p);
                b->p = nullptr;
                b->cap = 0;
            }
            staging_release(c);
            for (cudaEvent_t ev : c->events) cudaEventDestroy(ev);
            c->events.clear();
            cudaStreamDestroy(c->compute);
            cudaStreamDestroy(c->copy);
            if (c->pool) cudaMemPoolDestroy(c->pool);
            c->pool = nullptr;
            cudaGetLastError();
            c->ready = false;
            c->occ_cache.clear();
        }
    }
    if (prev >= 0) cudaSetDevice(prev);
}

}  // namespace nns

using namespace nns;

// ---------------------------------------------------------------------------------------------
// exported: lifetime, planning, device-resident building blocks
// (host-pointer entry points and index handles: ingest.cu; multi-GPU: multi.cu)
// ---------------------------------------------------------------------------------------------
extern "C" {

int nns_b200_version(void) { return NNS_B200_VERSION; }
const char* nns_b200_last_error(void) { return last_error_text(); }
unsigned long long nns_b200_launch_count(void) { return g_launches.load(); }

int nns_b200_init(int device)
{
    DeviceCtx* c;
    return ctx_get(device, &c);
}

int nns_b200_shutdown(void)
{
    ctx_release_all();
    return NNS_B200_OK;
}

int nns_b200_device_sms(int device)
{
    DeviceCtx* c;
    if (ctx_get(device, &c) != NNS_B200_OK) return -1;
    return c->num_sms;
}

size_t nns_b200_index_floats(int k, int n)
{
    if (k <= 0 || n <= 0) return 0;
    return (size_t)INDEX_HEADER_FLOATS + (size_t)ceil_div(n, LB) * (size_t)index_block_floats(k) +
           tensor_section_floats(k, n);
}

static float* section_of(int k, int n, float* d_index)
{
    return tensor_section_floats(k, n) ? d_index + INDEX_HEADER_FLOATS + (size_t)ceil_div(n, LB) * index_block_floats(k) : nullptr;
}

int nns_b200_index_build(int k, int n, const float* d_refs_aos, float* d_index, void* stream)
{
    if (k <= 0 || n < 0) return fail(NNS_B200_ERR_INVALID, "invalid shape k=%d n=%d", k, n);
    if (n > 0 && (!d_refs_aos || !d_index)) return fail(NNS_B200_ERR_INVALID, "NULL array");
    if (((uintptr_t)d_index & 15) != 0) return fail(NNS_B200_ERR_INVALID, "index must be 16-byte aligned");
    if (n == 0) return NNS_B200_OK;
    float* d_blocks = d_index + INDEX_HEADER_FLOATS;
    CU_TRY(launch_index_build(k, n, d_refs_aos, d_index, d_blocks, true, (cudaStream_t)stream));
    CU_TRY(tensor_index_build(k, n, d_index, d_blocks, section_of(k, n, d_index), (cudaStream_t)stream));
    count_launches(tensor_section_floats(k, n) ? 4 : 1);
    return NNS_B200_OK;
}

/* One slice of an index that several GPUs (processes) build together: references [j0, j0 + cn) of an
 * index laid out for n_total references.  The centre of the tensor section is fixed by the caller so
 * that every slice uses the same one; the maxima of the slice accumulate in the header words of
 * `part`.  After the slices have been exchanged (all-gather of the block / image ranges and of the
 * header words), nns_b200_index_finish folds the partial maxima. */
int nns_b200_index_build_part(int k, int n_total, int j0, int cn, int part_blocks, const float* d_refs_aos_part,
                              float* d_index, const float* centre, int part, void* stream)
{
    if (k <= 0 || n_total <= 0 || j0 < 0 || cn < 0 || (j0 % LB) != 0 || part_blocks < ceil_div(cn, LB) ||
        ((long long)j0 / LB + part_blocks) > ceil_div(n_total, LB))
        return fail(NNS_B200_ERR_INVALID, "invalid part k=%d n=%d j0=%d cn=%d blocks=%d", k, n_total, j0, cn, part_blocks);
    if (part < 0 || part >= MAX_PEERS) return fail(NNS_B200_ERR_INVALID, "part must be 0..%d", MAX_PEERS - 1);
    if (!d_index || (cn > 0 && !d_refs_aos_part)) return fail(NNS_B200_ERR_INVALID, "NULL array");
    cudaStream_t st = (cudaStream_t)stream;
    float* d_section = section_of(k, n_total, d_index);
    // this GPU's copy of both headers: zero, then the fixed centre
    TensorCentre c{};
    if (d_section) {
        if (!centre) return fail(NNS_B200_ERR_INVALID, "a part of an index with a tensor section needs a fixed centre");
        for (int t = 0; t < k && t < 512; ++t) c.c[t] = centre[t];
        CU_TRY(tensor_section_init(k, n_total, nullptr, d_section, &c, st));
    }
    BlockDsts dst{};
    float* d_blocks_part = d_index + INDEX_HEADER_FLOATS + (size_t)(j0 / LB) * index_block_floats(k);
    dst.p[0] = d_blocks_part;
    dst.count = 1;
    CU_TRY(launch_index_build_to(k, cn, d_refs_aos_part, d_index, HDR_PART_MAX + part, dst, true, st, part_blocks));
    if (d_section) {
        ImageDsts idst{};
        idst.p[0] = reinterpret_cast<unsigned char*>(d_section + TENSOR_HDR_FLOATS) + (size_t)(j0 / LB) * tensor_image_bytes_per_block(k);
        idst.count = 1;
        CU_TRY(tensor_image_build(k, cn, d_blocks_part, d_section, THDR_PART_MAX + part, THDR_PART_FLAGS + part, idst, st, part_blocks));
    }
    count_launches(d_section ? 3 : 1);
    return NNS_B200_OK;
}

int nns_b200_index_finish(int k, int n_total, float* d_index, int parts, void* stream)
{
    if (k <= 0 || n_total <= 0 || !d_index || parts < 1 || parts > MAX_PEERS) return fail(NNS_B200_ERR_INVALID, "invalid index");
    CU_TRY(launch_header_fold(d_index, section_of(k, n_total, d_index), parts, (cudaStream_t)stream));
    count_launches(1);
    return NNS_B200_OK;
}

/* byte ranges of an index that a part owns (for the exchange between GPUs): out6 = { blocks offset,
 * blocks bytes, image offset, image bytes (0 without a tensor section), index-header partial word
 * offset, section-header partial-max word offset } -- all in bytes from d_index */
int nns_b200_index_part_ranges(int k, int n_total, int j0, int part_blocks, int part, size_t* out6)
{
    if (k <= 0 || n_total <= 0 || !out6 || (j0 % LB) != 0) return fail(NNS_B200_ERR_INVALID, "invalid part");
    const size_t nb = (size_t)part_blocks, b0 = (size_t)j0 / LB;
    out6[0] = ((size_t)INDEX_HEADER_FLOATS + b0 * index_block_floats(k)) * 4;
    out6[1] = nb * index_block_floats(k) * 4;
    const bool sec = tensor_section_floats(k, n_total) != 0;
    const size_t sec0 = ((size_t)INDEX_HEADER_FLOATS + (size_t)ceil_div(n_total, LB) * index_block_floats(k)) * 4;
    out6[2] = sec ? sec0 + TENSOR_HDR_FLOATS * 4 + b0 * tensor_image_bytes_per_block(k) : 0;
    out6[3] = sec ? nb * tensor_image_bytes_per_block(k) : 0;
    out6[4] = (size_t)(HDR_PART_MAX + part) * 4;
    out6[5] = sec ? sec0 + (size_t)(THDR_PART_MAX + part) * 4 : 0;
    return NNS_B200_OK;
}

int nns_b200_keys_init(uint64_t* d_keys, int m, void* stream)
{
    if (m < 0 || (m > 0 && !d_keys)) return fail(NNS_B200_ERR_INVALID, "invalid keys");
    CU_TRY(launch_keys_init((u64*)d_keys, m, (cudaStream_t)stream));
    count_launches(m > 0 ? 1 : 0);
    return NNS_B200_OK;
}

int nns_b200_keys_unpack(const uint64_t* d_keys, int m, int* d_idx, float* d_dist, void* stream)
{
    if (m < 0 || (m > 0 && (!d_keys || !d_idx))) return fail(NNS_B200_ERR_INVALID, "invalid keys");
    CU_TRY(launch_keys_unpack((const u64*)d_keys, m, d_idx, d_dist, (cudaStream_t)stream));
    count_launches(m > 0 ? 1 : 0);
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
    return search_keys_on(c, k, m, n, d_queries, d_index, d_index + INDEX_HEADER_FLOATS, section_of(k, n, (float*)d_index),
                          index_base, (u64*)d_keys, flags, (cudaStream_t)stream);
}

int nns_b200_topk_keys(int k, int m, int n, int K, const float* d_queries, const float* d_index, int index_base,
                       uint64_t* d_keys, unsigned flags, void* stream)
{
    if (k <= 0 || k > TOPK_MAX_DIMS || m < 0 || n < 0) return fail(NNS_B200_ERR_INVALID, "invalid shape k=%d m=%d n=%d", k, m, n);
    if (K < 1 || K > TOPK_MAX_K) return fail(NNS_B200_ERR_UNSUPPORTED, "K must be 1..%d", TOPK_MAX_K);
    if (m > 0 && n > 0 && (!d_queries || !d_index || !d_keys)) return fail(NNS_B200_ERR_INVALID, "NULL array");
    if ((long long)index_base + n > 0x7fffffffLL) return fail(NNS_B200_ERR_INVALID, "index_base + n overflows int32");
    if (m == 0 || n == 0) return NNS_B200_OK;
    DeviceCtx* c;
    ST_TRY(ctx_get(-1, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    return topk_keys_on(c, k, m, n, K, d_queries, d_index + INDEX_HEADER_FLOATS, section_of(k, n, (float*)d_index), index_base,
                        (u64*)d_keys, flags, (cudaStream_t)stream);
}

int nns_b200_topk_unpack(const uint64_t* d_keys, int m, int K, int* d_idx, float* d_dist, void* stream)
{
    if (m < 0 || K < 1 || K > TOPK_MAX_K || (m > 0 && (!d_keys || !d_idx))) return fail(NNS_B200_ERR_INVALID, "invalid keys");
    CU_TRY(topk_unpack_launch((const u64*)d_keys, m, K, d_idx, d_dist, (cudaStream_t)stream));
    count_launches(m > 0 ? 1 : 0);
    return NNS_B200_OK;
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
    float* d_blocks = d_index + INDEX_HEADER_FLOATS;
    CU_TRY(launch_index_build(k, n, d_refs_aos, d_index, d_blocks, true, st));
    // the tensor section only when this search is planned onto the tcgen05 path
    float* d_section = nullptr;
    if (plan_wants_tensor(k, m, n, flags, c->num_sms)) {
        d_section = section_of(k, n, d_index);
        CU_TRY(tensor_index_build(k, n, d_index, d_blocks, d_section, st));
    }
    count_launches(n > 0 ? 3 : 2);  // + the unpack below
    ST_TRY(search_keys_on(c, k, m, n, d_queries, d_index, d_blocks, d_section, 0, d_keys, flags, st));
    CU_TRY(launch_keys_unpack(d_keys, m, d_idx, nullptr, st));
    return NNS_B200_OK;
}

int nns_b200_tensor_stats(unsigned* out4)
{
    unsigned* out3 = out4;
    if (!out3) return fail(NNS_B200_ERR_INVALID, "NULL out");
    DeviceCtx* c;
    ST_TRY(ctx_get(-1, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    out3[0] = out3[1] = out3[2] = out3[3] = 0;
    if (!c->stats.p) return NNS_B200_OK;
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(out3, c->stats.p, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost));
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
