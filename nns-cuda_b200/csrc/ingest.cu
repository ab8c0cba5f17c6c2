// ingest.cu -- the host-pointer side of the boundary: the drop-in symbol, pageable-memory staging,
// chunked upload overlapped with index build and search, and the device-resident index handle
// ("build once, query many").  Replaces the per-call cudaMalloc + thrust H2D + transpose of
// core.cu:351-370 (and its copies in every later variant).  All compute is in the CUDA kernels.
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <thread>
#include <time.h>

#include "host_state.h"

namespace nns {

// ---------------------------------------------------------------------------------------------
// parallel host memcpy (pageable -> pinned staging)
// ---------------------------------------------------------------------------------------------
// A small persistent pool: a single thread's memcpy (~10 GB/s) is what limits the driver's own
// pageable path; several threads reach the PCIe rate.  NNS_B200_COPY_THREADS overrides the size
// (0 = the calling thread only).
class CopyPool {
public:
    static CopyPool& get()
    {
        static CopyPool* p = new CopyPool();  // leaked on purpose: worker threads outlive static destructors
        return *p;
    }
    void copy(void* dst, const void* src, size_t bytes)
    {
        const size_t min_piece = (size_t)512 << 10;
        int parts = (int)std::min<size_t>((size_t)workers_.size() + 1, (bytes + min_piece - 1) / min_piece);
        if (parts <= 1) {
            memcpy(dst, src, bytes);
            return;
        }
        const size_t piece = ((bytes + parts - 1) / parts + 4095) & ~(size_t)4095;
        std::atomic<int> pending{0};
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (int i = 1; i < parts; ++i) {
                const size_t off = (size_t)i * piece;
                if (off >= bytes) break;
                jobs_.push_back(Job{(char*)dst + off, (const char*)src + off, std::min(piece, bytes - off), &pending});
                pending.fetch_add(1, std::memory_order_relaxed);
            }
        }
        cv_.notify_all();
        memcpy(dst, src, std::min(piece, bytes));
        // help with whatever is still queued (possibly other callers' pieces), then wait for ours
        for (;;) {
            Job j;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (jobs_.empty()) break;
                j = jobs_.front();
                jobs_.pop_front();
            }
            run(j);
        }
        while (pending.load(std::memory_order_acquire) != 0) std::this_thread::yield();
    }

private:
    struct Job {
        char* dst;
        const char* src;
        size_t bytes;
        std::atomic<int>* pending;
    };
    static void run(const Job& j)
    {
        memcpy(j.dst, j.src, j.bytes);
        j.pending->fetch_sub(1, std::memory_order_release);
    }
    CopyPool()
    {
        const char* e = getenv("NNS_B200_COPY_THREADS");
        int n = e ? atoi(e) : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2)) - 1;
        if (n < 0) n = 0;
        if (n > 32) n = 32;
        for (int i = 0; i < n; ++i) workers_.emplace_back([this]() { loop(); });
        for (auto& t : workers_) t.detach();
    }
    void loop()
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this]() { return !jobs_.empty(); });
                j = jobs_.front();
                jobs_.pop_front();
            }
            run(j);
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Job> jobs_;
    std::vector<std::thread> workers_;
};

// ---------------------------------------------------------------------------------------------
// staged upload
// ---------------------------------------------------------------------------------------------
static bool host_pointer_is_pinned(const void* p)
{
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static int staging_reserve(DeviceCtx* c)
{
    Staging& sg = c->stage;
    if (sg.slot[0]) return NNS_B200_OK;
    const char* e = getenv("NNS_B200_STAGE_MB");
    sg.slot_bytes = (size_t)(e ? std::max(1, atoi(e)) : 8) << 20;
    for (int i = 0; i < STAGE_SLOTS; ++i) {
        CU_TRY(cudaHostAlloc((void**)&sg.slot[i], sg.slot_bytes, cudaHostAllocDefault));
        CU_TRY(cudaEventCreateWithFlags(&sg.drained[i], cudaEventDisableTiming));
        sg.in_flight[i] = false;
    }
    return NNS_B200_OK;
}

void staging_release(DeviceCtx* c)
{
    Staging& sg = c->stage;
    for (int i = 0; i < STAGE_SLOTS; ++i) {
        if (sg.slot[i]) cudaFreeHost(sg.slot[i]);
        if (sg.drained[i]) cudaEventDestroy(sg.drained[i]);
        sg.slot[i] = nullptr;
        sg.drained[i] = nullptr;
        sg.in_flight[i] = false;
    }
}

int h2d_async(DeviceCtx* c, void* d_dst, const void* h_src, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return NNS_B200_OK;
    // pinned (or tiny) sources go straight to the copy engine
    if (bytes <= ((size_t)64 << 10) || host_pointer_is_pinned(h_src)) {
        CU_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        return NNS_B200_OK;
    }
    ST_TRY(staging_reserve(c));
    Staging& sg = c->stage;
    for (size_t off = 0; off < bytes; off += sg.slot_bytes) {
        const size_t piece = std::min(sg.slot_bytes, bytes - off);
        const int i = sg.next;
        sg.next = (sg.next + 1) % STAGE_SLOTS;
        if (sg.in_flight[i]) CU_TRY(cudaEventSynchronize(sg.drained[i]));  // the DMA out of this slot has finished
        CopyPool::get().copy(sg.slot[i], (const char*)h_src + off, piece);
        CU_TRY(cudaMemcpyAsync((char*)d_dst + off, sg.slot[i], piece, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaEventRecord(sg.drained[i], st));
        sg.in_flight[i] = true;
    }
    return NNS_B200_OK;
}

// mean of at most 4096 strided sample rows: any centre is valid for the tensor section, the sample
// mean is as good as the full one for the error bound and costs O(4096 k) on the host
void sample_centre_host(int k, int n, const float* r_points, TensorCentre* out)
{
    memset(out, 0, sizeof(*out));
    if (n <= 0 || k > 512) return;
    const int samples = n < 4096 ? n : 4096;
    double sum[512] = {0.0};
    for (int i = 0; i < samples; ++i) {
        const float* row = r_points + (size_t)((long long)i * n / samples) * k;
        for (int t = 0; t < k; ++t) {
            const float x = row[t];
            if (x >= -1e15f && x <= 1e15f) sum[t] += x;  // NaN / INF / huge coordinates do not steer the centre
        }
    }
    for (int t = 0; t < k; ++t) out->c[t] = (float)(sum[t] / samples);
}

// ---------------------------------------------------------------------------------------------
// one-shot search over host arrays
// ---------------------------------------------------------------------------------------------
// Reference chunk of the ingest pipeline, in points.  FP32 paths: ~32 MiB of AoS data.  tcgen05 path:
// every chunk is searched as an index of its own (own centre, own query image, own candidate seeds),
// which costs a fixed ~0.2 ms, so chunks are a quarter of the set but at least 24 MiB -- and for
// k > 32 the set is one chunk: every chunk re-scores its own near-minimum candidates exactly, which at
// k = 128 costs more than the upload it hides (B200, C4: 279 ms in four chunks vs 245 ms in one).
long long ingest_chunk_points(int k, int n, bool tensor)
{
    const long long row = (long long)k * 4;
    // NNS_B200_INGEST_CHUNK_MB (tuning): 0 = the whole set in one chunk, > 0 = that many MiB of AoS data per chunk
    static const long long forced = []() {
        const char* e = getenv("NNS_B200_INGEST_CHUNK_MB");
        return e ? atoll(e) : -1ll;
    }();
    if (forced == 0) return ((long long)n + LB - 1) / LB * LB > LB ? ((long long)n + LB - 1) / LB * LB : LB;
    if (forced > 0) {
        const long long c = ((forced << 20) / row) / LB * LB;
        return c < LB ? LB : c;
    }
    long long chunk = ((32ll << 20) / row) / LB * LB;
    if (tensor && k > LOWK_MAX_K) return ((long long)n + LB - 1) / LB * LB > LB ? ((long long)n + LB - 1) / LB * LB : LB;
    if (tensor) {
        const long long quarter = (((long long)n + 3) / 4 + LB - 1) / LB * LB;
        const long long floor24 = ((24ll << 20) / row) / LB * LB;
        chunk = quarter > floor24 ? quarter : floor24;
    }
    return chunk < LB ? LB : chunk;
}

// Host arrays -> device -> keys (h_keys != NULL) or indices (h_idx != NULL) on the host.
// References are ingested in chunks: the upload of chunk c+1 (copy stream; pageable sources through the
// pinned staging ring) overlaps the index build + search of chunk c (compute stream); every chunk
// accumulates into the same packed keys with its own index base.
int search_host_on(DeviceCtx* c, int k, int m, int n, const float* s, const float* r, int index_base,
                   u64* h_keys, int* h_idx, u64* ext_keys, float* h_dist)
{
    std::lock_guard<std::mutex> lk(c->mu);
    return search_host_locked(c, k, m, n, s, r, index_base, h_keys, h_idx, ext_keys, h_dist);
}

// NNS_B200_TRACE=1: host-clock phase times of every host-pointer search on stderr (tuning aid)
static bool ingest_trace()
{
    static const bool on = []() { const char* e = getenv("NNS_B200_TRACE"); return e && atoi(e) != 0; }();
    return on;
}
static double now_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int search_host_locked(DeviceCtx* c, int k, int m, int n, const float* s, const float* r, int index_base,
                       u64* h_keys, int* h_idx, u64* ext_keys, float* h_dist)
{
    const double t_start = now_ms();
    DeviceGuard guard;
    ST_TRY(guard.enter(c->device));
    const unsigned flags = host_flags();
    const bool tensor = plan_wants_tensor(k, m, n, flags, c->num_sms);
    const long long chunk = ingest_chunk_points(k, n, tensor);
    const int nchunks = n > 0 ? (int)((n + chunk - 1) / chunk) : 0;
    const size_t qbytes = (size_t)m * k * sizeof(float);
    const size_t rbytes = (size_t)n * k * sizeof(float);
    const size_t bf = index_block_floats(k);
    const size_t ibytes = ((size_t)INDEX_HEADER_FLOATS + (size_t)ceil_div(n, LB) * bf) * sizeof(float);
    // per-chunk tensor sections (header + images), chunk after chunk
    const size_t chunk_sec_floats = tensor ? tensor_section_floats(k, (int)std::min<long long>(chunk, n)) : 0;
    ST_TRY(buf_reserve(&c->q, qbytes));
    ST_TRY(buf_reserve(&c->r, rbytes));
    ST_TRY(buf_reserve(&c->index, ibytes));
    if (tensor) ST_TRY(buf_reserve(&c->tsec, chunk_sec_floats * sizeof(float) * (size_t)nchunks));
    ST_TRY(buf_reserve(&c->keys, (size_t)m * sizeof(u64)));
    ST_TRY(buf_reserve(&c->idx, (size_t)m * sizeof(int) * (h_dist ? 2 : 1)));
    float* d_q = (float*)c->q.p;
    float* d_r = (float*)c->r.p;
    float* d_index = (float*)c->index.p;
    u64* d_keys = (u64*)c->keys.p;
    int* d_idx = (int*)c->idx.p;
    ST_TRY(ctx_events(c, nchunks + 1));

    const double t_reserved = now_ms();
    // queries ride the copy stream ahead of the first reference chunk
    ST_TRY(h2d_async(c, d_q, s, qbytes, c->copy));
    CU_TRY(launch_keys_init(d_keys, m, c->compute));
    count_launches(h_idx ? 2 : 1);  // keys init (+ unpack below)
    for (int ci = 0; ci < nchunks; ++ci) {
        const long long j0 = (long long)ci * chunk;
        const int cn = (int)((n - j0) < chunk ? (n - j0) : chunk);
        const double tc0 = now_ms();
        ST_TRY(h2d_async(c, d_r + j0 * k, r + j0 * k, (size_t)cn * k * sizeof(float), c->copy));
        CU_TRY(cudaEventRecord(c->events[ci], c->copy));
        CU_TRY(cudaStreamWaitEvent(c->compute, c->events[ci], 0));
        float* d_blocks_c = d_index + INDEX_HEADER_FLOATS + (j0 / LB) * (long long)bf;
        CU_TRY(launch_index_build(k, cn, d_r + j0 * k, d_index, d_blocks_c, ci == 0, c->compute));
        count_launches(1);
        const double tc1 = now_ms();
        float* d_section_c = nullptr;
        if (tensor) {
            d_section_c = (float*)c->tsec.p + (size_t)ci * chunk_sec_floats;
            CU_TRY(tensor_index_build(k, cn, d_index, d_blocks_c, d_section_c, c->compute));
            count_launches(3);
        }
        const double tc2 = now_ms();
        ST_TRY(search_keys_on(c, k, m, cn, d_q, d_index, d_blocks_c, d_section_c, index_base + (int)j0, d_keys, flags,
                              c->compute));
        if (ingest_trace())
            fprintf(stderr, "nns_b200 trace:   chunk %d: upload+transpose enqueue %.2f ms, tensor build enqueue %.2f ms, search enqueue %.2f ms\n", ci,
                    tc1 - tc0, tc2 - tc1, now_ms() - tc2);
    }
    if (nchunks == 0) {  // n == 0: still wait for the query upload before the buffers are reused
        CU_TRY(cudaEventRecord(c->events[0], c->copy));
        CU_TRY(cudaStreamWaitEvent(c->compute, c->events[0], 0));
    }
    // ext_keys: an already initialised key array, possibly in a PEER GPU's memory (NVLink P2P).  The
    // search accumulated into this GPU's own keys; one merge kernel folds them into ext_keys with
    // system-scope atomics, i.e. the cross-GPU (dist, idx) reduction is m red.min.u64 over NVLink.
    if (ext_keys) {
        CU_TRY(launch_keys_merge(ext_keys, d_keys, m, c->compute));
        count_launches(1);
    }
    if (h_keys) CU_TRY(cudaMemcpyAsync(h_keys, d_keys, (size_t)m * sizeof(u64), cudaMemcpyDeviceToHost, c->compute));
    if (h_idx) {
        float* d_dist = h_dist ? (float*)(d_idx + m) : nullptr;  // second half of the idx buffer
        CU_TRY(launch_keys_unpack(d_keys, m, d_idx, d_dist, c->compute));
        CU_TRY(cudaMemcpyAsync(h_idx, d_idx, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, c->compute));
        if (h_dist) CU_TRY(cudaMemcpyAsync(h_dist, d_dist, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
    }
    const double t_enqueued = now_ms();
    const cudaError_t idle = ingest_trace() ? cudaStreamQuery(c->compute) : cudaSuccess;
    if (ingest_trace()) fprintf(stderr, "nns_b200 trace:   compute stream after the download returned: %s\n", idle == cudaSuccess ? "idle" : cudaGetErrorName(idle));
    cudaGetLastError();
    CU_TRY(stream_drain(c->compute));
    const double t_sync1 = now_ms();
    CU_TRY(stream_drain(c->copy));
    if (ingest_trace()) {
        unsigned long long res = 0, used = 0, dres = 0;
        cudaMemPoolGetAttribute(c->pool, cudaMemPoolAttrReservedMemCurrent, &res);
        cudaMemPoolGetAttribute(c->pool, cudaMemPoolAttrUsedMemCurrent, &used);
        cudaMemPool_t dp;
        if (cudaDeviceGetDefaultMemPool(&dp, c->device) == cudaSuccess) cudaMemPoolGetAttribute(dp, cudaMemPoolAttrReservedMemCurrent, &dres);
        fprintf(stderr, "nns_b200 trace: k=%d m=%d n=%d chunks=%d tensor=%d | reserve %.2f ms, enqueue + run %.2f ms, sync compute %.2f ms, sync copy %.2f ms"
                " | pool reserved %.0f MB used %.0f MB, default pool reserved %.0f MB\n", k, m, n, nchunks, (int)tensor, t_reserved - t_start,
                t_enqueued - t_reserved, t_sync1 - t_enqueued, now_ms() - t_sync1, res / 1048576.0, used / 1048576.0, dres / 1048576.0);
    }
    return NNS_B200_OK;
}

}  // namespace nns

using namespace nns;

// ---------------------------------------------------------------------------------------------
// index handle
// ---------------------------------------------------------------------------------------------
struct nns_b200_index {
    int k = 0, n = 0, device = 0;
    DeviceCtx* ctx = nullptr;
    float* d_index = nullptr;    // header + FP32 blocks
    float* d_section = nullptr;  // tensor section, built by the first search the planner puts on tcgen05
    bool has_section = false;
    std::mutex mu;
};

extern "C" {

int nns_b200_index_create(int k, int n, const float* r_points, int device, nns_b200_index_t** out)
{
    if (!out) return fail(NNS_B200_ERR_INVALID, "NULL out");
    *out = nullptr;
    if (k <= 0 || n < 0 || (n > 0 && !r_points)) return fail(NNS_B200_ERR_INVALID, "invalid index k=%d n=%d", k, n);
    if ((long long)k * n > 0x7fffffffLL * 4) return fail(NNS_B200_ERR_INVALID, "shape too large");
    DeviceCtx* c;
    ST_TRY(ctx_get(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard guard;
    ST_TRY(guard.enter(c->device));
    nns_b200_index* h = new nns_b200_index();
    h->k = k; h->n = n; h->device = c->device; h->ctx = c;
    const size_t bf = index_block_floats(k);
    const size_t ibytes = ((size_t)INDEX_HEADER_FLOATS + (size_t)ceil_div(n, LB) * bf) * sizeof(float);
    cudaError_t e = cudaMalloc((void**)&h->d_index, ibytes);
    if (e != cudaSuccess) { delete h; return fail_cuda(e, __FILE__, __LINE__); }
    // chunked ingest: upload (copy stream) overlapped with the transpose (compute stream)
    const long long chunk = ingest_chunk_points(k, n, false);
    const int nchunks = n > 0 ? (int)((n + chunk - 1) / chunk) : 0;
    int st = buf_reserve(&c->r, (size_t)n * k * sizeof(float));
    if (st == NNS_B200_OK) st = ctx_events(c, nchunks + 1);
    float* d_r = (float*)c->r.p;
    for (int ci = 0; ci < nchunks && st == NNS_B200_OK; ++ci) {
        const long long j0 = (long long)ci * chunk;
        const int cn = (int)((n - j0) < chunk ? (n - j0) : chunk);
        st = h2d_async(c, d_r + j0 * k, r_points + j0 * k, (size_t)cn * k * sizeof(float), c->copy);
        if (st != NNS_B200_OK) break;
        cudaError_t ce = cudaEventRecord(c->events[ci], c->copy);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(c->compute, c->events[ci], 0);
        if (ce == cudaSuccess)
            ce = launch_index_build(k, cn, d_r + j0 * k, h->d_index, h->d_index + INDEX_HEADER_FLOATS + (j0 / LB) * (long long)bf,
                                    ci == 0, c->compute);
        if (ce != cudaSuccess) st = fail_cuda(ce, __FILE__, __LINE__);
        count_launches(1);
    }
    if (st == NNS_B200_OK && n == 0) {
        cudaError_t ce = cudaMemsetAsync(h->d_index, 0, INDEX_HEADER_FLOATS * sizeof(float), c->compute);
        if (ce != cudaSuccess) st = fail_cuda(ce, __FILE__, __LINE__);
    }
    if (st == NNS_B200_OK) {
        cudaError_t ce = cudaStreamSynchronize(c->copy);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->compute);
        if (ce != cudaSuccess) st = fail_cuda(ce, __FILE__, __LINE__);
    }
    if (st != NNS_B200_OK) {
        cudaFree(h->d_index);
        delete h;
        return st;
    }
    *out = h;
    return NNS_B200_OK;
}

int nns_b200_index_search(nns_b200_index_t* h, int m, const float* s_points, int* results, float* distances)
{
    if (!h) return fail(NNS_B200_ERR_INVALID, "NULL index");
    if (m < 0 || (m > 0 && (!s_points || !results))) return fail(NNS_B200_ERR_INVALID, "invalid queries");
    if ((long long)h->k * m > 0x7fffffffLL * 4) return fail(NNS_B200_ERR_INVALID, "shape too large");
    if (m == 0) return NNS_B200_OK;
    DeviceCtx* c = h->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    std::lock_guard<std::mutex> lh(h->mu);
    DeviceGuard guard;
    ST_TRY(guard.enter(c->device));
    const int k = h->k, n = h->n;
    const unsigned flags = host_flags();
    ST_TRY(buf_reserve(&c->q, (size_t)m * k * sizeof(float)));
    ST_TRY(buf_reserve(&c->keys, (size_t)m * sizeof(u64)));
    ST_TRY(buf_reserve(&c->idx, (size_t)m * sizeof(int) * 2));
    ST_TRY(ctx_events(c, 1));
    float* d_q = (float*)c->q.p;
    u64* d_keys = (u64*)c->keys.p;
    int* d_idx = (int*)c->idx.p;
    float* d_dist = (float*)(d_idx + m);
    ST_TRY(h2d_async(c, d_q, s_points, (size_t)m * k * sizeof(float), c->copy));
    CU_TRY(cudaEventRecord(c->events[0], c->copy));
    CU_TRY(launch_keys_init(d_keys, m, c->compute));
    // the tensor section is added to the index by the first search that is planned onto tcgen05
    if (!h->has_section && plan_wants_tensor(k, m, n, flags, c->num_sms)) {
        if (!h->d_section) CU_TRY(cudaMalloc((void**)&h->d_section, tensor_section_floats(k, n) * sizeof(float)));
        CU_TRY(tensor_index_build(k, n, h->d_index, h->d_index + INDEX_HEADER_FLOATS, h->d_section, c->compute));
        h->has_section = true;
        count_launches(3);
    }
    CU_TRY(cudaStreamWaitEvent(c->compute, c->events[0], 0));
    ST_TRY(search_keys_on(c, k, m, n, d_q, h->d_index, h->d_index + INDEX_HEADER_FLOATS, h->has_section ? h->d_section : nullptr,
                          0, d_keys, flags, c->compute));
    CU_TRY(launch_keys_unpack(d_keys, m, d_idx, distances ? d_dist : nullptr, c->compute));
    count_launches(2);
    CU_TRY(cudaMemcpyAsync(results, d_idx, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, c->compute));
    if (distances) CU_TRY(cudaMemcpyAsync(distances, d_dist, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
    CU_TRY(stream_drain(c->compute));
    return NNS_B200_OK;
}

int nns_b200_index_size(const nns_b200_index_t* h, int* k, int* n)
{
    if (!h) return fail(NNS_B200_ERR_INVALID, "NULL index");
    if (k) *k = h->k;
    if (n) *n = h->n;
    return NNS_B200_OK;
}

int nns_b200_index_destroy(nns_b200_index_t* h)
{
    if (!h) return NNS_B200_OK;
    {
        std::lock_guard<std::mutex> lk(h->ctx->mu);
        DeviceGuard guard;
        if (guard.enter(h->device) == NNS_B200_OK) {
            cudaStreamSynchronize(h->ctx->compute);
            if (h->d_index) cudaFree(h->d_index);
            if (h->d_section) cudaFree(h->d_section);
        }
    }
    delete h;
    return NNS_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// one-shot host entry points
// ---------------------------------------------------------------------------------------------
int nns_b200_search_host(int k, int m, int n, const float* s_points, const float* r_points, int* results)
{
    ST_TRY(check_host_args(k, m, n, s_points, r_points, results));
    if (m == 0) return NNS_B200_OK;
    DeviceCtx* c;
    ST_TRY(ctx_get(-1, &c));
    return search_host_on(c, k, m, n, s_points, r_points, 0, nullptr, results, nullptr, nullptr);
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

/* K nearest neighbours over host arrays: chunked ingest like nns_b200_search_host, every chunk
 * accumulating into the same sorted key lists. */
int nns_b200_search_topk_host(int k, int m, int n, int K, const float* s_points, const float* r_points, int* indices,
                              float* distances)
{
    ST_TRY(check_host_args(k, m, n, s_points, r_points, indices));
    if (K < 1 || K > TOPK_MAX_K) return fail(NNS_B200_ERR_UNSUPPORTED, "K must be 1..%d", TOPK_MAX_K);
    if (k > TOPK_MAX_DIMS) return fail(NNS_B200_ERR_UNSUPPORTED, "top-K needs k <= %d", TOPK_MAX_DIMS);
    if (m == 0) return NNS_B200_OK;
    DeviceCtx* c;
    ST_TRY(ctx_get(-1, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard guard;
    ST_TRY(guard.enter(c->device));
    const unsigned flags = host_flags();
    // the tcgen05-screened search takes the whole set as one index (its sample pass fixes one threshold per query)
    const bool tensor = topk_wants_tensor(k, m, n, flags);
    const long long chunk = tensor ? std::max<long long>(LB, ((long long)n + LB - 1) / LB * LB) : ingest_chunk_points(k, n, false);
    const int nchunks = n > 0 ? (int)((n + chunk - 1) / chunk) : 0;
    const size_t bf = index_block_floats(k);
    const size_t cells = (size_t)m * K;
    ST_TRY(buf_reserve(&c->q, (size_t)m * k * sizeof(float)));
    ST_TRY(buf_reserve(&c->r, (size_t)n * k * sizeof(float)));
    ST_TRY(buf_reserve(&c->index, ((size_t)INDEX_HEADER_FLOATS + (size_t)ceil_div(n, LB) * bf) * sizeof(float)));
    if (tensor) ST_TRY(buf_reserve(&c->tsec, tensor_section_floats(k, n) * sizeof(float)));
    ST_TRY(buf_reserve(&c->keys, cells * sizeof(u64)));
    ST_TRY(buf_reserve(&c->idx, cells * sizeof(int) * 2));
    ST_TRY(ctx_events(c, nchunks + 1));
    float* d_q = (float*)c->q.p;
    float* d_r = (float*)c->r.p;
    float* d_index = (float*)c->index.p;
    u64* d_keys = (u64*)c->keys.p;
    int* d_idx = (int*)c->idx.p;
    float* d_dist = (float*)(d_idx + cells);
    ST_TRY(h2d_async(c, d_q, s_points, (size_t)m * k * sizeof(float), c->copy));
    if (cells > 0x7fffffffLL) return fail(NNS_B200_ERR_INVALID, "m * K too large");
    CU_TRY(launch_keys_init(d_keys, (int)cells, c->compute));
    count_launches(2);
    for (int ci = 0; ci < nchunks; ++ci) {
        const long long j0 = (long long)ci * chunk;
        const int cn = (int)((n - j0) < chunk ? (n - j0) : chunk);
        ST_TRY(h2d_async(c, d_r + j0 * k, r_points + j0 * k, (size_t)cn * k * sizeof(float), c->copy));
        CU_TRY(cudaEventRecord(c->events[ci], c->copy));
        CU_TRY(cudaStreamWaitEvent(c->compute, c->events[ci], 0));
        float* d_blocks_c = d_index + INDEX_HEADER_FLOATS + (j0 / LB) * (long long)bf;
        CU_TRY(launch_index_build(k, cn, d_r + j0 * k, d_index, d_blocks_c, ci == 0, c->compute));
        count_launches(1);
        float* d_section = nullptr;
        if (tensor) {
            d_section = (float*)c->tsec.p;
            CU_TRY(tensor_index_build(k, cn, d_index, d_blocks_c, d_section, c->compute));
            count_launches(3);
        }
        ST_TRY(topk_keys_on(c, k, m, cn, K, d_q, d_blocks_c, d_section, (int)j0, d_keys, flags, c->compute));
    }
    if (nchunks == 0) {
        CU_TRY(cudaEventRecord(c->events[0], c->copy));
        CU_TRY(cudaStreamWaitEvent(c->compute, c->events[0], 0));
    }
    CU_TRY(topk_unpack_launch(d_keys, m, K, d_idx, distances ? d_dist : nullptr, c->compute));
    CU_TRY(cudaMemcpyAsync(indices, d_idx, cells * sizeof(int), cudaMemcpyDeviceToHost, c->compute));
    if (distances) CU_TRY(cudaMemcpyAsync(distances, d_dist, cells * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
    CU_TRY(stream_drain(c->compute));
    CU_TRY(stream_drain(c->copy));
    return NNS_B200_OK;
}

int nns_b200_sample_centre(int k, int n, const float* r_points, float* centre_out)
{
    if (k <= 0 || k > 512 || n < 0 || !centre_out || (n > 0 && !r_points)) return fail(NNS_B200_ERR_INVALID, "invalid arguments");
    TensorCentre c;
    sample_centre_host(k, n, r_points, &c);
    for (int t = 0; t < k; ++t) centre_out[t] = c.c[t];
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
        const char* file;
        int line, code;
        last_cuda_error(&file, &line, &code);
        if (st == NNS_B200_ERR_CUDA || line)
            printf("Error: %s:%d, code:%d, reason: %s \n", file, line, code, cudaGetErrorString((cudaError_t)code));
        else
            printf("Error: nns_b200: %s \n", last_error_text());
        exit(1);
    }
    *results = out;
}

}  // extern "C"
