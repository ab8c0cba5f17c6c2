// multi.cu -- single-process multi-GPU search over host arrays: replaces v8/v9::cudaCall's OpenMP
// fan-out + per-GPU D2H + serial host merge (core.cu:761-853, 965-1057) on one NVSwitch box.
//
//  * query-sharded (shard_mode 0): every GPU needs the whole reference index.  Instead of G uploads of
//    the full set, GPU g uploads only slice g of the references over ITS PCIe link and builds that
//    slice of the index -- FP32 blocks and tcgen05 operand images -- with kernels that store each
//    row into the same slice of every peer's index (posted NVLink stores): host->device traffic is
//    n k 4 bytes in total, spread over G links, and the all-gather of the built index is fused into
//    the build kernels.  Then each GPU searches its query slice; no result collective.
//  * reference-sharded (shard_mode 1): GPU g uploads, builds and searches slice g for all queries;
//    the per-GPU packed (dist, idx) keys are folded into ONE key array on GPU 0 with system-scope
//    red.min.u64 over NVLink (keys_merge_kernel) -- no NCCL, no host merge; exact lowest-index ties.
// Without peer access both modes fall back to independent per-GPU searches and a host-side MIN.
#include <condition_variable>
#include <string>
#include <thread>

#include "host_state.h"

namespace nns {

static std::mutex g_multi_mu;  // one multi-GPU call at a time: it owns every GPU it spans

class HostBarrier {
public:
    explicit HostBarrier(int n) : n_(n) {}
    void wait()
    {
        std::unique_lock<std::mutex> lk(mu_);
        const int gen = gen_;
        if (++count_ == n_) {
            count_ = 0;
            ++gen_;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&]() { return gen_ != gen; });
        }
    }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    int n_, count_ = 0, gen_ = 0;
};

// peer access between every pair of the first G devices; false if any pair cannot
static bool enable_all_peers(int G)
{
    int prev = 0;
    if (cudaGetDevice(&prev) != cudaSuccess) return false;
    bool ok = true;
    for (int g = 0; g < G && ok; ++g) {
        if (cudaSetDevice(g) != cudaSuccess) { ok = false; break; }
        for (int h = 0; h < G && ok; ++h) {
            if (h == g) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, g, h) != cudaSuccess || !can) { ok = false; break; }
            const cudaError_t pe = cudaDeviceEnablePeerAccess(h, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) ok = false;
            cudaGetLastError();
        }
    }
    cudaSetDevice(prev);
    cudaGetLastError();
    return ok;
}

struct Slice { long long r0, rn; };
static Slice ref_slice(int n, int G, int g)
{
    const long long blocks = ceil_div(n, LB);
    const long long per = (blocks + G - 1) / G * LB;
    const long long r0 = (long long)g * per;
    const long long rn = r0 >= n ? 0 : ((n - r0) < per ? (n - r0) : per);
    return {r0 < n ? r0 : n, rn};
}

// query-sharded with the fused sharded ingest (see the file header).  Runs in the worker thread of GPU g.
struct GatherShared {
    int k, m, n, G;
    const float* s;
    const float* r;
    int* results;
    bool tensor;
    unsigned flags;
    TensorCentre centre;
    std::vector<DeviceCtx*> ctx;
    std::vector<cudaEvent_t> built;  // GPU g's slice has been stored into every index
    HostBarrier* barrier;
};

static int gather_worker(GatherShared& sh, int g, bool* reached_barriers)
{
    DeviceCtx* c = sh.ctx[g];
    DeviceGuard guard;
    ST_TRY(guard.enter(c->device));
    const int k = sh.k, n = sh.n, G = sh.G;
    const size_t bf = index_block_floats(k);
    const int per_q = ceil_div(sh.m, G);
    const int q0 = g * per_q;
    const int qn = q0 >= sh.m ? 0 : ((sh.m - q0) < per_q ? (sh.m - q0) : per_q);
    float* d_index = (float*)c->index.p;
    float* d_section = sh.tensor ? (float*)c->tsec.p : nullptr;
    const Slice sl = ref_slice(n, G, g);

    // 1. own headers: zero (+ the common centre); nobody may publish into them before that
    int st = NNS_B200_OK;
    cudaError_t e = cudaMemsetAsync(d_index, 0, INDEX_HEADER_FLOATS * sizeof(float), c->compute);
    if (e == cudaSuccess && d_section) e = tensor_section_init(k, n, nullptr, d_section, &sh.centre, c->compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->compute);
    if (e != cudaSuccess) st = fail_cuda(e, __FILE__, __LINE__);
    sh.barrier->wait();
    reached_barriers[0] = true;

    // 2. upload + build the own slice into every GPU's index
    if (st == NNS_B200_OK && qn > 0) st = h2d_async(c, c->q.p, sh.s + (size_t)q0 * k, (size_t)qn * k * sizeof(float), c->copy);
    if (st == NNS_B200_OK && sl.rn > 0) {
        const long long chunk = ingest_chunk_points(k, (int)sl.rn, false);
        const int nchunks = (int)((sl.rn + chunk - 1) / chunk);
        st = ctx_events(c, nchunks + 1);
        float* d_r = (float*)c->r.p;
        for (int ci = 0; ci < nchunks && st == NNS_B200_OK; ++ci) {
            const long long j0 = (long long)ci * chunk;            // within the slice
            const int cn = (int)((sl.rn - j0) < chunk ? (sl.rn - j0) : chunk);
            st = h2d_async(c, d_r + j0 * k, sh.r + (sl.r0 + j0) * k, (size_t)cn * k * sizeof(float), c->copy);
            if (st != NNS_B200_OK) break;
            e = cudaEventRecord(c->events[ci], c->copy);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(c->compute, c->events[ci], 0);
            const size_t b0 = (size_t)((sl.r0 + j0) / LB);
            BlockDsts bd{};
            ImageDsts id{};
            bd.count = id.count = G;
            // own index first
            for (int d = 0; d < G; ++d) {
                const int h = (g + d) % G;
                bd.p[d] = (float*)sh.ctx[h]->index.p + INDEX_HEADER_FLOATS + b0 * bf;
                if (d_section)
                    id.p[d] = reinterpret_cast<unsigned char*>((float*)sh.ctx[h]->tsec.p + TENSOR_HDR_FLOATS) +
                              b0 * tensor_image_bytes_per_block(k);
            }
            if (e == cudaSuccess) e = launch_index_build_to(k, cn, d_r + j0 * k, d_index, HDR_PART_MAX + g, bd, false, c->compute);
            if (e == cudaSuccess && d_section)
                e = tensor_image_build(k, cn, bd.p[0], d_section, THDR_PART_MAX + g, THDR_PART_FLAGS + g, id, c->compute);
            count_launches(d_section ? 2 : 1);
            if (e != cudaSuccess) st = fail_cuda(e, __FILE__, __LINE__);
        }
    }
    if (st == NNS_B200_OK) {
        HeaderPeers hp{};
        hp.count = G;
        hp.self = g;
        for (int h = 0; h < G; ++h) {
            hp.header[h] = (float*)sh.ctx[h]->index.p;
            hp.section[h] = sh.tensor ? (float*)sh.ctx[h]->tsec.p : nullptr;
        }
        e = launch_header_publish(hp, g, c->compute);
        count_launches(1);
        if (e != cudaSuccess) st = fail_cuda(e, __FILE__, __LINE__);
    }
    // the event is recorded even after a failure so that peers waiting on it cannot hang
    e = cudaEventRecord(sh.built[g], c->compute);
    if (e != cudaSuccess && st == NNS_B200_OK) st = fail_cuda(e, __FILE__, __LINE__);
    sh.barrier->wait();
    reached_barriers[1] = true;
    if (st != NNS_B200_OK) return st;

    // 3. every slice is in place once all G build events have fired: fold the headers, search
    for (int h = 0; h < G; ++h) CU_TRY(cudaStreamWaitEvent(c->compute, sh.built[h], 0));
    CU_TRY(launch_header_fold(d_index, d_section, G, c->compute));
    count_launches(1);
    if (qn > 0) {
        u64* d_keys = (u64*)c->keys.p;
        int* d_idx = (int*)c->idx.p;
        ST_TRY(ctx_events(c, 1));
        CU_TRY(cudaEventRecord(c->events[0], c->copy));  // the query upload
        CU_TRY(cudaStreamWaitEvent(c->compute, c->events[0], 0));
        CU_TRY(launch_keys_init(d_keys, qn, c->compute));
        ST_TRY(search_keys_on(c, k, qn, n, (const float*)c->q.p, d_index, d_index + INDEX_HEADER_FLOATS, d_section, 0, d_keys,
                              sh.flags, c->compute));
        CU_TRY(launch_keys_unpack(d_keys, qn, d_idx, nullptr, c->compute));
        count_launches(2);
        CU_TRY(cudaMemcpyAsync(sh.results + q0, d_idx, (size_t)qn * sizeof(int), cudaMemcpyDeviceToHost, c->compute));
    }
    CU_TRY(stream_drain(c->compute));
    CU_TRY(stream_drain(c->copy));
    return NNS_B200_OK;
}

}  // namespace nns

using namespace nns;

extern "C" int nns_b200_search_multi(int k, int m, int n, const float* s_points, const float* r_points, int* results,
                                     int num_gpus, int shard_mode)
{
    ST_TRY(check_host_args(k, m, n, s_points, r_points, results));
    if (shard_mode != 0 && shard_mode != 1) return fail(NNS_B200_ERR_INVALID, "shard_mode must be 0 or 1");
    if (m == 0) return NNS_B200_OK;
    int visible = 0;
    CU_TRY(cudaGetDeviceCount(&visible));
    if (num_gpus <= 0 || num_gpus > visible) num_gpus = visible;
    if (num_gpus > MAX_PEERS) num_gpus = MAX_PEERS;
    if (num_gpus <= 0) return fail(NNS_B200_ERR_CUDA, "no CUDA device");
    const int G = num_gpus;
    std::lock_guard<std::mutex> multi_lock(g_multi_mu);
    std::vector<DeviceCtx*> ctx(G);
    for (int g = 0; g < G; ++g) ST_TRY(ctx_get(g, &ctx[g]));
    // this call owns every GPU it spans: all context locks, in device order
    std::vector<std::unique_lock<std::mutex>> locks;
    for (int g = 0; g < G; ++g) locks.emplace_back(ctx[g]->mu);
    DeviceGuard restore;  // the worker threads set their own device; this thread's is restored on every path
    int cur = 0;
    CU_TRY(cudaGetDevice(&cur));
    ST_TRY(restore.enter(cur));

    const bool p2p = G > 1 && enable_all_peers(G);
    const unsigned flags = host_flags();
    std::vector<int> status(G, NNS_B200_OK);
    std::vector<std::string> msgs(G);
    std::vector<std::thread> th;
    auto first_failure = [&]() {
        for (int g = 0; g < G; ++g)
            if (status[g] != NNS_B200_OK) return fail(status[g], "gpu %d: %s", g, msgs[g].c_str());
        return (int)NNS_B200_OK;
    };

    if (shard_mode == 0 && p2p && n > 0) {
        // ---- query-sharded, sharded ingest + fused all-gather of the built index ----
        GatherShared sh{};
        sh.k = k; sh.m = m; sh.n = n; sh.G = G; sh.s = s_points; sh.r = r_points; sh.results = results;
        sh.flags = flags;
        sh.ctx = ctx;
        const int per_q = ceil_div(m, G);
        sh.tensor = plan_wants_tensor(k, per_q, n, flags, ctx[0]->num_sms);
        if (sh.tensor) sample_centre_host(k, n, r_points, &sh.centre);
        const size_t bf = index_block_floats(k);
        const size_t ibytes = ((size_t)INDEX_HEADER_FLOATS + (size_t)ceil_div(n, LB) * bf) * sizeof(float);
        const Slice s0 = ref_slice(n, G, 0);
        sh.built.resize(G);
        for (int g = 0; g < G; ++g) {
            DeviceGuard dg;
            ST_TRY(dg.enter(g));
            DeviceCtx* c = ctx[g];
            ST_TRY(buf_reserve(&c->index, ibytes));
            if (sh.tensor) ST_TRY(buf_reserve(&c->tsec, tensor_section_floats(k, n) * sizeof(float)));
            ST_TRY(buf_reserve(&c->q, (size_t)per_q * k * sizeof(float)));
            ST_TRY(buf_reserve(&c->r, (size_t)s0.rn * k * sizeof(float)));
            ST_TRY(buf_reserve(&c->keys, (size_t)per_q * sizeof(u64)));
            ST_TRY(buf_reserve(&c->idx, (size_t)per_q * sizeof(int)));
            CU_TRY(cudaEventCreateWithFlags(&sh.built[g], cudaEventDisableTiming));
        }
        HostBarrier barrier(G);
        sh.barrier = &barrier;
        for (int g = 0; g < G; ++g) {
            th.emplace_back([&, g]() {
                bool reached[2] = {false, false};
                status[g] = gather_worker(sh, g, reached);
                if (status[g] != NNS_B200_OK) msgs[g] = last_error_text();
                // a worker that failed before a barrier must still arrive, or its peers would wait forever
                for (int b = 0; b < 2; ++b)
                    if (!reached[b]) barrier.wait();
            });
        }
        for (auto& t : th) t.join();
        for (int g = 0; g < G; ++g) cudaEventDestroy(sh.built[g]);
        return first_failure();
    }

    if (shard_mode == 0) {
        // query-sharded without peer access: every GPU ingests the whole reference set
        const int per = ceil_div(m, G);
        for (int g = 0; g < G; ++g) {
            th.emplace_back([&, g]() {
                const int q0 = g * per;
                const int qn = q0 >= m ? 0 : ((m - q0) < per ? (m - q0) : per);
                if (qn > 0)
                    status[g] = search_host_locked(ctx[g], k, qn, n, s_points + (size_t)q0 * k, r_points, 0, nullptr,
                                                   results + q0, nullptr, nullptr);
                if (status[g] != NNS_B200_OK) msgs[g] = last_error_text();
            });
        }
        for (auto& t : th) t.join();
        return first_failure();
    }

    // ---- reference-sharded: GPU g owns a contiguous slice of whole reference blocks (core.cu:781-791
    // without the <= 0 tail defect D9); packed keys merged by integer MIN ----
    if (p2p) {
        DeviceCtx* c0 = ctx[0];
        u64* shared_keys = nullptr;
        {
            DeviceGuard guard;
            ST_TRY(guard.enter(0));
            ST_TRY(buf_reserve(&c0->peer_keys, (size_t)m * sizeof(u64)));
            shared_keys = (u64*)c0->peer_keys.p;
            CU_TRY(launch_keys_init(shared_keys, m, c0->compute));
            CU_TRY(cudaStreamSynchronize(c0->compute));
            count_launches(1);
        }
        for (int g = 0; g < G; ++g) {
            th.emplace_back([&, g]() {
                const Slice sl = ref_slice(n, G, g);
                if (sl.rn <= 0) return;
                status[g] = search_host_locked(ctx[g], k, m, (int)sl.rn, s_points, r_points + sl.r0 * k, (int)sl.r0, nullptr,
                                               nullptr, shared_keys, nullptr);
                if (status[g] != NNS_B200_OK) msgs[g] = last_error_text();
            });
        }
        for (auto& t : th) t.join();
        ST_TRY(first_failure());
        DeviceGuard guard;
        ST_TRY(guard.enter(0));
        ST_TRY(buf_reserve(&c0->idx, (size_t)m * sizeof(int)));
        CU_TRY(launch_keys_unpack(shared_keys, m, (int*)c0->idx.p, nullptr, c0->compute));
        CU_TRY(cudaMemcpyAsync(results, c0->idx.p, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, c0->compute));
        CU_TRY(stream_drain(c0->compute));
        count_launches(1);
        return NNS_B200_OK;
    }
    // no peer access: per-GPU keys to the host, merged there
    std::vector<std::vector<u64>> keys(G);
    for (int g = 0; g < G; ++g) {
        th.emplace_back([&, g]() {
            const Slice sl = ref_slice(n, G, g);
            if (sl.rn <= 0) return;
            keys[g].resize(m);
            status[g] = search_host_locked(ctx[g], k, m, (int)sl.rn, s_points, r_points + sl.r0 * k, (int)sl.r0, keys[g].data(),
                                           nullptr, nullptr, nullptr);
            if (status[g] != NNS_B200_OK) msgs[g] = last_error_text();
        });
    }
    for (auto& t : th) t.join();
    ST_TRY(first_failure());
    for (int i = 0; i < m; ++i) {
        u64 best = KEY_INIT;
        for (int g = 0; g < G; ++g)
            if (!keys[g].empty() && keys[g][i] < best) best = keys[g][i];
        results[i] = (int)(unsigned)(best & 0xffffffffull);
    }
    return NNS_B200_OK;
}
