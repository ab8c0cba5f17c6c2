// host_state.h -- host-side state shared by the translation units behind the C ABI (capi.cu: planning,
// per-device context, device-resident entry points; ingest.cu: host-pointer ingest, index handles;
// multi.cu: single-process multi-GPU search).  No compute here: every search runs in the CUDA kernels.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "../../include/nns_b200.h"
#include "nns_internal.h"

namespace nns {

// ---- errors (thread-local text, reference-style file:line for CUDA failures) ----
int fail(int status, const char* fmt, ...);
int fail_cuda(cudaError_t e, const char* file, int line);
const char* last_error_text();
void last_cuda_error(const char** file, int* line, int* code);

#define CU_TRY(call)                                                              \
    do {                                                                          \
        const cudaError_t e__ = (call);                                           \
        if (e__ != cudaSuccess) return ::nns::fail_cuda(e__, __FILE__, __LINE__); \
    } while (0)

#define ST_TRY(call)                          \
    do {                                      \
        const int st__ = (call);              \
        if (st__ != NNS_B200_OK) return st__; \
    } while (0)

void count_launches(unsigned long long n);

// ---- planning ----
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
typedef int (*occ_fn)(void* user, int k, int q, int mode, int warps, int stages);
int make_plan(int k, int m, int n, unsigned flags, int num_sms, occ_fn occ, void* occ_user, Plan* p);
unsigned host_flags();  // NNS_B200_FLAGS (environment): flags word of the host-pointer entry points

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- per-device state ----
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

// Pinned staging ring for pageable host arrays (the reference passes malloc'd memory, main.cu:27-34):
// worker threads memcpy the caller's pages into a pinned slot while the copy engine drains the
// previous slot, so a pageable upload runs at min(host memcpy, PCIe) instead of the driver's
// single-threaded staging.
constexpr int STAGE_SLOTS = 3;
struct Staging {
    unsigned char* slot[STAGE_SLOTS] = {nullptr, nullptr, nullptr};
    cudaEvent_t drained[STAGE_SLOTS] = {nullptr, nullptr, nullptr};  // the H2D copy out of the slot has completed
    bool in_flight[STAGE_SLOTS] = {false, false, false};
    size_t slot_bytes = 0;
    int next = 0;
};

struct DeviceCtx {
    int device = -1;
    int num_sms = 0;
    bool ready = false;
    std::mutex mu;  // serialises users of the cached buffers/streams of this device
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaMemPool_t pool = nullptr;  // private stream-ordered pool for the tensor path's scratch
    std::vector<cudaEvent_t> events;
    DevBuf q, r, index, tsec, keys, idx, stats, peer_keys;
    Staging stage;
    std::map<std::tuple<int, int, int, int, int>, int> occ_cache;
};

int ctx_get(int device, DeviceCtx** out);
int buf_reserve(DevBuf* b, size_t bytes);
int ctx_events(DeviceCtx* c, int count);

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

// Wait until everything enqueued on `st` has completed.  cudaStreamSynchronize on a stream that is ALREADY idle
// was measured to stall for 50-700 ms every few calls on B200 / driver 580 (the stream-ordered pool holds ~2 GB
// of freed scratch at that point; a download into pageable memory has just returned, so the stream is known to
// be idle: tools/e2e_probe.py with NNS_B200_TRACE=1).  Asking first costs a microsecond and avoids the call.
static inline cudaError_t stream_drain(cudaStream_t st)
{
    const cudaError_t q = cudaStreamQuery(st);
    if (q == cudaSuccess) return cudaSuccess;
    if (q != cudaErrorNotReady) return q;
    cudaGetLastError();
    return cudaStreamSynchronize(st);
}

// The hot path on device-resident data: plan, launch.  The caller must have made c->device current.
// d_section = tensor section of these n references (NULL: the tcgen05 path is not available).
int search_keys_on(DeviceCtx* c, int k, int m, int n, const float* d_queries, const float* d_header,
                   const float* d_blocks, const float* d_section, int index_base, u64* d_keys, unsigned flags,
                   cudaStream_t st);
// K nearest neighbours on device-resident data (d_section = NULL: FP32 kernel only)
bool topk_wants_tensor(int k, int m, int n, unsigned flags);
int topk_keys_on(DeviceCtx* c, int k, int m, int n, int K, const float* d_queries, const float* d_blocks, const float* d_section,
                 int index_base, u64* d_keys, unsigned flags, cudaStream_t st);
// does the planner put this search on the tcgen05 path (so that the caller builds a tensor section)?
bool plan_wants_tensor(int k, int m, int n, unsigned flags, int num_sms);

// ---- ingest (ingest.cu) ----
// Host array -> device, asynchronously on `st`; pageable sources go through the pinned staging ring.
// Returns once every piece has been handed to the copy engine (the source may then be reused only if
// it was pageable; pinned sources are read by the DMA until the stream reaches this point).
int h2d_async(DeviceCtx* c, void* d_dst, const void* h_src, size_t bytes, cudaStream_t st);
int search_host_on(DeviceCtx* c, int k, int m, int n, const float* s, const float* r, int index_base, u64* h_keys,
                   int* h_idx, u64* ext_keys, float* h_dist);
// the same for a caller that already holds c->mu (multi.cu holds the locks of every GPU of a call)
int search_host_locked(DeviceCtx* c, int k, int m, int n, const float* s, const float* r, int index_base, u64* h_keys,
                       int* h_idx, u64* ext_keys, float* h_dist);
long long ingest_chunk_points(int k, int n, bool tensor);
int check_host_args(int k, int m, int n, const void* s, const void* r, const void* out);
void sample_centre_host(int k, int n, const float* r_points, TensorCentre* out);
void staging_release(DeviceCtx* c);

}  // namespace nns
