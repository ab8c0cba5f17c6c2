// tools/lowk_tune.cu -- measures code-generation variants of the low-k search kernel on a B200:
// queries per thread (Q), register budget (MINB CTAs/SM), quads per loop body (UNROLL), software-
// pipelined argmin (PIPE), consumer warps and ring depth.  Every variant must produce bit-identical
// packed keys.  Build: make -C tools lowk_tune   Run: tools/lowk_tune [k] [m] [n]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../nns-cuda_b200/csrc/nns_internal.h"
using namespace nns;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); exit(1);} }while(0)

struct Run { const float* q; int m; const float* header; const float* blocks; int nblocks; u64* keys; int num_sms; };
static std::vector<u64> g_ref_keys;
static int g_only = -1, g_counter = 0;  // argv[4]: run only the variant with this ordinal

template <int K, int Q, int MINB, int UNROLL, bool FILTER, int G = 1>
void variant(const Run& r, int W, int stages)
{
    const int ordinal = g_counter++;
    if (g_only >= 0 && ordinal != g_only) return;
    const void* kern = FILTER ? (const void*)lowk_filter_kernel<K, Q, MINB, UNROLL, G>
                              : (const void*)lowk_exact_kernel<K, Q, false, MINB, UNROLL>;
    const size_t smem = (size_t)LOWK_BAR_BYTES + (size_t)stages * lowk_tile_bytes(K);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = (W + 1) * 32;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    const int qb = 32 * W * Q;
    const int nqb = (r.m + qb - 1) / qb;
    const int slots = r.num_sms * occ;
    const int tb = lowk_tb(K);
    // splits: whole waves, whole tiles
    int best_s = 1, best_bps = r.nblocks; double best_c = 1e300;
    for (int s = 1; s <= (r.nblocks + tb - 1) / tb && s <= 4096; ++s) {
        int b = (r.nblocks + s - 1) / s; b = (b + tb - 1) / tb * tb;
        int se = (r.nblocks + b - 1) / b;
        double waves = (double)(((long long)nqb * se + slots - 1) / slots);
        double c = waves * (b + 4.0);
        if (c < best_c * 0.999) { best_c = c; best_s = se; best_bps = b; }
    }
    dim3 grid(nqb, best_s);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(launch_keys_init(r.keys, r.m, 0));
        CK(cudaEventRecord(e0));
        int zero = 0;
        void* ah[] = {(void*)&r.q, (void*)&r.m, (void*)&r.header, (void*)&r.blocks, (void*)&r.nblocks, (void*)&best_bps, (void*)&zero, (void*)&stages, (void*)&r.keys};
        void* an[] = {(void*)&r.q, (void*)&r.m, (void*)&r.blocks, (void*)&r.nblocks, (void*)&best_bps, (void*)&zero, (void*)&stages, (void*)&r.keys};
        CK(cudaLaunchKernel(kern, grid, dim3(threads), FILTER ? ah : an, smem, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best_ms = std::min(best_ms, ms);
    }
    std::vector<u64> h(r.m);
    CK(cudaMemcpy(h.data(), r.keys, (size_t)r.m * 8, cudaMemcpyDeviceToHost));
    long long bad = 0;
    if (g_ref_keys.empty()) g_ref_keys = h; else for (int i = 0; i < r.m; ++i) bad += (h[i] != g_ref_keys[i]);
    const double pairs = (double)r.m * (double)r.nblocks * LB;
    const double rate = pairs / (best_ms * 1e-3);
    const double peak = 148.0 * 128 * 1.965e9 / (2.0 * K);
    printf("#%02d K=%2d Q=%d MINB=%d UNR=%d FILTER=%d G=%d W=%d st=%d | regs=%3d occ=%d grid=(%d,%d) | %8.3f ms  %6.3f Tpair/s  %5.1f%% of FP32 peak  mismatches=%lld\n",
           ordinal, K, Q, MINB, UNROLL, (int)FILTER, G, W, stages, fa.numRegs, occ, nqb, best_s, best_ms, rate / 1e12, 100.0 * rate / peak, bad);
    fflush(stdout);
}

template <int K, int Q, int MINB>
void sweep_codegen(const Run& r)
{
    variant<K, Q, MINB, 2, false>(r, 8, 4);
    variant<K, Q, MINB, 1, true, 1>(r, 8, 4);
    variant<K, Q, MINB, 2, true, 1>(r, 8, 4);
    variant<K, Q, MINB, 4, true, 1>(r, 8, 4);
    variant<K, Q, MINB, 1, true, 2>(r, 8, 4);
    variant<K, Q, MINB, 2, true, 2>(r, 8, 4);
    variant<K, Q, MINB, 1, true, 4>(r, 8, 4);
    variant<K, Q, MINB, 2, true, 4>(r, 8, 4);
    variant<K, Q, MINB, 2, true, 2>(r, 4, 4);
}

int main(int argc, char** argv)
{
    const int k = argc > 1 ? atoi(argv[1]) : 3;
    const int m = argc > 2 ? atoi(argv[2]) : 65536;
    const int n = argc > 3 ? atoi(argv[3]) : 1048576;
    g_only = argc > 4 ? atoi(argv[4]) : -1;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs=%d | k=%d m=%d n=%d\n", prop.name, prop.multiProcessorCount, k, m, n);
    std::vector<float> hq((size_t)m * k), hr((size_t)n * k);
    unsigned long long st = 88172645463325252ull;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (float)((st >> 40) * (1.0 / 16777216.0)); };
    for (auto& v : hq) v = rnd();
    for (auto& v : hr) v = rnd();
    float *dq, *dr, *dindex; u64* dkeys;
    const int nblocks = (n + LB - 1) / LB;
    CK(cudaMalloc(&dq, hq.size() * 4)); CK(cudaMalloc(&dr, hr.size() * 4));
    CK(cudaMalloc(&dindex, ((size_t)nblocks * (k + 1) * LB + INDEX_HEADER_FLOATS) * 4)); CK(cudaMalloc(&dkeys, (size_t)m * 8));
    CK(cudaMemcpy(dq, hq.data(), hq.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dr, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice));
    CK(launch_index_build(k, n, dr, dindex, dindex + INDEX_HEADER_FLOATS, true, 0));
    CK(cudaDeviceSynchronize());
    Run r{dq, m, dindex, dindex + INDEX_HEADER_FLOATS, nblocks, dkeys, prop.multiProcessorCount};
    if (k == 3) {
        sweep_codegen<3, 4, 1>(r); sweep_codegen<3, 4, 2>(r); sweep_codegen<3, 4, 3>(r);
        sweep_codegen<3, 8, 1>(r); sweep_codegen<3, 8, 2>(r);
        sweep_codegen<3, 6, 1>(r); sweep_codegen<3, 6, 2>(r);
    } else if (k == 16) {
        sweep_codegen<16, 2, 1>(r); sweep_codegen<16, 2, 2>(r); sweep_codegen<16, 4, 1>(r); sweep_codegen<16, 4, 2>(r);
        sweep_codegen<16, 3, 1>(r);
    } else {
        printf("only k=3 and k=16 are instantiated here\n");
    }
    return 0;
}
