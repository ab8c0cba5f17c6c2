// tools/ubench_sync.cu -- cost (clk per instruction, one warp per scheduler and 4 warps per scheduler) of the
// synchronisation instructions on the tcgen05 screen's per-tile path: tcgen05.wait::ld with nothing
// outstanding, tcgen05.fence::before/after_thread_sync, __syncwarp, mbarrier.arrive, mbarrier.try_wait
// on a completed phase (default and with a suspend-time hint), clock64.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); return 1;} }while(0)

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* cycles, unsigned* sink)
{
    __shared__ __align__(8) unsigned long long bar[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[warp]);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    __syncthreads();
    unsigned acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (MODE == 1) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (MODE == 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (MODE == 3) __syncwarp();
            if (MODE == 4) { if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
            if (MODE == 5 || MODE == 6) {
                // phase parity 1 of a fresh barrier counts as complete: returns true at once
                uint32_t done;
                if (MODE == 5)
                    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(b), "r"(1u) : "memory");
                else
                    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(b), "r"(1u), "r"(0x989680u) : "memory");
                acc += done;
            }
            if (MODE == 7) acc += (unsigned)clock64();
            if (MODE == 8) { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncwarp(); }
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> int run(const char* name, int sms, long long* d_cyc, unsigned* d_sink)
{
    const int iters = 2000;
    for (int warps : {4, 16}) {
        k<MODE><<<sms, warps * 32>>>(20, d_cyc, d_sink);
        k<MODE><<<sms, warps * 32>>>(iters, d_cyc, d_sink);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(sms);
        CK(cudaMemcpy(h.data(), d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
        std::sort(h.begin(), h.end());
        printf("%-52s warps/SM=%2d  clk per instruction (per warp) = %7.1f\n", name, warps, (double)h[sms / 2] / (iters * 8.0));
    }
    return 0;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    long long* d_cyc; unsigned* d_sink;
    CK(cudaMalloc(&d_cyc, sms * sizeof(long long))); CK(cudaMalloc(&d_sink, 64));
    run<0>("tcgen05.wait::ld (nothing outstanding)", sms, d_cyc, d_sink);
    run<1>("tcgen05.fence::before_thread_sync", sms, d_cyc, d_sink);
    run<2>("tcgen05.fence::after_thread_sync", sms, d_cyc, d_sink);
    run<3>("__syncwarp", sms, d_cyc, d_sink);
    run<4>("mbarrier.arrive (lane 0)", sms, d_cyc, d_sink);
    run<5>("mbarrier.try_wait, completed phase", sms, d_cyc, d_sink);
    run<6>("mbarrier.try_wait + suspend hint, completed phase", sms, d_cyc, d_sink);
    run<7>("clock64", sms, d_cyc, d_sink);
    run<8>("fence::before_thread_sync + __syncwarp", sms, d_cyc, d_sink);
    return 0;
}
