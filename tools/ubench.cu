// tools/ubench.cu -- instruction-throughput microbenchmarks for the FP32 distance/argmin inner
// loop on sm_100a.  Measures, per SM and per SM-clock cycle, how many warp-instructions of
// each kind retire when every SM runs WARPS warps; all numbers are from clock64() inside the
// kernel, so they are independent of DVFS.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); exit(1);} }while(0)

#define R8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
// scalar
#define FFMA_(i)  asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(x), "f"(y));
#define FFMA_SQ(i) asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(a[i]) : "f"(b[i]));
#define FADD_(i)  asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(x));
#define FMNMX_(i) asm volatile("min.f32 %0, %0, %1;" : "+f"(c[i]) : "f"(b[i]));
#define FMNMX3_(i) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(c[i]) : "f"(b[i]), "f"(x));
#define FSETP_(i) asm volatile("{.reg .pred p; setp.lt.f32 p, %1, %2; selp.u32 %0, 1, %0, p;}" : "+r"(ci[i]) : "f"(b[i]), "f"(x));
#define IADD_(i)  asm volatile("add.u32 %0, %0, %1;" : "+r"(ci[i]) : "r"(xi));
#define VIMNMX_(i) asm volatile("min.u32 %0, %0, %1;" : "+r"(ci[i]) : "r"(cj[i]));
#define VIMNMX3_(i) asm volatile("{.reg .u32 t; min.u32 t, %1, %2; min.u32 %0, %0, t;}" : "+r"(ci[i]) : "r"(cj[i]), "r"(xi));
#define ISETP_(i) asm volatile("{.reg .pred p; setp.lt.u32 p, %1, %2; @p add.u32 %0, %0, 1;}" : "+r"(ci[i]) : "r"(cj[i]), "r"(xi));
// packed
#define FFMA2_(i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(px), "l"(py));
#define FFMA2_SQ(i) asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(p[i]) : "l"(pb[i]));
#define FADD2_(i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(px));
#define FMUL2_(i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(px));

enum { T_FFMA, T_FFMA_SQ, T_FADD, T_FMNMX, T_FMNMX3, T_FSETP, T_IADD, T_FFMA2, T_FFMA2_SQ, T_FADD2, T_FMUL2,
       T_FFMA2_FMNMX, T_FFMA_FMNMX, T_FFMA2_2FMNMX, T_FFMA2_IADD, T_FFMA2_FSETP, T_FFMA2_LDS, T_LDS128, T_FFMA_IADD, T_VIMNMX, T_VIMNMX3, T_FFMA2_VIMNMX, T_FFMA2_2VIMNMX, T_FFMA2_VIMNMX3, T_FFMA2_ISETP, T_LOOPMIX, T_COUNT };
const char* names[] = {"FFMA(3reg)", "FFMA(d*d+a)", "FADD", "FMNMX", "FMNMX3", "FSETP+SEL", "IADD", "FFMA2(3reg)", "FFMA2(d*d+a)", "FADD2", "FMUL2",
       "FFMA2+FMNMX 1:1", "FFMA+FMNMX 1:1", "FFMA2+2xFMNMX", "FFMA2+IADD 1:1", "FFMA2+(FSETP+SEL)", "FFMA2 x4 + LDS.128", "LDS.128 bcast", "FFMA+IADD 1:1", "VIMNMX", "VIMNMX3(2 min.u32)", "FFMA2+VIMNMX 1:1", "FFMA2+2xVIMNMX", "FFMA2+VIMNMX3", "FFMA2+(ISETP+@p IADD)", "6xF2+VIMNMX+VIMNMX3+ISETP"};
// warp-instructions per R8 body per test (for reporting)
const int per_body[] = {8,8,8,8,8,16,8,8,8,8,8, 16,16,24,16,24, 56, 24, 16, 8, 8, 16, 24, 16, 24, 80};

template <int T>
__global__ void __launch_bounds__(1024) bench(int iters, float x, float y, unsigned xi, long long* cycles, float* sink)
{
    __shared__ __align__(16) float sm[1024];
    sm[threadIdx.x % 1024] = x;
    __syncthreads();
    float a[8], b[8], c[8]; unsigned ci[8], cj[8]; u64 p[8], pb[8]; u64 px, py;
    for (int i = 0; i < 8; ++i) { a[i] = i * x; b[i] = (i + 1) * y; c[i] = 1e30f; ci[i] = i; cj[i] = xi * i + 77; 
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[i]), "f"(b[i]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pb[i]) : "f"(b[i]), "f"(a[i])); }
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(px) : "f"(x), "f"(y));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(py) : "f"(y), "f"(x));
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (T == T_FFMA) { R8(FFMA_) }
            if (T == T_FFMA_SQ) { R8(FFMA_SQ) }
            if (T == T_FADD) { R8(FADD_) }
            if (T == T_FMNMX) { R8(FMNMX_) }
            if (T == T_FMNMX3) { R8(FMNMX3_) }
            if (T == T_FSETP) { R8(FSETP_) }
            if (T == T_IADD) { R8(IADD_) }
            if (T == T_FFMA2) { R8(FFMA2_) }
            if (T == T_FFMA2_SQ) { R8(FFMA2_SQ) }
            if (T == T_FADD2) { R8(FADD2_) }
            if (T == T_FMUL2) { R8(FMUL2_) }
#define MIX1(i) FFMA2_(i) FMNMX_(i)
#define MIX2(i) FFMA_(i) FMNMX_(i)
#define MIX3(i) FFMA2_(i) FMNMX_(i) FMNMX3_(i)
#define MIX4(i) FFMA2_(i) IADD_(i)
#define MIX5(i) FFMA2_(i) FSETP_(i)
#define MIX6(i) FFMA_(i) IADD_(i)
            if (T == T_FFMA2_FMNMX) { R8(MIX1) }
            if (T == T_FFMA_FMNMX) { R8(MIX2) }
            if (T == T_FFMA2_2FMNMX) { R8(MIX3) }
            if (T == T_FFMA2_IADD) { R8(MIX4) }
            if (T == T_FFMA2_FSETP) { R8(MIX5) }
            if (T == T_FFMA_IADD) { R8(MIX6) }
#define MIX8(i) FFMA2_(i) VIMNMX_(i)
#define MIX9(i) FFMA2_(i) VIMNMX_(i) VIMNMX_(i)
#define MIX10(i) FFMA2_(i) VIMNMX3_(i)
#define MIX11(i) FFMA2_(i) ISETP_(i)
#define MIX12(i) FADD2_(i) FFMA2_SQ(i) FADD2_(i) FFMA2_SQ(i) VIMNMX_(i) FADD2_(i) FFMA2_SQ(i) VIMNMX3_(i) ISETP_(i)
            if (T == T_VIMNMX) { R8(VIMNMX_) }
            if (T == T_VIMNMX3) { R8(VIMNMX3_) }
            if (T == T_FFMA2_VIMNMX) { R8(MIX8) }
            if (T == T_FFMA2_2VIMNMX) { R8(MIX9) }
            if (T == T_FFMA2_VIMNMX3) { R8(MIX10) }
            if (T == T_FFMA2_ISETP) { R8(MIX11) }
            if (T == T_LOOPMIX) { R8(MIX12) }
            if (T == T_FFMA2_LDS) {
#define LDSB(i) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w) : "r"(sbase + 16u*((i+u*8+it)&63))); c[i] = v.x + v.z; b[i] = v.y + v.w; }
#define MIX7(i) LDSB(i) FFMA2_(i) FFMA2_SQ(i) FFMA2_(i) FFMA2_SQ(i)
                R8(MIX7)
            }
            if (T == T_LDS128) { R8(LDSB) }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) { float lo, hi; asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += a[i] + b[i] + c[i] + ci[i] + cj[i] + lo + hi; }
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int T> void run(int warps, int sms, long long* d_cyc, float* d_sink)
{
    const int iters = 2000;
    bench<T><<<sms, warps * 32>>>(100, 1.0001f, 0.9999f, 3, d_cyc, d_sink);
    bench<T><<<sms, warps * 32>>>(iters, 1.0001f, 0.9999f, 3, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(sms); CK(cudaMemcpy(h.data(), d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    double cyc = (double)h[sms / 2];
    double winstr = (double)iters * 4 * per_body[T] * warps;   // warp-instructions per SM
    printf("%-22s warps/SM=%2d  cycles=%9.0f  warp-instr/clk/SM=%6.3f  (per SMSP %5.3f)\n", names[T], warps, cyc, winstr / cyc, winstr / cyc / 4);
}

template <int T> void sweep(int sms, long long* d_cyc, float* d_sink) { run<T>(4, sms, d_cyc, d_sink); run<T>(8, sms, d_cyc, d_sink); run<T>(16, sms, d_cyc, d_sink); run<T>(32, sms, d_cyc, d_sink); }

int main(int argc, char**)
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s  SMs=%d  clock=%d kHz  smem/SM=%zu  regs/SM=%d  L2=%d\n", prop.name, prop.multiProcessorCount, prop.clockRate, prop.sharedMemPerMultiprocessor, prop.regsPerMultiprocessor, prop.l2CacheSize);
    int sms = prop.multiProcessorCount;
    long long* d_cyc; float* d_sink; CK(cudaMalloc(&d_cyc, sms * sizeof(long long))); CK(cudaMalloc(&d_sink, 64));
    if (argc > 1) {
    sweep<T_FFMA>(sms, d_cyc, d_sink); sweep<T_FFMA_SQ>(sms, d_cyc, d_sink); sweep<T_FADD>(sms, d_cyc, d_sink);
    sweep<T_FMNMX>(sms, d_cyc, d_sink); sweep<T_FMNMX3>(sms, d_cyc, d_sink); sweep<T_FSETP>(sms, d_cyc, d_sink); sweep<T_IADD>(sms, d_cyc, d_sink);
    sweep<T_FFMA2>(sms, d_cyc, d_sink); sweep<T_FFMA2_SQ>(sms, d_cyc, d_sink); sweep<T_FADD2>(sms, d_cyc, d_sink); sweep<T_FMUL2>(sms, d_cyc, d_sink);
    sweep<T_FFMA2_FMNMX>(sms, d_cyc, d_sink); sweep<T_FFMA_FMNMX>(sms, d_cyc, d_sink); sweep<T_FFMA2_2FMNMX>(sms, d_cyc, d_sink);
    sweep<T_FFMA2_IADD>(sms, d_cyc, d_sink); sweep<T_FFMA_IADD>(sms, d_cyc, d_sink); sweep<T_FFMA2_FSETP>(sms, d_cyc, d_sink); sweep<T_FFMA2_LDS>(sms, d_cyc, d_sink); sweep<T_LDS128>(sms, d_cyc, d_sink);
    }
    sweep<T_VIMNMX>(sms, d_cyc, d_sink); sweep<T_VIMNMX3>(sms, d_cyc, d_sink); sweep<T_FFMA2_VIMNMX>(sms, d_cyc, d_sink); sweep<T_FFMA2_2VIMNMX>(sms, d_cyc, d_sink);
    sweep<T_FFMA2_VIMNMX3>(sms, d_cyc, d_sink); sweep<T_FFMA2_ISETP>(sms, d_cyc, d_sink); sweep<T_LOOPMIX>(sms, d_cyc, d_sink);
    return 0;
}
