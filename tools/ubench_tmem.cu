// tools/ubench_tmem.cu -- TMEM-load (tcgen05.ld / LDTM) and min-reduction throughput on sm_100a:
// what bounds the epilogue of the tcgen05 screen (csrc/tensor_search.cu) when the contraction is
// short (k <= 32).  Every SM runs one CTA that owns all 512 TMEM columns; W warps (warp w reads
// the lane quarter w % 4) loop over the columns.  Cycles from clock64(), per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tmem ubench_tmem.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fmin3(float a, float b, float c)
{
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ int imin3(int a, int b, int c)
{
    int d;
    asm("{.reg .s32 t; min.s32 t, %1, %2; min.s32 %0, t, %3;}" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// MODE 0: loads only (x32, wait after each)          1: loads only, 4 x32 then one wait
// MODE 2: 4 x32 + wait + FMNMX3 tree (the kernel's epilogue)
// MODE 3: 4 x32 + wait + half the values through FMNMX3, half through integer VIMNMX3
// MODE 4: 4 x32 + wait + all integer VIMNMX3         5: x16 loads only, wait after each
// MODE 6: software-pipelined: load chunk c+1 while reducing chunk c (FMNMX3)
// MODE 7: like 6 with the mixed float/int reduction
template <int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) tm_bench(int iters, long long* cycles, float* sink)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    // zero the columns so that the values are ordinary floats
    {
        for (int c = 0; c < 512; c += 1) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(base + c), "r"(__float_as_uint(1.0f + c)) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    float facc = 3.0e38f;
    int iacc = 0x7fffffff;
    uint32_t xacc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        // one "tile": 128 columns; the four tiles of TMEM in turn
        const uint32_t taddr = base + (uint32_t)((it & 3) * 128);
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld32(taddr + c * 32, v);
                tmem_wait();
                xacc ^= v[0] ^ v[31];
            }
        } else if (MODE == 5) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint32_t v[16];
                tmem_ld16(taddr + c * 16, v);
                tmem_wait();
                xacc ^= v[0] ^ v[15];
            }
        } else if (MODE >= 6 && MODE <= 10) {
            uint32_t v[2][32];
            tmem_ld32(taddr, v[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                tmem_wait();
                if (c < 3) tmem_ld32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
                const uint32_t (&cur)[32] = v[c & 1];
                if (MODE == 6) {
                    float c0 = facc, c1 = facc, c2 = facc, c3 = facc;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        c0 = fmin3(c0, __uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1]));
                        c1 = fmin3(c1, __uint_as_float(cur[8 + 2 * j]), __uint_as_float(cur[8 + 2 * j + 1]));
                        c2 = fmin3(c2, __uint_as_float(cur[16 + 2 * j]), __uint_as_float(cur[16 + 2 * j + 1]));
                        c3 = fmin3(c3, __uint_as_float(cur[24 + 2 * j]), __uint_as_float(cur[24 + 2 * j + 1]));
                    }
                    facc = fminf(fmin3(c0, c1, c2), c3);
                } else if (MODE == 8) {
                    uint32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        x0 ^= cur[2 * j] ^ cur[2 * j + 1];
                        x1 ^= cur[8 + 2 * j] ^ cur[8 + 2 * j + 1];
                        x2 ^= cur[16 + 2 * j] ^ cur[16 + 2 * j + 1];
                        x3 ^= cur[24 + 2 * j] ^ cur[24 + 2 * j + 1];
                    }
                    xacc ^= x0 ^ x1 ^ x2 ^ x3;
                } else if (MODE == 9) {
                    // no loop-carried dependency inside the chunk: chains start from the data
                    float c0 = fmin3(__uint_as_float(cur[0]), __uint_as_float(cur[1]), __uint_as_float(cur[2]));
                    float c1 = fmin3(__uint_as_float(cur[8]), __uint_as_float(cur[9]), __uint_as_float(cur[10]));
                    float c2 = fmin3(__uint_as_float(cur[16]), __uint_as_float(cur[17]), __uint_as_float(cur[18]));
                    float c3 = fmin3(__uint_as_float(cur[24]), __uint_as_float(cur[25]), __uint_as_float(cur[26]));
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        c0 = fmin3(c0, __uint_as_float(cur[3 + 2 * j]), __uint_as_float(cur[4 + 2 * j]));
                        c1 = fmin3(c1, __uint_as_float(cur[11 + 2 * j]), __uint_as_float(cur[12 + 2 * j]));
                        c2 = fmin3(c2, __uint_as_float(cur[19 + 2 * j]), __uint_as_float(cur[20 + 2 * j]));
                        c3 = fmin3(c3, __uint_as_float(cur[27 + 2 * j]), __uint_as_float(cur[28 + 2 * j]));
                    }
                    c0 = fminf(c0, __uint_as_float(cur[7]));
                    c1 = fminf(c1, __uint_as_float(cur[15]));
                    c2 = fminf(c2, __uint_as_float(cur[23]));
                    c3 = fminf(c3, __uint_as_float(cur[31]));
                    facc = fminf(facc, fminf(fmin3(c0, c1, c2), c3));
                } else if (MODE == 10) {
                    // plain 2-input FMNMX, 8 independent chains
                    float c[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) c[j] = fminf(__uint_as_float(cur[4 * j]), __uint_as_float(cur[4 * j + 1]));
#pragma unroll
                    for (int j = 0; j < 8; ++j) c[j] = fminf(c[j], __uint_as_float(cur[4 * j + 2]));
#pragma unroll
                    for (int j = 0; j < 8; ++j) c[j] = fminf(c[j], __uint_as_float(cur[4 * j + 3]));
                    facc = fminf(facc, fminf(fminf(fminf(c[0], c[1]), fminf(c[2], c[3])), fminf(fminf(c[4], c[5]), fminf(c[6], c[7]))));
                } else {
                    float c0 = facc, c1 = facc;
                    int i0 = iacc, i1 = iacc;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        c0 = fmin3(c0, __uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1]));
                        i0 = imin3(i0, (int)cur[8 + 2 * j], (int)cur[8 + 2 * j + 1]);
                        c1 = fmin3(c1, __uint_as_float(cur[16 + 2 * j]), __uint_as_float(cur[16 + 2 * j + 1]));
                        i1 = imin3(i1, (int)cur[24 + 2 * j], (int)cur[24 + 2 * j + 1]);
                    }
                    facc = fminf(c0, c1);
                    iacc = min(i0, i1);
                }
            }
        } else {
            uint32_t v[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld32(taddr + c * 32, v[c]);
            tmem_wait();
            if (MODE == 1) {
#pragma unroll
                for (int c = 0; c < 4; ++c) xacc ^= v[c][0] ^ v[c][31];
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t (&cur)[32] = v[c];
                    if (MODE == 2) {
                        float c0 = facc, c1 = facc, c2 = facc, c3 = facc;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            c0 = fmin3(c0, __uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1]));
                            c1 = fmin3(c1, __uint_as_float(cur[8 + 2 * j]), __uint_as_float(cur[8 + 2 * j + 1]));
                            c2 = fmin3(c2, __uint_as_float(cur[16 + 2 * j]), __uint_as_float(cur[16 + 2 * j + 1]));
                            c3 = fmin3(c3, __uint_as_float(cur[24 + 2 * j]), __uint_as_float(cur[24 + 2 * j + 1]));
                        }
                        facc = fminf(fmin3(c0, c1, c2), c3);
                    } else if (MODE == 3) {
                        float c0 = facc, c1 = facc;
                        int i0 = iacc, i1 = iacc;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            c0 = fmin3(c0, __uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1]));
                            i0 = imin3(i0, (int)cur[8 + 2 * j], (int)cur[8 + 2 * j + 1]);
                            c1 = fmin3(c1, __uint_as_float(cur[16 + 2 * j]), __uint_as_float(cur[16 + 2 * j + 1]));
                            i1 = imin3(i1, (int)cur[24 + 2 * j], (int)cur[24 + 2 * j + 1]);
                        }
                        facc = fminf(c0, c1);
                        iacc = min(i0, i1);
                    } else {
                        int i0 = iacc, i1 = iacc, i2 = iacc, i3 = iacc;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            i0 = imin3(i0, (int)cur[2 * j], (int)cur[2 * j + 1]);
                            i1 = imin3(i1, (int)cur[8 + 2 * j], (int)cur[8 + 2 * j + 1]);
                            i2 = imin3(i2, (int)cur[16 + 2 * j], (int)cur[16 + 2 * j + 1]);
                            i3 = imin3(i3, (int)cur[24 + 2 * j], (int)cur[24 + 2 * j + 1]);
                        }
                        iacc = min(min(i0, i1), min(i2, i3));
                    }
                }
            }
        }
    }
    const long long t1 = clock64();
    if (facc == 123.456f || iacc == 12345 || xacc == 0x12345u) sink[0] = facc + iacc + xacc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (threadIdx.x == 0 && blockIdx.x == 0) { sink[1] = facc; sink[2] = (float)iacc; sink[3] = (float)xacc; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

// register-only min throughput: FMNMX3 alone, VIMNMX3 alone, alternating
template <int MODE>
__global__ void __launch_bounds__(1024, 1) min_bench(int iters, long long* cycles, float* sink, float x)
{
    float f[8];
    int g[8];
    float a = x, b = x * 2;
    int ia = __float_as_int(x), ib = ia + 3;
    for (int i = 0; i < 8; ++i) { f[i] = 1e30f + i; g[i] = 0x7f000000 + i; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0 || MODE == 2) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(b));
                if (MODE == 1 || MODE == 2) asm volatile("{.reg .s32 t; min.s32 t, %0, %1; min.s32 %0, t, %2;}" : "+r"(g[i]) : "r"(ia), "r"(ib));
                if (MODE == 3) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(b)); asm volatile("min.s32 %0, %0, %1;" : "+r"(g[i]) : "r"(ia)); }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += f[i] + g[i];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static double median_cycles(long long* d_cyc, int sms)
{
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    return (double)h[sms / 2];
}

template <int MODE> void run_tm(const char* name, int sms, long long* d_cyc, float* d_sink)
{
    const int iters = 4000;
    for (int warps : {4, 8, 16}) {
        if (warps <= 8) {
            tm_bench<MODE, 256><<<sms, warps * 32>>>(50, d_cyc, d_sink);
            tm_bench<MODE, 256><<<sms, warps * 32>>>(iters, d_cyc, d_sink);
        } else {
            tm_bench<MODE, 512><<<sms, warps * 32>>>(50, d_cyc, d_sink);
            tm_bench<MODE, 512><<<sms, warps * 32>>>(iters, d_cyc, d_sink);
        }
        CK(cudaDeviceSynchronize());
        const double cyc = median_cycles(d_cyc, sms);
        const double values = (double)iters * 128 * 32 * warps;  // 32-bit values read per SM
        float hs[4];
        CK(cudaMemcpy(hs, d_sink, sizeof(hs), cudaMemcpyDeviceToHost));
        printf("%-44s warps=%2d cycles=%9.0f  values/clk/SM=%7.2f  B/clk/SM=%7.1f  clk per 256x128 tile=%7.1f  [facc=%g iacc=%g]\n", name, warps, cyc,
               values / cyc, 4 * values / cyc, 32768.0 / (values / cyc), hs[1], hs[2]);
    }
}

template <int MODE> void run_min(const char* name, int vals_per_body, int sms, long long* d_cyc, float* d_sink)
{
    const int iters = 4000;
    for (int warps : {4, 8, 16, 32}) {
        min_bench<MODE><<<sms, warps * 32>>>(50, d_cyc, d_sink, 1.5f);
        min_bench<MODE><<<sms, warps * 32>>>(iters, d_cyc, d_sink, 1.5f);
        CK(cudaDeviceSynchronize());
        const double cyc = median_cycles(d_cyc, sms);
        const double values = (double)iters * 4 * 8 * vals_per_body * 32 * warps;
        printf("%-44s warps=%2d cycles=%9.0f  new values/clk/SM=%7.2f\n", name, warps, cyc, values / cyc);
    }
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s SMs=%d\n", prop.name, sms);
    long long* d_cyc;
    float* d_sink;
    CK(cudaMalloc(&d_cyc, sms * sizeof(long long)));
    CK(cudaMalloc(&d_sink, 64));
    run_min<0>("FMNMX3 (2 new values / instr)", 2, sms, d_cyc, d_sink);
    run_min<1>("VIMNMX3 = 2 x min.s32 (2 new values)", 2, sms, d_cyc, d_sink);
    run_min<2>("FMNMX3 + VIMNMX3 alternating (4 new values)", 4, sms, d_cyc, d_sink);
    run_min<3>("FMNMX3 + VIMNMX alternating (3 new values)", 3, sms, d_cyc, d_sink);
    run_tm<0>("LDTM x32, wait each", sms, d_cyc, d_sink);
    run_tm<1>("LDTM 4 x x32, one wait", sms, d_cyc, d_sink);
    run_tm<5>("LDTM x16, wait each", sms, d_cyc, d_sink);
    run_tm<2>("LDTM 4 x x32 + FMNMX3 tree", sms, d_cyc, d_sink);
    run_tm<3>("LDTM 4 x x32 + FMNMX3/VIMNMX3 halves", sms, d_cyc, d_sink);
    run_tm<4>("LDTM 4 x x32 + VIMNMX3", sms, d_cyc, d_sink);
    run_tm<6>("LDTM pipelined x32 + FMNMX3", sms, d_cyc, d_sink);
    run_tm<7>("LDTM pipelined x32 + FMNMX3/VIMNMX3", sms, d_cyc, d_sink);
    run_tm<8>("LDTM pipelined x32 + LOP3 xor", sms, d_cyc, d_sink);
    run_tm<9>("LDTM pipelined x32 + FMNMX3 (no carried dep)", sms, d_cyc, d_sink);
    run_tm<10>("LDTM pipelined x32 + FMNMX 2-input", sms, d_cyc, d_sink);
    return 0;
}
