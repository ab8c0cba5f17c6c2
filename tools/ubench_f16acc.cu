// tools/ubench_f16acc.cu -- can the screen's epilogue work on 16-bit accumulators?
//  (A) throughput of the packed 16-bit min/max instructions (HMNMX2, VIMNMX3.U16x2) against FMNMX3, alone and mixed
//  (B) tcgen05.mma kind::f16 with an F16 accumulator: where the value sits in the 32-bit TMEM cell, what
//      tcgen05.ld ... pack::16b returns, how the accumulation rounds
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../nns-cuda_b200/csrc -o ubench_f16acc ubench_f16acc.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "nns_common.cuh"
#include "tensor_common.cuh"
using namespace nns;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); exit(1);} }while(0)

#define R8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
#define FMNMX3_(i) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(c[i]) : "f"(b[i]), "f"(x));
#define VMX3_(i) asm volatile("{.reg .b32 t; max.u16x2 t, %0, %1; max.u16x2 %0, t, %2;}" : "+r"(ci[i]) : "r"(cj[i]), "r"(xi));
#define HMX2_(i) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(ch[i]) : "r"(cj[i]));
#define HMXB_(i) asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(ch[i]) : "r"(cj[i]));
#define PACK_(i) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(ch[i]) : "f"(b[i]), "f"(c[i]));
enum { T_F3, T_V3, T_H2, T_HB, T_V3_H2, T_F3_H2, T_F3_V3, T_PACK, T_PACK_F3, T_2V3_H2, T_COUNT };
const char* names[] = {"FMNMX3", "VIMNMX3.U16x2", "HMNMX2", "HMNMX2.BF16", "VIMNMX3.U16x2 + HMNMX2 1:1", "FMNMX3 + HMNMX2 1:1",
                       "FMNMX3 + VIMNMX3.U16x2 1:1", "F2FP.PACK", "F2FP.PACK + FMNMX3 1:1", "2 VIMNMX3.U16x2 + HMNMX2"};
const int per_body[] = {8, 8, 8, 8, 16, 16, 16, 8, 16, 24};

template <int T>
__global__ void __launch_bounds__(1024) bench(int iters, float x, float y, unsigned xi, long long* cycles, float* sink)
{
    float b[8], c[8]; unsigned ci[8], cj[8], ch[8];
    for (int i = 0; i < 8; ++i) { b[i] = (i + 1) * y; c[i] = 1e30f; ci[i] = i; cj[i] = xi * i + 77; ch[i] = xi + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (T == T_F3) { R8(FMNMX3_) }
            if (T == T_V3) { R8(VMX3_) }
            if (T == T_H2) { R8(HMX2_) }
            if (T == T_HB) { R8(HMXB_) }
#define M1(i) VMX3_(i) HMX2_(i)
#define M2(i) FMNMX3_(i) HMX2_(i)
#define M3(i) FMNMX3_(i) VMX3_(i)
#define M4(i) PACK_(i) FMNMX3_(i)
#define M5(i) VMX3_(i) HMX2_(i) VMX3_(i)
            if (T == T_V3_H2) { R8(M1) }
            if (T == T_F3_H2) { R8(M2) }
            if (T == T_F3_V3) { R8(M3) }
            if (T == T_PACK) { R8(PACK_) }
            if (T == T_PACK_F3) { R8(M4) }
            if (T == T_2V3_H2) { R8(M5) }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += c[i] + (float)ci[i] + (float)ch[i];
    if (s == 12345.678f) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int T>
void run(int warps, int sms)
{
    long long* d_cyc; float* d_sink;
    CK(cudaMalloc(&d_cyc, sms * sizeof(long long))); CK(cudaMalloc(&d_sink, 4));
    const int iters = 4096;
    bench<T><<<sms, warps * 32>>>(iters, 1.5f, 0.25f, 3u, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    bench<T><<<sms, warps * 32>>>(iters, 1.5f, 0.25f, 3u, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    std::vector<long long> cyc(sms);
    CK(cudaMemcpy(cyc.data(), d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    std::sort(cyc.begin(), cyc.end());
    const double instr = (double)iters * 4 * per_body[T] * warps;
    printf("%-32s warps/SM %2d: %.2f warp-instr/clk/SM\n", names[T], warps, instr / (double)cyc[sms / 2]);
    cudaFree(d_cyc); cudaFree(d_sink);
}

// ---------------------------------------------------------------------------------------------
// (B) one M = 128, N = 64, K = 16 MMA (+ a second accumulating one) with 16-bit accumulators
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32_pack(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// ab_bf16: operands BF16 (else F16); acc: 0 = F16 accumulator, 1 = F32
__global__ void __launch_bounds__(128) mma_probe(const unsigned short* a_img, const unsigned short* b_img, const int ab_bf16, const int c_f32,
                                                 const int second, uint32_t* raw, uint32_t* packed)
{
    __shared__ __align__(1024) unsigned char sm[128 * 32 + 64 * 32];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    unsigned char* a_s = sm;
    unsigned char* b_s = sm + 128 * 32;
    // interleave layout [chunk][row][16 B]
    for (int i = threadIdx.x; i < 128 * 16; i += 128) {
        const int row = i / 16, col = i % 16;
        reinterpret_cast<unsigned short*>(a_s + (col / 8) * 128 * 16 + row * 16)[col % 8] = a_img[i];
    }
    for (int i = threadIdx.x; i < 64 * 16; i += 128) {
        const int row = i / 16, col = i % 16;
        reinterpret_cast<unsigned short*>(b_s + (col / 8) * 64 * 16 + row * 16)[col % 8] = b_img[i];
    }
    const uint32_t bar_a = smem_u32(&bar);
    if (threadIdx.x == 0) { mbar_init(bar_a, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc512(smem_u32(&slot));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t fmt = ab_bf16 ? 1u : 0u;
        const uint32_t idesc = ((uint32_t)(c_f32 ? 1 : 0) << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const u64 ad = umma_desc_interleave(smem_u32(a_s), 128), bd = umma_desc_interleave(smem_u32(b_s), 64);
        tc_mma_bf16(tm, ad, bd, idesc, 0);
        for (int i = 0; i < second; ++i) tc_mma_bf16(tm, ad, bd, idesc, 1);
        tc_commit(bar_a);
    }
    mbar_wait(bar_a, 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5;
    const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
    uint32_t v[64], p[32];
    tmem_ld64(lane_base, v);
    tmem_ld_wait();
    tmem_ld32_pack(lane_base, p);
    tmem_ld_wait();
    for (int i = 0; i < 64; ++i) raw[threadIdx.x * 64 + i] = v[i];
    for (int i = 0; i < 32; ++i) packed[threadIdx.x * 32 + i] = p[i];
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc512(tm);
}

static unsigned short h_f16(float f) { __half h = __float2half_rn(f); unsigned short u; memcpy(&u, &h, 2); return u; }
static unsigned short h_bf16(float f) { __nv_bfloat16 h = __float2bfloat16_rn(f); unsigned short u; memcpy(&u, &h, 2); return u; }
static float f16_f(unsigned short u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }

static void probe(int ab_bf16, int c_f32, int second, int pattern)
{
    std::vector<float> A(128 * 16), B(64 * 16);
    for (int r = 0; r < 128; ++r)
        for (int j = 0; j < 16; ++j) {
            if (pattern == 0) A[r * 16 + j] = (float)((r + j) % 7) * 0.25f - 0.5f;                // exact small values
            else if (pattern == 1) A[r * 16 + j] = j == 0 ? 1.0f : 0.015625f;                      // 1 + 15 * 2^-12 * scale
            else A[r * 16 + j] = j == 0 ? 1.0f : (j == 1 ? -1.0f : 0.015625f);                     // cancellation then small terms
        }
    for (int c = 0; c < 64; ++c)
        for (int j = 0; j < 16; ++j) {
            if (pattern == 0) B[c * 16 + j] = (float)((c * 3 + j) % 5) * 0.5f - 1.0f;
            else if (pattern == 1) B[c * 16 + j] = j == 0 ? 1.0f : 0.015625f * (float)(c + 1);
            else B[c * 16 + j] = j < 2 ? 1.0f : 0.015625f * (float)(c + 1);
        }
    std::vector<unsigned short> a(128 * 16), b(64 * 16);
    for (size_t i = 0; i < a.size(); ++i) a[i] = ab_bf16 ? h_bf16(A[i]) : h_f16(A[i]);
    for (size_t i = 0; i < b.size(); ++i) b[i] = ab_bf16 ? h_bf16(B[i]) : h_f16(B[i]);
    unsigned short *d_a, *d_b; uint32_t *d_raw, *d_p;
    CK(cudaMalloc(&d_a, a.size() * 2)); CK(cudaMalloc(&d_b, b.size() * 2));
    CK(cudaMalloc(&d_raw, 128 * 64 * 4)); CK(cudaMalloc(&d_p, 128 * 32 * 4));
    CK(cudaMemcpy(d_a, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_b, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_raw, 0xEE, 128 * 64 * 4)); CK(cudaMemset(d_p, 0xEE, 128 * 32 * 4));
    mma_probe<<<1, 128>>>(d_a, d_b, ab_bf16, c_f32, second, d_raw, d_p);
    cudaError_t e = cudaDeviceSynchronize();
    printf("\n[probe] operands %s, accumulator %s, %d accumulating repeats, pattern %d: %s\n", ab_bf16 ? "BF16" : "F16", c_f32 ? "F32" : "F16", second, pattern, cudaGetErrorString(e));
    if (e != cudaSuccess) exit(1);
    std::vector<uint32_t> raw(128 * 64), p(128 * 32);
    CK(cudaMemcpy(raw.data(), d_raw, raw.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(p.data(), d_p, p.size() * 4, cudaMemcpyDeviceToHost));
    for (int r : {0, 1, 37, 127}) {
        printf(" row %3d:", r);
        for (int c : {0, 1, 2, 3, 31, 32, 33, 63}) {
            double want = 0;
            for (int j = 0; j < 16; ++j) want += (double)A[r * 16 + j] * (double)B[c * 16 + j];
            want *= (1 + second);
            const uint32_t cell = raw[r * 64 + c];
            float asf; memcpy(&asf, &cell, 4);
            printf("  c%-2d want %.6f cell %08x lo16 %.6f hi16 %.6f f32 %.6f |", c, want, cell, f16_f(cell & 0xffff), f16_f(cell >> 16), asf);
        }
        printf("\n   packed regs 0,1,15,16,31: ");
        for (int i : {0, 1, 15, 16, 31}) printf(" r%-2d %08x (%.5f, %.5f)", i, p[r * 32 + i], f16_f(p[r * 32 + i] & 0xffff), f16_f(p[r * 32 + i] >> 16));
        printf("\n");
    }
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_raw); cudaFree(d_p);
}

// (C) is an F16-accumulator MMA instruction "exact sum of the 16 products and the accumulator, rounded to nearest"?
// Random F16 operands (mixed magnitudes, both signs), one overwriting + `second` accumulating instructions, every
// accumulator compared BIT FOR BIT with that model evaluated in FP64 on the host.
static double round_to_f16(double x)  // round to nearest even, F16 normal / subnormal grid, no double rounding
{
    if (x == 0.0 || !std::isfinite(x)) return x;
    int e;
    std::frexp(x, &e);                // |x| = f * 2^e, f in [0.5, 1)
    int q = e - 11;                   // ulp exponent for 11 significant bits
    if (q < -24) q = -24;             // subnormal grid 2^-24
    const double r = std::nearbyint(std::ldexp(x, -q));  // default rounding mode: to nearest even
    const double y = std::ldexp(r, q);
    return std::fabs(y) > 65504.0 ? std::copysign(INFINITY, x) : y;
}
static void model_check(int second, unsigned seed)
{
    std::vector<float> A(128 * 16), B(64 * 16);
    srand(seed);
    auto rnd = [&]() {
        const double u = (rand() / (double)RAND_MAX) * 2.0 - 1.0;
        const int sh = rand() % 12;   // magnitudes over 12 binades
        return (float)std::ldexp(u, -sh + 3);
    };
    for (auto& v : A) v = f16_f(h_f16(rnd()));
    for (auto& v : B) v = f16_f(h_f16(rnd()));
    std::vector<unsigned short> a(A.size()), b(B.size());
    for (size_t i = 0; i < a.size(); ++i) a[i] = h_f16(A[i]);
    for (size_t i = 0; i < b.size(); ++i) b[i] = h_f16(B[i]);
    unsigned short *d_a, *d_b; uint32_t *d_raw, *d_p;
    CK(cudaMalloc(&d_a, a.size() * 2)); CK(cudaMalloc(&d_b, b.size() * 2));
    CK(cudaMalloc(&d_raw, 128 * 64 * 4)); CK(cudaMalloc(&d_p, 128 * 32 * 4));
    CK(cudaMemcpy(d_a, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_b, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
    mma_probe<<<1, 128>>>(d_a, d_b, 0, 0, second, d_raw, d_p);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> raw(128 * 64);
    CK(cudaMemcpy(raw.data(), d_raw, raw.size() * 4, cudaMemcpyDeviceToHost));
    long mism = 0, off_by_one_toward_zero = 0;
    double worst_rel = 0;
    for (int r = 0; r < 128; ++r)
        for (int c = 0; c < 64; ++c) {
            double dot = 0;
            for (int j = 0; j < 16; ++j) dot += (double)A[r * 16 + j] * (double)B[c * 16 + j];  // exact: 22-bit products, 16 terms
            double acc = round_to_f16(dot);
            for (int i = 0; i < second; ++i) acc = round_to_f16(acc + dot);
            const float got = f16_f((unsigned short)(raw[r * 64 + c] & 0xffff));
            if ((double)got != acc) {
                ++mism;
                if (std::fabs((double)got) < std::fabs(acc)) ++off_by_one_toward_zero;
                if (acc != 0) worst_rel = std::max(worst_rel, std::fabs(((double)got - acc) / acc));
                if (mism <= 5) printf("   mismatch r%d c%d: model %.9g hardware %.9g (exact %.12g)\n", r, c, acc, (double)got, dot * (1 + second));
            }
        }
    printf("[model] %d accumulating repeats, seed %u: %ld of 8192 accumulators differ from round-to-nearest(exact sum) (%ld smaller in magnitude), worst relative difference %.3g\n",
           second, seed, mism, off_by_one_toward_zero, worst_rel);
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_raw); cudaFree(d_p);
}

int main(int argc, char** argv)
{
    if (argc > 1) {  // model check only
        for (unsigned seed = 1; seed <= 4; ++seed) { model_check(0, seed); model_check(1, seed); model_check(3, seed); }
        return 0;
    }
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    for (int w : {16, 32}) {
        run<T_F3>(w, sms); run<T_V3>(w, sms); run<T_H2>(w, sms); run<T_HB>(w, sms); run<T_V3_H2>(w, sms); run<T_F3_H2>(w, sms);
        run<T_F3_V3>(w, sms); run<T_PACK>(w, sms); run<T_PACK_F3>(w, sms); run<T_2V3_H2>(w, sms);
    }
    probe(0, 1, 0, 0);   // sanity: F16 operands, F32 accumulator
    probe(0, 0, 0, 0);   // F16 accumulator
    probe(0, 0, 0, 1);   // rounding inside one instruction
    probe(0, 0, 3, 1);   // rounding across accumulating instructions
    probe(0, 0, 0, 2);   // cancellation
    probe(1, 0, 0, 0);   // BF16 operands with an F16 accumulator: legal?
    return 0;
}
