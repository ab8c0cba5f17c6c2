// tensor_longk.cu -- the tcgen05 screen for long contractions, 128 < k <= TENSOR_MAX_K (509).
//
// Same screen as tensor_search.cu (BF16 operands, FP32 accumulators in TMEM, fused per-query minimum
// epilogue, candidate records, exact FP32 re-score by tensor_rescore_kernel), but the contraction is a
// K-LOOP: the operand images are KB = ceil((k + 3) / 64) blocks of 64 columns (the three |r'|^2 columns
// ride in the padding of the last block), the query strip's A image stays resident in shared memory for
// the whole CTA, and the B ring holds ONE 64-column block of one reference tile per stage (16 KiB), so
// its depth no longer depends on k.  One accumulator unit = one 128-reference tile; the MMA issuer walks
// tile -> block -> 4 K = 16 steps (-> accumulator half), committing the stage after each block and the
// accumulator after the last.  With KB >= 3 a tile is >= 24 MMAs of 64 clk: the tensor pipe is the
// bound and neither the issue rate nor the epilogue matters, so descriptors are built at run time.
//   k <= 317 (KB <= 5): 256 query rows per CTA (two M = 128 accumulator halves, A <= 160 KiB)
//   k <= 509 (KB <= 8): 128 query rows per CTA (one half, A <= 128 KiB); the two epilogue warps that
//                       share a TMEM lane quarter split the tile's 128 columns
// Replaces, for these k, the FP32 `wide` kernel whose 4 queries per CTA re-read the index m/4 times.
#include "tensor_common.cuh"

namespace nns {

constexpr int LK_BLOCK_BYTES = T_BN * 128;  // one 64-column block of one reference tile
constexpr int LK_MAX_STAGES = 8;

template <int HALVES>
__global__ void __maxnreg__(T_MAX_REGS)
tensor_screen_longk_kernel(const unsigned char* __restrict__ qimage, const int m, const unsigned char* __restrict__ rimage,
                           const int ntiles, const int tiles_per_split, const int KB, const int stages,
                           const float* __restrict__ band, unsigned* __restrict__ approx_min, const CandBuf cb)
{
    constexpr int ROWS = 128 * HALVES;
    constexpr int NBUF = 2;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(T_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t a_bytes = (uint32_t)KB * ROWS * 128;
    const uint32_t b_bytes = (uint32_t)KB * LK_BLOCK_BYTES;  // one reference tile, all blocks (global image stride)
    unsigned char* a_smem = smem;
    unsigned char* b_smem = smem + a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + (size_t)stages * LK_BLOCK_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * LK_MAX_STAGES + 8);
    unsigned* s_cand_count = tmem_slot + 1;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t b_full = bar0, b_empty = bar0 + 8 * LK_MAX_STAGES;
    const uint32_t acc_full = bar0 + 8 * 2 * LK_MAX_STAGES, acc_empty = acc_full + 8 * NBUF, a_full = acc_empty + 8 * NBUF;

    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const int t0 = (int)blockIdx.y * tiles_per_split;
    const int nt = min(ntiles, t0 + tiles_per_split) - t0;
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
    __shared__ unsigned s_abort;
    if (threadIdx.x == 0) s_abort = *reinterpret_cast<volatile const unsigned*>(cb.status + 1);
    __syncthreads();
    if (nt <= 0 || s_abort != 0u) {
        if (threadIdx.x == 0) cb.cta_count[cta] = 0;
        return;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, 1); }
        for (int i = 0; i < NBUF; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, T_TEAM_WARPS); }
        mbar_init(a_full, 1);
        mbar_fence_init();
        *s_cand_count = 0;
    }
    if (warp == 1) tmem_alloc512(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- producer ----------------
        if (lane == 0) {
            mbar_arrive_expect_tx(a_full, a_bytes);
            for (int kb = 0; kb < KB; ++kb)  // the strip's A image, block by block (32 / 16 KiB each)
                bulk_g2s(smem_u32(a_smem + (size_t)kb * ROWS * 128), qimage + (size_t)blockIdx.x * a_bytes + (size_t)kb * ROWS * 128,
                         ROWS * 128, a_full);
            int s = 0;
            uint32_t ph = 0;
            for (int t = 0; t < nt; ++t)
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(b_empty + 8 * s, ph ^ 1u);
                    mbar_arrive_expect_tx(b_full + 8 * s, LK_BLOCK_BYTES);
                    bulk_g2s(smem_u32(b_smem + (size_t)s * LK_BLOCK_BYTES),
                             rimage + (size_t)(t0 + t) * b_bytes + (size_t)kb * LK_BLOCK_BYTES, LK_BLOCK_BYTES, b_full + 8 * s);
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (elect_one_sync()) {
            mbar_wait(a_full, 0);
            const u64 adesc0 = umma_desc_sw128(smem_u32(a_smem));
            const u64 bdesc0 = umma_desc_sw128(smem_u32(b_smem));
            int s = 0;
            uint32_t ph = 0;
            for (int u = 0; u < nt; ++u) {
                const int buf = u & 1;
                mbar_wait_mma(acc_empty + 8 * buf, (uint32_t)(((u >> 1) & 1) ^ 1));  // its team drained this buffer
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait_mma(b_full + 8 * s, ph);
                    tc_fence_after();
                    const u64 bd = bdesc0 + (u64)((uint32_t)(s * LK_BLOCK_BYTES) >> 4);
                    const u64 ad = adesc0 + (u64)((uint32_t)(kb * ROWS * 128) >> 4);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                        for (int h = 0; h < HALVES; ++h)
                            tc_mma_bf16(tmem_base + (uint32_t)((buf * HALVES + h) * T_BN), ad + (u64)((h * 128 * 128 + ks * 32) >> 4),
                                        bd + (u64)((ks * 32) >> 4), IDESC, (uint32_t)((kb | ks) != 0));
                    tc_commit(b_empty + 8 * s);  // stage free once its MMAs have read it
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
                tc_commit(acc_full + 8 * buf);  // accumulator complete
            }
        }
    } else {
        // ---------------- epilogue: thread = query row ----------------
        const int e = warp - T_SERVICE_WARPS;
        const int team = e >> 3;        // accumulator buffer / tile parity
        const int lq = warp & 3;        // TMEM lane quarter this warp may access
        const int hsel = (e >> 2) & 1;  // HALVES = 2: accumulator half; HALVES = 1: which half of the tile's columns
        const int half = HALVES == 2 ? hsel : 0;
        const int row = half * 128 + lq * 32 + lane;
        const long long q = (long long)blockIdx.x * ROWS + row;
        const float my_band = (q < m) ? band[q] : -inf_f();
        float run_min = (q < m) ? ord2f(approx_min[q]) : inf_f();
        float thresh = run_min + my_band;
        const uint32_t lane_base = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(half * T_BN);
        const int c_lo = HALVES == 2 ? 0 : 2 * hsel, c_hi = HALVES == 2 ? 4 : 2 * hsel + 2;  // 32-column chunks of this warp
        uint32_t v[32];
        for (int u = team; u < nt; u += 2) {
            const int buf = u & 1;
            const uint32_t taddr = lane_base + (uint32_t)(buf * HALVES * T_BN);
            const int unit0 = (t0 + u) * (T_BN / 32);
            mbar_wait_hot(acc_full + 8 * buf, (uint32_t)((u >> 1) & 1));
            tc_fence_after();
#pragma unroll 1
            for (int c = c_lo; c < c_hi; ++c) {
                tmem_ld32(taddr + c * 32, v);
                tmem_ld_wait();
                if (c + 1 == c_hi) {  // every TMEM read of this warp for this unit has completed
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + 8 * buf);
                }
                float c0 = min3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
                float c1 = min3(__uint_as_float(v[8]), __uint_as_float(v[9]), __uint_as_float(v[10]));
                float c2 = min3(__uint_as_float(v[16]), __uint_as_float(v[17]), __uint_as_float(v[18]));
                float c3 = min3(__uint_as_float(v[24]), __uint_as_float(v[25]), __uint_as_float(v[26]));
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    c0 = min3(c0, __uint_as_float(v[3 + 2 * j]), __uint_as_float(v[4 + 2 * j]));
                    c1 = min3(c1, __uint_as_float(v[11 + 2 * j]), __uint_as_float(v[12 + 2 * j]));
                    c2 = min3(c2, __uint_as_float(v[19 + 2 * j]), __uint_as_float(v[20 + 2 * j]));
                    c3 = min3(c3, __uint_as_float(v[27 + 2 * j]), __uint_as_float(v[28 + 2 * j]));
                }
                c0 = min3(c0, __uint_as_float(v[7]), __uint_as_float(v[15]));
                c2 = min3(c2, __uint_as_float(v[23]), __uint_as_float(v[31]));
                const float cm = fminf(min3(c0, c1, c2), c3);
                if (cm <= thresh) {
                    TensorCand cnd;
                    cnd.q = (int)q; cnd.unit = unit0 + c; cnd.smin = cm;
                    cand_emit(cb, s_cand_count, cta, cnd);
                    if (cm < run_min && !cb.fixed_threshold) {
                        run_min = cm;
                        thresh = run_min + my_band;
                        atomicMin(approx_min + q, f2ord(run_min));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc512(tmem_base);
    if (threadIdx.x == 0) {
        const unsigned used = *s_cand_count;
        cb.cta_count[cta] = min(used, cb.region_cap);
        atomicAdd(cb.status, used);
    }
}

int tensor_longk_rows(int KB) { return KB <= 5 ? 256 : 128; }

// ring depth that fits beside the resident A image
static int longk_stages(int KB, int rows)
{
    const size_t budget = (size_t)227 * 1024 - 1024 - (size_t)KB * rows * 128 - ((2 * LK_MAX_STAGES + 8) * 8 + 16);
    int s = (int)(budget / LK_BLOCK_BYTES);
    return s > LK_MAX_STAGES ? LK_MAX_STAGES : s;
}

cudaError_t tensor_longk_launch(int KB, dim3 grid, cudaStream_t st, const unsigned char* qimage, int m, const unsigned char* rimage,
                                int ntiles, int tps, const float* band, unsigned* amin, const CandBuf& cb)
{
    const int rows = tensor_longk_rows(KB);
    const int stages = longk_stages(KB, rows);
    if (stages < 2) return cudaErrorInvalidValue;
    const size_t smem = (size_t)KB * rows * 128 + (size_t)stages * LK_BLOCK_BYTES + (2 * LK_MAX_STAGES + 8) * 8 + 16;
    cudaError_t e;
    if (rows == 256) {
        e = cudaFuncSetAttribute(tensor_screen_longk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tensor_screen_longk_kernel<2><<<grid, T_THREADS, smem, st>>>(qimage, m, rimage, ntiles, tps, KB, stages, band, amin, cb);
    } else {
        e = cudaFuncSetAttribute(tensor_screen_longk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tensor_screen_longk_kernel<1><<<grid, T_THREADS, smem, st>>>(qimage, m, rimage, ntiles, tps, KB, stages, band, amin, cb);
    }
    return cudaGetLastError();
}

}  // namespace nns
