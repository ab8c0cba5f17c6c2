// tensor_search.cu -- the tcgen05 path.  Built for k > 32 (BASELINE config C4, k = 128): the dense
// contraction -2 Q.R^T runs on the 5th-generation tensor cores (tcgen05.mma, BF16 operands, FP32
// accumulators in TMEM), a fused epilogue reduces every 128-reference tile to per-query minima of the
// approximate score S~ = |r'|^2 - 2 q'.r' and emits (query, 32-reference unit) candidates, and an exact
// FP32 re-score of the candidates in V0's subtract-square-accumulate form (core.cu:38-43) decides the
// answer.  The m x n score matrix never leaves TMEM.  With split-precision BF16 columns the same
// screen also serves k <= 32 on large problems (C2, C3), where it is bound by one FMNMX3 lane-slot
// per pair instead of the k FFMA2 lane-slots of the FP32 screened kernel.
//
// Exactness.  q' = fl(q - c), r' = fl(r - c) are the inputs centred on the reference mean c
// (distances are translation invariant; centring shrinks the operands ~4x for data in [0,1]).
// The tensor cores see bf16(-2q') and bf16(r') (hi + lo parts for k <= 42); E(q) bounds |S~ - S| for
// every reference plus the gap between V0's FP32 distance and the real one
// (tensor_query_image_kernel).  A unit is a candidate when its minimum S~ is within 2E of the running
// minimum, and is re-scored when it is within 2E of the final minimum: the unit holding V0's answer
// always qualifies, as does every unit holding an exactly tied reference, and the re-score keeps the
// lowest index through the packed-key atomicMin.  If the candidate buffer overflows (adversarial data:
// all points identical, clusters far denser than the screen resolves) a device flag makes the FP32
// kernel launched right behind redo the search -- there is no host round trip -- and screen CTAs that
// have not started yet give up at once.
//
// Kernel structure (one CTA per 256-query strip x reference range, 18 warps, one CTA per SM):
//   warp 0   producer: 1-D bulk copies (TMA) of the pre-swizzled BF16 images: the strip's A tile
//            once, then the B tiles through an mbarrier ring (G tiles per stage: 4 KiB tiles are
//            moved four at a time so that the ring costs one barrier round trip per 16 KiB)
//   warp 1   MMA issuer: one elected thread issues tcgen05.mma.kind::f16 M=128 N=128 K=16, two
//            accumulator halves (query rows 0-127 / 128-255) x two buffers = all 512 TMEM columns
//            (k <= 9: N=64 units, two per tile, in four buffers); tcgen05.commit publishes the
//            accumulator and releases the B stage
//   warps 2-17 epilogue, two teams of 8 warps: team i owns the TMEM buffers of parity i and reduces
//            the accumulator units u % 2 == i; thread = query row; tcgen05.ld 32 (64) columns at a
//            time, FMNMX3 tree, candidate test.  |r'|^2 is folded into the contraction (A carries
//            1,1,1; B carries |r'|^2 split into three BF16 terms), so the epilogue touches neither
//            shared memory nor the FP32 pipe.
// Why two teams: for short contractions the tile pipeline is a chain of latencies -- commit ->
// epilogue wake-up -> four TMEM-load round trips -> release -> issuer wake-up -> MMA issue -> MMA --
// of ~1300 clk per TMEM buffer, of which only ~150 clk per warp are FMNMX3 issue slots
// (tools/tensor_trace.py, profiles/r1_tensor_trace_*.txt).  Two teams keep both buffers' chains
// running concurrently with four epilogue warps per scheduler.
// Operand images are K-major with the 128-byte swizzle (Swizzle<3,4,3>), or K-major "interleaved"
// 8 x 16 B core matrices for the 16-column steps, written by the prep kernels exactly as the UMMA
// shared-memory descriptors expect them, so plain bulk copies suffice (no tensor maps).
// SASS: UTCHMMA / LDTM / UBLKCP.
//
// F16 mode (round 2; THDR_MODE = TMODE_F16, plain layout, 10 <= k <= 128): F16 operands scaled by powers of two and F16
// accumulators -- the epilogue reads two accumulators per register (tcgen05.ld ... pack::16b) and reduces them with
// HMNMX2, which retires values at twice the rate of FMNMX3; its 58 registers leave room for three MMA-issuing threads
// and three epilogue teams, with the query image in tensor memory and one accumulator buffer per (issuer, team).  The
// scaling, the range guarantees and the error bound are in tensor_common.cuh ("F16 mode"); DESIGN.md 3.3 has the
// measurements that led there.
#include <algorithm>

#include "tensor_common.cuh"

namespace nns {

// Split precision (k <= TENSOR_SPLIT_MAX_K).  BF16 keeps 8 significant bits; splitting each centred
// coordinate into hi = bf16(x) and lo = bf16(x - hi) and laying the contraction dimension out as
//     queries:    [ qh (k) | qh (k) | ql (k) ]        references: [ rh (k) | rl (k) | rh (k) ]
// makes one BF16 MMA pass accumulate qh.rh + qh.rl + ql.rh, i.e. q'.r' up to terms of relative size
// ~3 * 2^-18: the screen's error bound E shrinks ~250x, so that it is selective even for k = 3 with
// millions of references (nearest-neighbour distances ~1e-5 of the data extent).
//
// Precision mode.  For TENSOR_PLAIN_MIN_K <= k <= TENSOR_SPLIT_MAX_K both layouts exist: split columns
// (contraction 3k + 3) or plain BF16 (k + 3, a third of the tensor work; E ~250x larger).  Which one is
// cheaper depends on the data: the screen stays selective as long as the band 2E is small against the
// nearest-neighbour distances (uniform k = 16, n = 16.7 M: 2E = 0.03 vs d^2 = 0.16 -> 3 references inside
// the band), but on data of low intrinsic dimension the plain band holds thousands.  tensor_index_build
// decides per index with a probe (tensor_mode_kernel) and leaves the choice in the section header
// (THDR_MODE); every kernel of the path is launched in both variants and the one that does not match the
// header word exits at once -- no host round trip.
// source dimension and part (0 = hi, 1 = lo) of data column `col` (< ndata); dimension -1 = none
__device__ __forceinline__ void image_column(int k, int ndata, bool split, int col, bool query, int& dim, int& part)
{
    if (col >= ndata) { dim = -1; part = 0; return; }
    if (!split) { dim = col; part = 0; return; }
    const int seg = col / k;
    dim = col - seg * k;
    part = query ? (seg == 2 ? 1 : 0) : (seg == 1 ? 1 : 0);
}
__device__ __forceinline__ __nv_bfloat16 bf16_part(float x, int part)
{
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    if (part == 0) return hi;
    const float rem = x - __bfloat162float(hi);  // exact in FP32
    return __float2bfloat16_rn((fabsf(x) < inf_f()) ? rem : 0.0f);
}

// 16-bit operand element: BF16 (hi / lo part) or, in F16 mode, F16 of the scaled value
__device__ __forceinline__ unsigned short operand_bits(float x, int part, bool f16, float scale)
{
    if (f16) return __half_as_ushort(__float2half_rn(x * scale));
    return __bfloat16_as_ushort(bf16_part(x, part));
}

// ---------------------------------------------------------------------------------------------
// reference-side preparation (part of index_build for k <= 128)
// ---------------------------------------------------------------------------------------------
// The images are built from the FP32 tiled-SoA blocks (coalesced rows; the AoS upload is not needed
// again), so a tensor section can be added to an index at any time.  Any centre is valid (distances
// are translation invariant; it only scales the error bound E), so it is the mean of a strided sample
// of at most TENSOR_CENTRE_BLOCKS reference blocks -- or a centre fixed by the caller, which lets
// several GPUs / ingest chunks build slices of one section independently.
constexpr int TENSOR_CENTRE_BLOCKS = 1024;

// grid.x sampled blocks (block b * stride); warp w sums rows w, w + 4, ... of its block
__global__ void __launch_bounds__(128)
tensor_colsum_kernel(const float* __restrict__ blocks, const int n, const int k, const int stride,
                     float* __restrict__ sums, unsigned* __restrict__ count)
{
    const long long b = (long long)blockIdx.x * stride;
    const int jn = (int)min((long long)LB, n - b * LB);
    if (jn <= 0) return;
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const float* blk = blocks + (size_t)b * (k + 1) * LB;
    for (int t = warp; t < k; t += 4) {
        float s = 0.0f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int r = lane + 32 * e;
            const float x = (r < jn) ? __ldg(blk + (size_t)t * LB + r) : 0.0f;
            if (fabsf(x) <= 1e15f) s += x;  // NaN / INF / huge coordinates do not steer the centre
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) atomicAdd(sums + t, s);
    }
    if (threadIdx.x == 0) atomicAdd(count, (unsigned)jn);
}

__global__ void tensor_centre_kernel(float* __restrict__ hdr, const int k, const TensorCentre fixed, const int use_fixed)
{
    const unsigned cnt = reinterpret_cast<const unsigned*>(hdr)[THDR_MAX + 2];  // samples summed by tensor_colsum_kernel
    for (int t = threadIdx.x; t < THDR_MAX; t += blockDim.x) {
        float c = use_fixed ? (t < k ? fixed.c[t] : 0.0f) : ((t < k && cnt > 0) ? hdr[t] / (float)cnt : 0.0f);
        if (!(fabsf(c) <= 1e15f)) {  // NaN / INF / huge input: the tensor path is disabled for this index
            c = 0.0f;
            atomicOr(reinterpret_cast<unsigned*>(hdr) + THDR_FLAGS, 1u);
        }
        hdr[t] = c;
    }
}

// one CTA per 128-reference block: BF16 image of r' = fl(r - c) (TensorGeom layout) with |r'|^2
// (FP32, split into three BF16 terms; +INF for padded lanes) in columns norm_col .. norm_col + 2.
// F16 mode (f16 != 0): F16 image of s r' with s^2 |r'|^2, s = hdr[THDR_SCALE]; a reference outside the range the
// scaling was chosen for (s |r'| > F16_R_MAX) flags the section unusable.
// Like index_build_kernel it can store to the same slice of several peer GPUs' sections.
__global__ void __launch_bounds__(128)
tensor_ref_image_kernel(const float* __restrict__ blocks, const int n, const int k, const TensorGeom g,
                        float* __restrict__ hdr, const int max_word, const int flag_word, const ImageDsts dst,
                        const unsigned* __restrict__ mode_word, const unsigned my_mode, const int f16)
{
    if (tensor_mode_mismatch(mode_word, my_mode)) return;
    const long long b = blockIdx.x;
    const int row = threadIdx.x;  // one thread per reference
    const long long j = b * T_BN + row;
    const bool valid = j < n;
    const float* col = blocks + (size_t)b * (k + 1) * LB + row;  // coordinate t of this reference: col[t * LB]
    const size_t img_off = (size_t)b * image_bytes(T_BN, g.KB, g.KS);
    const float s = f16 ? hdr[THDR_SCALE] : 1.0f;
    // |r'|^2 first (ascending dimensions), because its columns may share a chunk with data columns
    float rn = 0.0f;
    bool bad = false;
    for (int t = 0; t < k; ++t) {
        float x = 0.0f;
        if (valid) {
            x = __fsub_rn(__ldg(col + (size_t)t * LB), hdr[t]);
            // a NaN / INF coordinate only poisons its own column (that reference cannot win in
            // V0 either); finite but huge values would overflow the error-bound arithmetic
            if (fabsf(x) > 1e15f && fabsf(x) < inf_f()) bad = true;
        }
        rn = __fmaf_rn(x, x, rn);
    }
    const float rv = valid ? rn : inf_f();  // padded lanes can never be a tile minimum
    unsigned short n_hi, n_mid, n_lo;
    bool wipe = false;  // F16: a finite reference out of range -> +INF score, no INF * 0 in the contraction
    if (f16) {
        const float nv = rv * s * s;  // power-of-two scaling: exact
        if (valid && rn < inf_f() && !(nv <= F16_R_MAX * F16_R_MAX)) { bad = true; wipe = true; }
        const __half h_hi = __float2half_rn(wipe ? inf_f() : nv);
        const float rem1 = (__half2float(h_hi) < inf_f()) ? nv - __half2float(h_hi) : 0.0f;
        const __half h_mid = __float2half_rn(rem1);
        const __half h_lo = __float2half_rn(rem1 - __half2float(h_mid));
        n_hi = __half_as_ushort(h_hi); n_mid = __half_as_ushort(h_mid); n_lo = __half_as_ushort(h_lo);
    } else {
        const __nv_bfloat16 b_hi = __float2bfloat16_rn(rv);
        const float rem1 = (rv < inf_f()) ? rv - __bfloat162float(b_hi) : 0.0f;
        const __nv_bfloat16 b_mid = __float2bfloat16_rn(rem1);
        const __nv_bfloat16 b_lo = __float2bfloat16_rn(rem1 - __bfloat162float(b_mid));
        n_hi = __bfloat16_as_ushort(b_hi); n_mid = __bfloat16_as_ushort(b_mid); n_lo = __bfloat16_as_ushort(b_lo);
    }
    const int chunks = g.KB * 8 + g.KS * 2;
    for (int ch = 0; ch < chunks; ++ch) {
        __align__(16) unsigned short v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = ch * 8 + e;
            int dim, part;
            image_column(k, g.ndata, g.split != 0, c, false, dim, part);
            float x = 0.0f;
            if (valid && !wipe && dim >= 0) x = __fsub_rn(__ldg(col + (size_t)dim * LB), hdr[dim]);
            unsigned short o = operand_bits(x, part, f16 != 0, s);
            if (c == g.norm_col) o = n_hi;
            if (c == g.norm_col + 1) o = n_mid;
            if (c == g.norm_col + 2) o = n_lo;
            v[e] = o;
        }
        const size_t o = img_off + image_chunk_at(T_BN, g.KB, row, ch);
        for (int d = 0; d < dst.count; ++d) *reinterpret_cast<uint4*>(dst.p[d] + o) = *reinterpret_cast<const uint4*>(v);
    }
    if (g.en) {  // the norm as ONE F16 number per reference, in an array behind all tile images (read by the screen's epilogue)
        const size_t noff = (size_t)((n + T_BN - 1) / T_BN) * image_bytes(T_BN, g.KB, g.KS) + (size_t)j * 2;
        for (int d = 0; d < dst.count; ++d) *reinterpret_cast<unsigned short*>(dst.p[d] + noff) = n_hi;
    }
    unsigned bits = (valid && rn < inf_f()) ? __float_as_uint(rn) : 0u;  // NaN / INF norms excluded
    bits = __reduce_max_sync(0xffffffffu, bits);
    if ((threadIdx.x & 31) == 0 && bits) atomicMax(reinterpret_cast<unsigned*>(hdr) + max_word, bits);
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(reinterpret_cast<unsigned*>(hdr) + flag_word, 1u);
}

// ---------------------------------------------------------------------------------------------
// query-side preparation (per search call)
// ---------------------------------------------------------------------------------------------
// one CTA per 256-query strip: BF16 image of -2 q' (TensorGeom layout, 1 in the norm columns),
// band[q] = 2 E(q), approx_min[q] = +INF (ordered encoding).
// F16 mode: F16 image of -2 t s q' with t in the norm columns, qscale[q] = t s^2 (0 = the query cannot be screened)
__global__ void __launch_bounds__(256)
tensor_query_image_kernel(const float* __restrict__ queries, const int m, const int k, const TensorGeom g,
                          const float* __restrict__ hdr, unsigned char* __restrict__ image,
                          float* __restrict__ band, unsigned* __restrict__ approx_min, float* __restrict__ qscale,
                          const unsigned* __restrict__ mode_word, const unsigned my_mode, const int f16)
{
    if (tensor_mode_mismatch(mode_word, my_mode)) return;
    const int row = threadIdx.x, rows = blockDim.x;  // rows per strip: 256, or 128 for the longest contractions
    const long long q = (long long)blockIdx.x * rows + row;
    const bool valid = q < m;
    const int KP = g.KB * 64 + g.KS * 16;  // contraction length seen by the MMA
    unsigned char* img = image + (size_t)blockIdx.x * image_bytes(rows, g.KB, g.KS);
    float qn = 0.0f;
    for (int t = 0; t < k; ++t) {
        const float x = valid ? __fsub_rn(__ldg(queries + q * k + t), hdr[t]) : 0.0f;
        qn = __fmaf_rn(x, x, qn);
    }
    const float r2 = __uint_as_float(reinterpret_cast<const unsigned*>(hdr)[THDR_MAX]);
    const float a = sqrtf(qn), rmax = sqrtf(r2);
    const float s = f16 ? hdr[THDR_SCALE] : 1.0f;
    float tq = f16 ? tensor_f16_query_scale(s * a, s * rmax) : 1.0f;        // 0: not representable
    if (f16 && !(tq * s * s < inf_f())) tq = 0.0f;                            // u_q = t s^2 must be an FP32 number (extents below ~1e-17)
    const float xs = f16 ? -2.0f * tq * s : -2.0f;                            // power of two: exact
    const int chunks = g.KB * 8 + g.KS * 2;
    for (int ch = 0; ch < chunks; ++ch) {
        __align__(16) unsigned short v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int col = ch * 8 + e;
            int dim, part;
            image_column(k, g.ndata, g.split != 0, col, true, dim, part);
            float x = 0.0f;
            if (valid && dim >= 0) x = __fsub_rn(__ldg(queries + q * k + dim), hdr[dim]);
            unsigned short o = f16 ? __half_as_ushort(__float2half_rn(xs * x)) : __bfloat16_as_ushort(bf16_part(-2.0f * x, part));
            if (col >= g.norm_col && col < g.norm_col + 3) o = f16 ? __half_as_ushort(__float2half_rn(tq)) : __bfloat16_as_ushort(__float2bfloat16_rn(1.0f));
            v[e] = o;
        }
        *reinterpret_cast<uint4*>(img + image_chunk_at(rows, g.KB, row, ch)) = *reinterpret_cast<const uint4*>(v);
    }
    if (valid) {
        // E(q) >= |S~ - S| + |d_V0 - D'| for every reference: tensor_error_bound / tensor_error_bound_f16 (tensor_common.cuh)
        const float E = f16 ? tensor_error_bound_f16(KP, k, a, rmax, s, tq > 0.0f ? tq : 1.0f, g.en != 0) : tensor_error_bound(g.split != 0, KP, a, rmax);
        const bool flagged = (reinterpret_cast<const unsigned*>(hdr)[THDR_FLAGS] & 1u) != 0;  // NaN / INF / huge / out-of-range references
        // |r'|^2 below 1e-30 is computed from (nearly) denormal FP32 products: the 2^-24 relative terms of E do not hold
        const bool usable = !flagged && (qn <= 1e30f) && (r2 <= 1e30f) && !(r2 > 0.0f && r2 < 1e-30f) && (tq > 0.0f);  // false for NaN too
        band[q] = usable ? 2.0f * E : inf_f();
        approx_min[q] = f2ord(inf_f());
        if (qscale) qscale[q] = f16 ? (tq > 0.0f ? tq * s * s : 1.0f) : 1.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
// G = reference tiles per TMA stage: short contractions (KB = 0) move 4 KiB / 8 KiB tiles, and one
// mbarrier round trip per tile on the MMA issuer's critical path costs more than the MMAs themselves
// SUB = accumulator units per reference tile: the 128-reference tile is issued as SUB MMAs of
// N = 128 / SUB columns into 2 * SUB TMEM buffers.  An epilogue team owns the buffers of its parity,
// so with SUB = 2 it reduces one of its two buffers while the other is being refilled (with one
// buffer per team the team idles for the refill: 23 % of its time in the ncu source view of the
// SUB = 1 build at k = 3).  Only for KB = 0: an N = 64 MMA re-reads the A operand twice as often, and
// at M = 128 the shared-memory operand fetch already runs at its 128 B/clk limit when the tensor
// pipe is the bound (k >= 10).
// TS = the A operand (the strip's query image) lives in TENSOR MEMORY instead of shared memory: the
// epilogue warps of team 0 copy it there once (tcgen05.st), and every MMA then fetches only B from shared
// memory.  At M = 128 the shared-memory operand fetch of the SS form runs at its 128 B/clk limit, which is
// what ruled out N = 64 units (A re-read twice as often) for the longer contractions; with A in TMEM the
// units can be 64 references wide and NBUF = 3 accumulator buffers fit beside A (2 * 3 * 64 + A columns
// <= 512), so an epilogue team has two unit periods to drain a buffer instead of one.
// ISS = MMA-issuing threads (warps 1 .. ISS): a thread's mbarrier wait returns ~250 clk after its last
// tcgen05.commit, so ONE issuer paces 64-reference units at ~265 clk each while their epilogue needs ~150
// (tools/tensor_trace2.py: the epilogue warps start waiting before the unit is even issued).  With ISS = 2
// (only for SUB = 2) issuer i issues sub-unit i of every tile into the buffers of team i; both wait for a B
// stage and both commit its release.
// F16 = F16 operands and F16 accumulators (THDR_MODE = TMODE_F16): the epilogue reads the accumulators two per
// register (tcgen05.ld ... pack::16b) and reduces them with HMNMX2, which retires twice the values per clock of
// FMNMX3; scores are in units of qscale[q] (a power of two) and are compared / recorded unscaled.
// TEAMS = epilogue teams of 8 warps (F16 only; the FP32-accumulator epilogues are written for T_TEAMS): a team's units are
// a serial chain per warp (wait, TMEM load ~140 clk, release, reduce, test: ~300 clk), so units complete at
// ~300 / TEAMS clk; the 16-bit epilogue needs 60 registers, which leaves room for a third team (28 warps, 72 registers).
constexpr int screen_max_regs(int iss, int teams) { return (16384 / (32 * ((1 + iss + teams * T_TEAM_WARPS + 3) / 4))) & ~7; }  // registers are per scheduler
__host__ __device__ constexpr int screen_lcm(int a, int b) { int x = a; while (x % b) x += a; return x; }
// EN (F16 only) = the images carry no norm columns: the epilogue adds t * s^2 |r'|^2 from rnorm[reference] (one F16 per
// reference, read through L1 -- every thread of a warp reads the same 16 bytes) with one HFMA2 per two columns.
template <int KB, int KS, int T_STAGES, int G, int SUB, int NBUF, bool TS, int ISS, bool F16, int TEAMS, bool EN>
__global__ void __maxnreg__(screen_max_regs(ISS, TEAMS))
tensor_screen_kernel(const unsigned char* __restrict__ qimage, const int m, const unsigned char* __restrict__ rimage,
                     const int ntiles, const int tiles_per_split,
                     const float* __restrict__ band, unsigned* __restrict__ approx_min, const float* __restrict__ qscale,
                     const unsigned short* __restrict__ rnorm,
                     const CandBuf cb, const unsigned* __restrict__ mode_word, const unsigned my_mode)
{
    static_assert(!EN || F16, "the epilogue adds the norm only in the F16 mode");
    if (tensor_mode_mismatch(mode_word, my_mode)) return;  // the other precision variant of this launch pair runs
    // KB 64-column swizzled blocks (one 128-byte swizzle row each), then KS interleaved 16-column steps
    constexpr uint32_t A_MAIN = KB * T_BM * 128, B_MAIN = KB * T_BN * 128;
    constexpr uint32_t A_BYTES = (uint32_t)image_bytes(T_BM, KB, KS);  // 72 KiB at KB = 2, KS = 1
    constexpr uint32_t B_BYTES = (uint32_t)image_bytes(T_BN, KB, KS);  // 36 KiB at KB = 2, KS = 1
    constexpr uint32_t STAGE_BYTES = G * B_BYTES;
    constexpr int SN = T_BN / SUB;             // references (TMEM columns) per accumulator unit
    constexpr int STEPS = KB * 4 + KS;         // K = 16 MMA steps per unit and accumulator half
    constexpr uint32_t A_COL0 = 2 * NBUF * SN; // TS: first TMEM column of the A operand, [half][step][8 columns]
    static_assert(!TS || A_COL0 + 2 * STEPS * 8 <= 512, "accumulators + A operand exceed the 512 TMEM columns");
    static_assert(TS || 2 * NBUF * SN <= 512, "accumulators exceed the 512 TMEM columns");
    constexpr uint32_t A_SMEM = TS ? 0u : A_BYTES;
    constexpr int SERVICE = 1 + ISS;           // warp 0 = TMA producer, warps 1 .. ISS = MMA issuers
    static_assert(ISS == 1 || (ISS <= G * SUB && ((NBUF == 4 && (ISS == 2 || ISS == 4)) || (NBUF == 3 && ISS == 3 && F16))),
                  "several issuers: issuer i takes the units u % ISS == i and must own whole buffers");
    // Accumulator barriers.  A barrier may only ever be waited on by ONE team, in phase order (a waiter one phase ahead
    // misreads its parity).  With 2 or 4 buffers a buffer belongs to one team; with 3 buffers the two teams alternate on
    // a buffer, which is only safe while one thread issues all units in order -- several issuers complete their units
    // in any order, so then there is one barrier pair per (buffer, team): unit u uses barrier u % NB, NB = lcm(NBUF, 2).
    constexpr int SPIN = F16 ? NNS_T_SPIN_F16 : NNS_T_SPIN;
    constexpr int NB = ISS > 1 ? screen_lcm(NBUF, TEAMS) : NBUF;
    static_assert(F16 || TEAMS == T_TEAMS, "only the 16-bit epilogue takes a team count");
    constexpr int CPU = SN / 32;               // 32-column chunks per unit
    // instruction descriptor: D = F32, A = B = BF16 (F16 mode: D = A = B = F16, format code 0), both K-major, N = SN, M = 128
    constexpr uint32_t IDESC = (F16 ? 0u : ((1u << 4) | (1u << 7) | (1u << 10))) | ((uint32_t)(SN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_smem = smem;
    unsigned char* b_smem = smem + A_SMEM;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_SMEM + T_STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * T_MAX_STAGES + 8);
    unsigned* s_cand_count = tmem_slot + 1;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t b_full = bar0, b_empty = bar0 + 8 * T_STAGES;
    const uint32_t acc_full = bar0 + 8 * 2 * T_STAGES, acc_empty = acc_full + 8 * NB, a_full = acc_empty + 8 * NB;
    static_assert(2 * T_STAGES + 2 * NB + 1 <= 2 * T_MAX_STAGES + 8, "mbarrier area");

    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const int t0 = (int)blockIdx.y * tiles_per_split;
    const int nt = min(ntiles, t0 + tiles_per_split) - t0;
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
    // nothing to do, or an earlier CTA ran out of candidate space (dense near-ties: the FP32 kernel
    // launched after the re-score redoes the whole search, so the rest of this pass would be wasted)
    // (one thread reads the flag for the whole CTA: it can flip between two threads' reads, and a CTA
    // of which only some threads leave would hang at the first barrier)
    __shared__ unsigned s_abort;
    if (threadIdx.x == 0) s_abort = *reinterpret_cast<volatile const unsigned*>(cb.status + 1);
    __syncthreads();
    if (nt <= 0 || s_abort != 0u) {
        if (threadIdx.x == 0) cb.cta_count[cta] = 0;
        return;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < T_STAGES; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, ISS); }
        for (int i = 0; i < NB; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, T_TEAM_WARPS); }
        mbar_init(a_full, TS ? T_TEAM_WARPS : 1);
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
            if (!TS) {
                mbar_arrive_expect_tx(a_full, A_BYTES);
                bulk_g2s(smem_u32(a_smem), qimage + (size_t)blockIdx.x * A_BYTES, A_BYTES, a_full);
            }
            int s = 0;
            uint32_t ph = 0;
            for (int g0 = 0; g0 < nt; g0 += G) {
                mbar_wait(b_empty + 8 * s, ph ^ 1u);
                T_TRACE(0, g0);
                const uint32_t bytes = (uint32_t)min(G, nt - g0) * B_BYTES;
                mbar_arrive_expect_tx(b_full + 8 * s, bytes);
                bulk_g2s(smem_u32(b_smem + (size_t)s * STAGE_BYTES), rimage + (size_t)(t0 + g0) * B_BYTES, bytes, b_full + 8 * s);
                if (++s == T_STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp < SERVICE) {
        // ---------------- MMA issuer(s) ----------------
        // ONE thread issues every MMA of the CTA, and for short contractions its own instruction
        // stream paces the tile pipeline (ncu source view of the first two-team build: ~550 clk per
        // tile of dependent uniform-datapath arithmetic, ~10 clk per instruction, and an
        // ELECT / branch wrapper of ~45 clk around every tcgen05 instruction issued under
        // `lane == 0`).  Hence: the thread is chosen with elect.sync (ptxas then knows that exactly one
        // thread is active), every descriptor is built once before the loop (per tile only the 14-bit
        // start-address field of the B descriptors moves), and the loop is unrolled over the two
        // TMEM buffers so that buffer addresses and barrier addresses are immediates.
        if (elect_one_sync()) {
            mbar_wait(a_full, 0);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(a_smem), b_addr0 = smem_u32(b_smem);
            u64 adesc_sw[KB > 0 ? KB : 1][4][2], adesc_il[KS > 0 ? KS : 1][2];
            uint32_t bdesc_sw_lo[KB > 0 ? KB : 1][4], bdesc_il_lo[KS > 0 ? KS : 1];
            uint32_t bdesc_sw_hi = 0, bdesc_il_hi = 0;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const u64 bd = umma_desc_sw128(b_addr0 + kb * (T_BN * 128) + ks * 32);
                    bdesc_sw_lo[kb][ks] = (uint32_t)bd;
                    bdesc_sw_hi = (uint32_t)(bd >> 32);
#pragma unroll
                    for (int h = 0; h < 2; ++h) adesc_sw[kb][ks][h] = umma_desc_sw128(a_addr + kb * (T_BM * 128) + h * (128 * 128) + ks * 32);
                }
#pragma unroll
            for (int x = 0; x < KS; ++x) {
                const u64 bd = umma_desc_interleave(b_addr0 + B_MAIN + x * (2 * T_BN * 16), T_BN);
                bdesc_il_lo[x] = (uint32_t)bd;
                bdesc_il_hi = (uint32_t)(bd >> 32);
#pragma unroll
                for (int h = 0; h < 2; ++h) adesc_il[x][h] = umma_desc_interleave(a_addr + A_MAIN + x * (2 * T_BM * 16) + h * (128 * 16), T_BM);
            }
            int s = 0, j = 0;
            uint32_t ph = 0;
            uint32_t boff16 = 0;  // (byte offset of the current B tile inside the ring) >> 4: added to the descriptors' start-address field
            // one accumulator unit: references [sub * SN, sub * SN + SN) of the current tile into buffer `buf`
            // tile_first / tile_last: this is the issuing thread's first / last unit of the tile
            auto issue_unit = [&](const int u, const int buf, const int sub, const bool tile_first, const bool tile_last) {
                T_TRACE(1, u);
                mbar_wait_mma<SPIN>(acc_empty + 8 * buf, (uint32_t)(((u / NBUF) & 1) ^ 1));  // its team drained this buffer
                if (j == 0 && tile_first) mbar_wait_mma<SPIN>(b_full + 8 * s, ph);             // TMA landed this stage
                T_TRACE(2, u);
                tc_fence_after();
                // rows sub * SN .. of the B tile: SN rows of 128 B in a swizzled block, of 16 B in an interleaved chunk
                const uint32_t sw16 = boff16 + (uint32_t)(sub * SN * 128 >> 4), il16 = boff16 + (uint32_t)(sub * SN * 16 >> 4);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const u64 bdesc = ((u64)bdesc_sw_hi << 32) | (u64)(bdesc_sw_lo[kb][ks] + sw16);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (NNS_T_EXPERIMENT >= 3) continue;
                            if (TS) tc_mma_bf16_ts(tmem_base + (uint32_t)((buf * 2 + h) * SN), tmem_base + A_COL0 + (uint32_t)((h * STEPS + kb * 4 + ks) * 8), bdesc, IDESC, (uint32_t)((kb | ks) != 0));
                            else tc_mma_bf16(tmem_base + (uint32_t)((buf * 2 + h) * SN), adesc_sw[kb][ks][h], bdesc, IDESC, (uint32_t)((kb | ks) != 0));
                        }
                    }
#pragma unroll
                for (int x = 0; x < KS; ++x) {  // interleaved steps (the last columns carry |r'|^2)
                    const u64 bdesc = ((u64)bdesc_il_hi << 32) | (u64)(bdesc_il_lo[x] + il16);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (NNS_T_EXPERIMENT >= 3) continue;
                        if (TS) tc_mma_bf16_ts(tmem_base + (uint32_t)((buf * 2 + h) * SN), tmem_base + A_COL0 + (uint32_t)((h * STEPS + KB * 4 + x) * 8), bdesc, IDESC, (uint32_t)((KB | x) != 0));
                        else tc_mma_bf16(tmem_base + (uint32_t)((buf * 2 + h) * SN), adesc_il[x][h], bdesc, IDESC, (uint32_t)((KB | x) != 0));
                    }
                }
                tc_commit(acc_full + 8 * buf);   // accumulator complete
                T_TRACE(3, u);
                if (tile_last) {                 // this thread is done with the tile
                    if (j == G - 1 || u / SUB == nt - 1) {
                        tc_commit(b_empty + 8 * s);  // stage free once the MMAs of its tiles have read it
                        j = 0;
                        if (++s == T_STAGES) { s = 0; ph ^= 1u; }
                        boff16 = (uint32_t)s * (STAGE_BYTES >> 4);
                    } else {
                        ++j;
                        boff16 += B_BYTES >> 4;
                    }
                }
            };
            if constexpr (ISS == 1) {
                // unrolled over the TMEM buffers: buffer and barrier addresses are immediates
                for (int u = 0; u < nt * SUB; u += NBUF) {
#pragma unroll
                    for (int i = 0; i < NBUF; ++i)
                        if (u + i < nt * SUB) {
                            const int sub = (NBUF % SUB == 0) ? i % SUB : (u + i) % SUB;
                            issue_unit(u + i, i, sub, sub == 0, sub == SUB - 1);
                        }
                }
            } else {
                // issuer `id` takes the units u % ISS == id (sub-unit id % 2 of every (ISS / 2)-th tile; with four issuers
                // each owns ONE accumulator buffer).  The issuing thread's own instruction stream is slow (one active
                // thread, dependent scalar code: tools/tensor_trace2.py shows ~370 clk between its commit and its next
                // wait), so what keeps the tensor pipe fed is the number of threads that issue, not their speed.
                const int id = warp - 1;
                const int nunits = nt * SUB;
                constexpr int UPGRP = G * SUB;         // units per B stage (G tiles): every issuer has at least one of them
                int sg = 0, held = -1;                 // stage / tile group this thread currently reads
                uint32_t phg = 0;
                int nb = id % NB;                      // u % NB
                uint32_t par = 0;                      // (u / NB) & 1
                for (int u = id; u < nunits; u += ISS) {
                    const int grp = u / UPGRP;
                    if (grp != held) {
                        if (held >= 0) {               // done with the previous stage once this thread's MMAs have read it
                            tc_commit(b_empty + 8 * sg);
                            if (++sg == T_STAGES) { sg = 0; phg ^= 1u; }
                        }
                        mbar_wait_mma<SPIN>(b_full + 8 * sg, phg);   // TMA landed this stage
                        held = grp;
                    }
                    static_assert(NB == NBUF || NB == 2 * NBUF, "barrier <-> buffer mapping");
                    const int buf = (NB == NBUF) ? nb : (nb >= NBUF ? nb - NBUF : nb);
                    // the previous user of this buffer, unit u - NBUF, has been drained (fresh barrier: parity 1 passes)
                    const int eb = (NB == NBUF) ? nb : (nb >= NBUF ? nb - NBUF : nb + NBUF);
                    const uint32_t epar = (NB == NBUF) ? (par ^ 1u) : (nb >= NBUF ? par : (par ^ 1u));
                    T_TRACE(1, u);
                    mbar_wait_mma<SPIN>(acc_empty + 8 * eb, epar);
                    T_TRACE(2, u);
                    tc_fence_after();
                    const int uin = u - grp * UPGRP;   // unit within the group
                    const uint32_t b16 = (uint32_t)sg * (STAGE_BYTES >> 4) + (uint32_t)(uin / SUB) * (B_BYTES >> 4);
                    const uint32_t sw16 = b16 + (uint32_t)((uin % SUB) * SN * 128 >> 4), il16 = b16 + (uint32_t)((uin % SUB) * SN * 16 >> 4);
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const u64 bdesc = ((u64)bdesc_sw_hi << 32) | (u64)(bdesc_sw_lo[kb][ks] + sw16);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                if (TS) tc_mma_bf16_ts(tmem_base + (uint32_t)((buf * 2 + h) * SN), tmem_base + A_COL0 + (uint32_t)((h * STEPS + kb * 4 + ks) * 8), bdesc, IDESC, (uint32_t)((kb | ks) != 0));
                                else tc_mma_bf16(tmem_base + (uint32_t)((buf * 2 + h) * SN), adesc_sw[kb][ks][h], bdesc, IDESC, (uint32_t)((kb | ks) != 0));
                            }
                        }
#pragma unroll
                    for (int xs = 0; xs < KS; ++xs) {
                        const u64 bdesc = ((u64)bdesc_il_hi << 32) | (u64)(bdesc_il_lo[xs] + il16);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (TS) tc_mma_bf16_ts(tmem_base + (uint32_t)((buf * 2 + h) * SN), tmem_base + A_COL0 + (uint32_t)((h * STEPS + KB * 4 + xs) * 8), bdesc, IDESC, (uint32_t)((KB | xs) != 0));
                            else tc_mma_bf16(tmem_base + (uint32_t)((buf * 2 + h) * SN), adesc_il[xs][h], bdesc, IDESC, (uint32_t)((KB | xs) != 0));
                        }
                    }
                    tc_commit(acc_full + 8 * nb);   // accumulator complete
                    T_TRACE(3, u);
                    nb += ISS;
                    if (nb >= NB) { nb -= NB; par ^= 1u; }
                }
                if (held >= 0) tc_commit(b_empty + 8 * sg);
            }
        }
    } else {
        // ---------------- epilogue: thread = query row ----------------
        // Two teams of 8 warps.  Team i owns TMEM buffer i and reduces the tiles t % 2 == i, so a
        // team has TWO tile periods for the serial path of one tile (wait, four TMEM-load round
        // trips, 76 FMNMX3, release -- ~700 clk measured with tools/tensor_trace.py, of which only
        // 150 clk are issue slots); with 4 epilogue warps per scheduler the ALU pipe stays busy.
        const int e = warp - SERVICE;
        const int team = e >> 3;                // TMEM buffer / tile parity (0 when T_TEAMS == 1)
        const int lq = warp & 3;                // TMEM lane quarter this warp may access
        const int half = (e >> 2) & 1;          // accumulator half (rows 0-127 / 128-255)
        const int row = half * 128 + lq * 32 + lane;
        const long long q = (long long)blockIdx.x * T_BM + row;
        const float my_band = (q < m) ? band[q] : -inf_f();  // rows past m never qualify
        // other CTAs (reference splits, earlier waves) may already have lowered this query's minimum
        float run_min = (q < m) ? ord2f(approx_min[q]) : inf_f();
        float thresh = run_min + my_band;
        const uint32_t lane_base = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(half * SN);
        if (TS && team == 0) {
            // this thread's row of the query image -> its TMEM lane, 8 columns (two 16-byte chunks) per K step
            const unsigned char* qrow = qimage + (size_t)blockIdx.x * A_BYTES;
            const uint32_t a_lane = tmem_base + ((uint32_t)(lq * 32) << 16) + A_COL0 + (uint32_t)(half * STEPS * 8);
#pragma unroll
            for (int st = 0; st < STEPS; ++st) {
                const uint4 lo = __ldg(reinterpret_cast<const uint4*>(qrow + image_chunk_at(T_BM, KB, row, 2 * st)));
                const uint4 hi = __ldg(reinterpret_cast<const uint4*>(qrow + image_chunk_at(T_BM, KB, row, 2 * st + 1)));
                tmem_st8(a_lane + (uint32_t)(st * 8), lo, hi);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
        // One TMEM chunk in flight per warp: with four epilogue warps per scheduler the load latency
        // of one warp is covered by the reductions of the other three.  (A register double buffer
        // measured no faster at k = 3 and slower at k = 128, where the box runs at its power cap.)
        // a 64-column unit (SUB = 2) is read with ONE 64-column load; 128-column units 32 columns at a time
        constexpr bool LD64 = NNS_T_LD64 < 0 ? (CPU == 2) : (NNS_T_LD64 != 0);
        uint32_t v[LD64 ? 64 : 32];
#if NNS_T_EXPERIMENT >= 2
#pragma unroll
        for (int i = 0; i < (LD64 ? 64 : 32); ++i) v[i] = 0x7f800000u;
#define tmem_ld32(a, b) ((void)0)
#define tmem_ld64(a, b) ((void)0)
#endif
        // reduce one 32-column chunk, emit it as a candidate when it is within the band
        auto reduce_chunk = [&](const uint32_t (&cur)[32], const int unit32) {
#if NNS_T_EXPERIMENT >= 1
            const float cm = fminf(__uint_as_float(cur[0]), __uint_as_float(cur[31]));
#else
            // four independent FMNMX3 chains (depth 4) + a 2-level combine
            float c0 = min3(__uint_as_float(cur[0]), __uint_as_float(cur[1]), __uint_as_float(cur[2]));
            float c1 = min3(__uint_as_float(cur[8]), __uint_as_float(cur[9]), __uint_as_float(cur[10]));
            float c2 = min3(__uint_as_float(cur[16]), __uint_as_float(cur[17]), __uint_as_float(cur[18]));
            float c3 = min3(__uint_as_float(cur[24]), __uint_as_float(cur[25]), __uint_as_float(cur[26]));
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                c0 = min3(c0, __uint_as_float(cur[3 + 2 * j]), __uint_as_float(cur[4 + 2 * j]));
                c1 = min3(c1, __uint_as_float(cur[11 + 2 * j]), __uint_as_float(cur[12 + 2 * j]));
                c2 = min3(c2, __uint_as_float(cur[19 + 2 * j]), __uint_as_float(cur[20 + 2 * j]));
                c3 = min3(c3, __uint_as_float(cur[27 + 2 * j]), __uint_as_float(cur[28 + 2 * j]));
            }
            c0 = fminf(c0, __uint_as_float(cur[7]));
            c1 = fminf(c1, __uint_as_float(cur[15]));
            c2 = fminf(c2, __uint_as_float(cur[23]));
            c3 = fminf(c3, __uint_as_float(cur[31]));
            const float cm = fminf(min3(c0, c1, c2), c3);
#endif
            // candidates have 32-reference granularity (one TMEM chunk): 4x less to re-score than a tile
            if (NNS_T_EXPERIMENT < 2 && cm <= thresh) {
                TensorCand cnd;
                cnd.q = (int)q; cnd.unit = unit32; cnd.smin = cm;
                cand_emit(cb, s_cand_count, cta, cnd);
                if (cm < run_min && !cb.fixed_threshold) {
                    run_min = cm;
                    thresh = run_min + my_band;
                    atomicMin(approx_min + q, f2ord(run_min));
                }
            }
        };
        // ---- pipelined epilogue for 64-reference units (SUB = 2) ----
        // ncu source view of the one-load-per-unit loop on C2 (profiles/r2_ncu_c2_screen.txt): a quarter of an
        // epilogue warp's time went to branch resolution (three taken branches per unit: the loop back edge and
        // the two jumps over the candidate path), 16 % to waiting for its TMEM load, 14 % to waiting for the next
        // accumulator.  Here a unit is read as two 32-column halves into two register sets that rotate, so that
        // a load is in flight under every 32-value reduction; the loop body covers two units, and the four chunk
        // minima are tested against the threshold ONCE (a threshold that is one unit stale only admits more
        // candidates), so a trip has two taken branches instead of six.
        if constexpr (F16) {
            // ---- 16-bit accumulators: 64 columns per load (32 registers), two loads (128 columns) per loop trip and ONE
            // candidate test per trip; 15 HMNMX2 reduce a 32-reference chunk to one register holding two partial minima ----
            static_assert(SN % 64 == 0, "F16 epilogue reads 64 columns at a time");
            constexpr int GPU_ = SN / 64;            // 64-column groups per accumulator unit
            constexpr int UPT = GPU_ == 1 ? 2 : 1;   // units per trip
            uint32_t w[32];
            const int nunits = nt * SUB;
            const float uq = (q < m) ? __ldg(qscale + q) : 1.0f, inv_uq = 1.0f / uq;  // powers of two
            float thresh_s = thresh * uq;            // the threshold in accumulator units
            uint32_t tq2 = 0u;                       // EN: (t, t) as F16, t = u_q / s^2 -- the weight of s^2 |r'|^2 in this row
            if constexpr (EN) {
                const float s = reinterpret_cast<const float*>(mode_word)[THDR_SCALE - THDR_MODE];
                const unsigned short th = __half_as_ushort(__float2half_rn(uq / (s * s)));
                tq2 = (uint32_t)th | ((uint32_t)th << 16);
            }
            auto hmin16 = [&](const int o) -> uint32_t {
                uint32_t c0 = hmin2(w[o + 0], w[o + 1]), c1 = hmin2(w[o + 2], w[o + 3]), c2 = hmin2(w[o + 4], w[o + 5]), c3 = hmin2(w[o + 6], w[o + 7]);
                c0 = hmin2(c0, w[o + 8]); c1 = hmin2(c1, w[o + 9]); c2 = hmin2(c2, w[o + 10]); c3 = hmin2(c3, w[o + 11]);
                c0 = hmin2(c0, w[o + 12]); c1 = hmin2(c1, w[o + 13]); c2 = hmin2(c2, w[o + 14]); c3 = hmin2(c3, w[o + 15]);
                return hmin2(hmin2(c0, c1), hmin2(c2, c3));
            };
            auto emit_h = [&](const uint32_t mh, const int unit32) {  // rare path
                const float cm = hmin2_to_float(mh) * inv_uq;
                if (cm <= thresh) {
                    TensorCand cnd;
                    cnd.q = (int)q; cnd.unit = unit32; cnd.smin = cm;
                    cand_emit(cb, s_cand_count, cta, cnd);
                    if (cm < run_min && !cb.fixed_threshold) {
                        run_min = cm;
                        thresh = run_min + my_band;
                        thresh_s = thresh * uq;
                        atomicMin(approx_min + q, f2ord(run_min));
                    }
                }
            };
            int nb = team;                           // u % NB of the next unit of this team (team < TEAMS <= NB)
            uint32_t par = 0;                        // (u / NB) & 1
#pragma unroll 1
            for (int u = team; u < nunits; u += UPT * TEAMS) {
                uint32_t mm[4];
#pragma unroll
                for (int j = 0; j < UPT; ++j) {
                    const int uj = u + j * TEAMS;
                    if (j > 0 && uj >= nunits) {
#pragma unroll
                        for (int gq = 0; gq < 2 * GPU_; ++gq) mm[2 * GPU_ * j + gq] = 0x7c007c00u;  // +INF, +INF
                        continue;
                    }
                    const int bar = nb;
                    const uint32_t bpar = par;
                    const int buf = (NB == NBUF) ? nb : (nb >= NBUF ? nb - NBUF : nb);
                    nb += TEAMS;
                    if (nb >= NB) { nb -= NB; par ^= 1u; }
                    if (lane == 0) T_TRACE(24 + (e & 7), uj);   // [24..31] started waiting (trace rows are per warp of the unit's team)
                    mbar_wait_hot<SPIN>(acc_full + 8 * bar, bpar);
                    if (lane == 0) T_TRACE(4 + (e & 7), uj);    // [4..11] accumulator ready
                    tc_fence_after();
#pragma unroll
                    for (int gq = 0; gq < GPU_; ++gq) {
                        tmem_ld64_pack16(lane_base + (uint32_t)(buf * 2 * SN + gq * 64), w);
                        uint4 nv[EN ? 8 : 1];
                        if constexpr (EN) {  // the 64 norms of these columns: issued under the TMEM load
                            const uint4* np = reinterpret_cast<const uint4*>(rnorm + ((size_t)(t0 + uj / SUB) * T_BN + (size_t)((uj % SUB) * SN + gq * 64)));
#pragma unroll
                            for (int i4 = 0; i4 < 8; ++i4) nv[i4] = __ldg(np + i4);
                        }
                        tmem_ld_wait();
                        if constexpr (EN) {
#pragma unroll
                            for (int i4 = 0; i4 < 8; ++i4) {
                                w[4 * i4 + 0] = hfma2(tq2, nv[i4].x, w[4 * i4 + 0]);
                                w[4 * i4 + 1] = hfma2(tq2, nv[i4].y, w[4 * i4 + 1]);
                                w[4 * i4 + 2] = hfma2(tq2, nv[i4].z, w[4 * i4 + 2]);
                                w[4 * i4 + 3] = hfma2(tq2, nv[i4].w, w[4 * i4 + 3]);
                            }
                        }
                        if (gq == GPU_ - 1) {  // every TMEM read of this warp for the unit has completed
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(acc_empty + 8 * bar);
                            if (lane == 0) T_TRACE(12 + (e & 7), uj);   // [12..19] released
                        }
                        mm[2 * (GPU_ * j + gq)] = hmin16(0);
                        mm[2 * (GPU_ * j + gq) + 1] = hmin16(16);
                    }
                }
                const float all = hmin2_to_float(hmin2(hmin2(mm[0], mm[1]), hmin2(mm[2], mm[3])));
                if (all <= thresh_s) {
#pragma unroll
                    for (int j = 0; j < UPT; ++j) {
                        const int uj = u + j * TEAMS;
                        if (j > 0 && uj >= nunits) continue;
                        const int unit0 = t0 * (T_BN / 32) + uj * CPU;
#pragma unroll
                        for (int c = 0; c < 2 * GPU_; ++c) emit_h(mm[2 * GPU_ * j + c], unit0 + c);
                    }
                }
            }
        } else if constexpr (NNS_T_PIPE == 1 && CPU == 2 && NNS_T_EXPERIMENT == 0) {
            uint32_t va[32], vb[32];
            const int nunits = nt * SUB;
            auto min32 = [&](const uint32_t (&cur)[32]) -> float {
                float c0 = min3(__uint_as_float(cur[0]), __uint_as_float(cur[1]), __uint_as_float(cur[2]));
                float c1 = min3(__uint_as_float(cur[8]), __uint_as_float(cur[9]), __uint_as_float(cur[10]));
                float c2 = min3(__uint_as_float(cur[16]), __uint_as_float(cur[17]), __uint_as_float(cur[18]));
                float c3 = min3(__uint_as_float(cur[24]), __uint_as_float(cur[25]), __uint_as_float(cur[26]));
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    c0 = min3(c0, __uint_as_float(cur[3 + 2 * j]), __uint_as_float(cur[4 + 2 * j]));
                    c1 = min3(c1, __uint_as_float(cur[11 + 2 * j]), __uint_as_float(cur[12 + 2 * j]));
                    c2 = min3(c2, __uint_as_float(cur[19 + 2 * j]), __uint_as_float(cur[20 + 2 * j]));
                    c3 = min3(c3, __uint_as_float(cur[27 + 2 * j]), __uint_as_float(cur[28 + 2 * j]));
                }
                c0 = min3(c0, __uint_as_float(cur[7]), __uint_as_float(cur[15]));
                c2 = min3(c2, __uint_as_float(cur[23]), __uint_as_float(cur[31]));
                return min3(min3(c0, c1, c2), c3, c3);
            };
            auto emit = [&](const float cm, const int unit32) {  // rare path
                if (cm <= thresh) {
                    TensorCand cnd;
                    cnd.q = (int)q; cnd.unit = unit32; cnd.smin = cm;
                    cand_emit(cb, s_cand_count, cta, cnd);
                    if (cm < run_min && !cb.fixed_threshold) {
                        run_min = cm;
                        thresh = run_min + my_band;
                        atomicMin(approx_min + q, f2ord(run_min));
                    }
                }
            };
            auto acquire = [&](const int u) {  // unit u's accumulator is complete
                mbar_wait_hot<SPIN>(acc_full + 8 * (u % NBUF), (uint32_t)((u / NBUF) & 1));
                tc_fence_after();
            };
            auto release = [&](const int u) {  // every TMEM read of this warp for unit u has completed
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + 8 * (u % NBUF));
            };
            auto col0 = [&](const int u) { return lane_base + (uint32_t)((u % NBUF) * 2 * SN); };
            // one unit: on entry va = first half (complete), vb = second half (load in flight);
            // on exit the same for unit un (when it exists)
            auto step = [&](const int u, const int un, float& m0, float& m1) {
                m0 = min32(va);
                tmem_ld_wait_for(vb);
                release(u);
                if (un < nunits) {
                    acquire(un);
                    tmem_ld32(col0(un), va);
                }
                m1 = min32(vb);
                if (un < nunits) {
                    tmem_ld_wait_for(va);
                    tmem_ld32(col0(un) + 32, vb);
                }
            };
            int u = (T_TEAMS == 2 ? team : 0);
            if (u < nunits) {
                acquire(u);
                tmem_ld32(col0(u), va);
                tmem_ld_wait_for(va);
                tmem_ld32(col0(u) + 32, vb);
            }
#pragma unroll 1
            for (; u < nunits; u += 2 * T_TEAMS) {
                const int u1 = u + T_TEAMS, u2 = u + 2 * T_TEAMS;
                float m0, m1, m2 = inf_f(), m3 = inf_f();
                step(u, u1, m0, m1);
                if (u1 < nunits) step(u1, u2, m2, m3);
                if (min3(fminf(m0, m1), m2, m3) <= thresh) {
                    const int unit0 = t0 * (T_BN / 32) + u * CPU;
                    emit(m0, unit0);
                    emit(m1, unit0 + 1);
                    if (u1 < nunits) {
                        emit(m2, unit0 + T_TEAMS * CPU);
                        emit(m3, unit0 + T_TEAMS * CPU + 1);
                    }
                }
            }
        } else if constexpr (NNS_T_PIPE == 3 && CPU == 2 && NNS_T_EXPERIMENT == 0 && T_TEAMS == 1) {
            // ---- variant (one team of 8 warps, up to 168 registers per thread): two rotating 64-column register sets,
            // the load of the next unit in flight under the reduction of the current one ----
            uint32_t va[64], vb[64];
            const int nunits = nt * SUB;
            auto min64 = [&](const uint32_t (&w)[64], float& m0, float& m1) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int o = 32 * hh;
                    float c0 = min3(__uint_as_float(w[o + 0]), __uint_as_float(w[o + 1]), __uint_as_float(w[o + 2]));
                    float c1 = min3(__uint_as_float(w[o + 8]), __uint_as_float(w[o + 9]), __uint_as_float(w[o + 10]));
                    float c2 = min3(__uint_as_float(w[o + 16]), __uint_as_float(w[o + 17]), __uint_as_float(w[o + 18]));
                    float c3 = min3(__uint_as_float(w[o + 24]), __uint_as_float(w[o + 25]), __uint_as_float(w[o + 26]));
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        c0 = min3(c0, __uint_as_float(w[o + 3 + 2 * j]), __uint_as_float(w[o + 4 + 2 * j]));
                        c1 = min3(c1, __uint_as_float(w[o + 11 + 2 * j]), __uint_as_float(w[o + 12 + 2 * j]));
                        c2 = min3(c2, __uint_as_float(w[o + 19 + 2 * j]), __uint_as_float(w[o + 20 + 2 * j]));
                        c3 = min3(c3, __uint_as_float(w[o + 27 + 2 * j]), __uint_as_float(w[o + 28 + 2 * j]));
                    }
                    c0 = min3(c0, __uint_as_float(w[o + 7]), __uint_as_float(w[o + 15]));
                    c2 = min3(c2, __uint_as_float(w[o + 23]), __uint_as_float(w[o + 31]));
                    (hh == 0 ? m0 : m1) = min3(min3(c0, c1, c2), c3, c3);
                }
            };
            auto emit3 = [&](const float cm, const int unit32) {
                if (cm <= thresh) {
                    TensorCand cnd;
                    cnd.q = (int)q; cnd.unit = unit32; cnd.smin = cm;
                    cand_emit(cb, s_cand_count, cta, cnd);
                    if (cm < run_min && !cb.fixed_threshold) {
                        run_min = cm;
                        thresh = run_min + my_band;
                        atomicMin(approx_min + q, f2ord(run_min));
                    }
                }
            };
            auto acquire3 = [&](const int u) {
                mbar_wait_hot<SPIN>(acc_full + 8 * (u % NBUF), (uint32_t)((u / NBUF) & 1));
                tc_fence_after();
            };
            auto release3 = [&](const int u) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + 8 * (u % NBUF));
            };
            auto col3 = [&](const int u) { return lane_base + (uint32_t)((u % NBUF) * 2 * SN); };
            if (nunits > 0) { acquire3(0); tmem_ld64(col3(0), va); }
#pragma unroll 1
            for (int u = 0; u < nunits; u += 2) {
                float m0, m1, m2 = inf_f(), m3 = inf_f();
                tmem_ld_wait_for64(va);
                release3(u);
                if (u + 1 < nunits) { acquire3(u + 1); tmem_ld64(col3(u + 1), vb); }
                min64(va, m0, m1);
                if (u + 1 < nunits) {
                    tmem_ld_wait_for64(vb);
                    release3(u + 1);
                    if (u + 2 < nunits) { acquire3(u + 2); tmem_ld64(col3(u + 2), va); }
                    min64(vb, m2, m3);
                }
                if (min3(fminf(m0, m1), m2, m3) <= thresh) {
                    const int unit0 = t0 * (T_BN / 32) + u * CPU;
                    emit3(m0, unit0);
                    emit3(m1, unit0 + 1);
                    if (u + 1 < nunits) { emit3(m2, unit0 + 2); emit3(m3, unit0 + 3); }
                }
            }
        } else if constexpr (NNS_T_PIPE == 2 && CPU == 2 && NNS_T_EXPERIMENT == 0) {
            // ---- variant: one 64-column load per unit as in the default path, but two units per loop trip and ONE
            // candidate test per two units (fewer taken branches; no change to the TMEM access pattern) ----
            uint32_t w[64];
            const int nunits = nt * SUB;
            auto min32w = [&](const int o) -> float {
                float c0 = min3(__uint_as_float(w[o + 0]), __uint_as_float(w[o + 1]), __uint_as_float(w[o + 2]));
                float c1 = min3(__uint_as_float(w[o + 8]), __uint_as_float(w[o + 9]), __uint_as_float(w[o + 10]));
                float c2 = min3(__uint_as_float(w[o + 16]), __uint_as_float(w[o + 17]), __uint_as_float(w[o + 18]));
                float c3 = min3(__uint_as_float(w[o + 24]), __uint_as_float(w[o + 25]), __uint_as_float(w[o + 26]));
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    c0 = min3(c0, __uint_as_float(w[o + 3 + 2 * j]), __uint_as_float(w[o + 4 + 2 * j]));
                    c1 = min3(c1, __uint_as_float(w[o + 11 + 2 * j]), __uint_as_float(w[o + 12 + 2 * j]));
                    c2 = min3(c2, __uint_as_float(w[o + 19 + 2 * j]), __uint_as_float(w[o + 20 + 2 * j]));
                    c3 = min3(c3, __uint_as_float(w[o + 27 + 2 * j]), __uint_as_float(w[o + 28 + 2 * j]));
                }
                c0 = min3(c0, __uint_as_float(w[o + 7]), __uint_as_float(w[o + 15]));
                c2 = min3(c2, __uint_as_float(w[o + 23]), __uint_as_float(w[o + 31]));
                return min3(min3(c0, c1, c2), c3, c3);
            };
            auto emit2 = [&](const float cm, const int unit32) {
                if (cm <= thresh) {
                    TensorCand cnd;
                    cnd.q = (int)q; cnd.unit = unit32; cnd.smin = cm;
                    cand_emit(cb, s_cand_count, cta, cnd);
                    if (cm < run_min && !cb.fixed_threshold) {
                        run_min = cm;
                        thresh = run_min + my_band;
                        atomicMin(approx_min + q, f2ord(run_min));
                    }
                }
            };
            auto unit = [&](const int u, float& m0, float& m1) {
                const int buf = u % NBUF;
                if (lane == 0) T_TRACE(24 + (e & 7), u);   // [24..31] started waiting (trace rows are per warp of the unit's team)
                mbar_wait_hot<SPIN>(acc_full + 8 * buf, (uint32_t)((u / NBUF) & 1));
                if (lane == 0) T_TRACE(4 + (e & 7), u);    // [4..11] accumulator ready
                tc_fence_after();
                tmem_ld64(lane_base + (uint32_t)(buf * 2 * SN), w);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + 8 * buf);
                if (lane == 0) T_TRACE(12 + (e & 7), u);   // [12..19] released
                m0 = min32w(0);
                m1 = min32w(32);
            };
            constexpr int TRIP = NNS_T_TRIP;  // units per loop trip and per candidate test
#pragma unroll 1
            for (int u = (T_TEAMS == 2 ? team : 0); u < nunits; u += TRIP * T_TEAMS) {
                float mm[2 * TRIP];
                float all = inf_f();
#pragma unroll
                for (int j = 0; j < TRIP; ++j) {
                    mm[2 * j] = mm[2 * j + 1] = inf_f();
                    if (j == 0 || u + j * T_TEAMS < nunits) unit(u + j * T_TEAMS, mm[2 * j], mm[2 * j + 1]);
                    all = min3(all, mm[2 * j], mm[2 * j + 1]);
                }
                if (all <= thresh) {
                    const int unit0 = t0 * (T_BN / 32) + u * CPU;
#pragma unroll
                    for (int j = 0; j < TRIP; ++j)
                        if (j == 0 || u + j * T_TEAMS < nunits) {
                            emit2(mm[2 * j], unit0 + j * T_TEAMS * CPU);
                            emit2(mm[2 * j + 1], unit0 + j * T_TEAMS * CPU + 1);
                        }
                }
            }
        } else
        // team i reduces the units u = i, i + 2, ...: buffer u % NBUF (its own parity), 32-reference
        // candidate units t0 * 4 + u * CPU + c
        for (int u = (T_TEAMS == 2 ? team : 0); u < nt * SUB; u += T_TEAMS) {
            const int buf = u % NBUF;
            const uint32_t taddr = lane_base + (uint32_t)(buf * 2 * SN);
            const int unit0 = t0 * (T_BN / 32) + u * CPU;
            mbar_wait_hot<SPIN>(acc_full + 8 * buf, (uint32_t)((u / NBUF) & 1));
            if (lane == 0) T_TRACE(4 + e, u);
            tc_fence_after();
            if constexpr (LD64) {
                static_assert(!LD64 || CPU % 2 == 0, "64-column loads need an even number of chunks per unit");
#pragma unroll 1
                for (int c = 0; c < CPU; c += 2) {
                    tmem_ld64(taddr + c * 32, *reinterpret_cast<uint32_t(*)[64]>(&v[0]));
                    tmem_ld_wait();
                    if (c + 2 == CPU) {
                        // every TMEM read of this warp for this unit has completed: hand the accumulator
                        // back to the MMA issuer before reducing
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty + 8 * buf);
                        if (lane == 0) T_TRACE(4 + 16 + e, u);
                    }
                    reduce_chunk(*reinterpret_cast<const uint32_t(*)[32]>(&v[0]), unit0 + c);
                    reduce_chunk(*reinterpret_cast<const uint32_t(*)[32]>(&v[LD64 ? 32 : 0]), unit0 + c + 1);
                }
            } else {
                // The chunk loop is deliberately NOT unrolled: unrolled, ptxas hoists all loads to the top
                // of the unit and spills the loop invariants to make room for the destination registers.
#pragma unroll 1
                for (int c = 0; c < CPU; ++c) {
                    tmem_ld32(taddr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                    tmem_ld_wait();
                    if (c + 1 == CPU) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty + 8 * buf);
                        if (lane == 0) T_TRACE(4 + 16 + e, u);
                    }
                    reduce_chunk(*reinterpret_cast<const uint32_t(*)[32]>(&v[0]), unit0 + c);
                }
            }
        }
#if NNS_T_EXPERIMENT >= 2
#undef tmem_ld32
#undef tmem_ld64
#endif
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

// ---------------------------------------------------------------------------------------------
// exact re-score: candidates are filtered against the FINAL approximate minimum by one lane each,
// the survivors are evaluated by the whole warp, one lane per reference of the 32-reference unit
// ---------------------------------------------------------------------------------------------
template <bool EXACT>
__global__ void __launch_bounds__(256)
tensor_rescore_kernel(const float* __restrict__ queries, const int k, const float* __restrict__ blocks,
                      const int index_base, const CandBuf cb, const float* __restrict__ band,
                      const unsigned* __restrict__ approx_min, u64* __restrict__ keys)
{
    const unsigned common = *cb.common_count;
    if (common > cb.common_cap || cb.status[1] != 0u) {  // the FP32 kernel launched after the last batch takes over
        if (blockIdx.x == 0 && threadIdx.x == 0) cb.status[1] = 1u;
        return;
    }
    const int lane = (int)(threadIdx.x & 31);
    const unsigned wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const unsigned gpr = cb.region_cap / 32;                   // 32-record groups per CTA region
    const unsigned region_groups = cb.n_ctas * gpr;
    const unsigned groups = region_groups + (common + 31) / 32;
    for (unsigned gi = wid; gi < groups; gi += nw) {
        size_t first;
        unsigned count;
        if (gi < region_groups) {
            const unsigned cta = gi / gpr, s0 = (gi - cta * gpr) * 32;
            const unsigned used = cb.cta_count[cta];
            if (s0 >= used) continue;
            first = (size_t)cta * cb.region_cap + s0;
            count = min(32u, used - s0);
        } else {
            const unsigned s0 = (gi - region_groups) * 32;
            first = (size_t)cb.n_ctas * cb.region_cap + s0;
            count = min(32u, common - s0);
        }
        TensorCand mine;
        mine.q = 0; mine.unit = 0; mine.smin = 0.0f;
        bool live = false;
        if ((unsigned)lane < count) {
            mine = cb.rec[first + lane];
            live = mine.smin <= ord2f(approx_min[mine.q]) + band[mine.q];
        }
        unsigned mask = __ballot_sync(0xffffffffu, live);
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const int cq = __shfl_sync(0xffffffffu, mine.q, src);
            const int unit = __shfl_sync(0xffffffffu, mine.unit, src);
            const int jl = unit * 32 + lane;  // reference index within this index
            const float* col = blocks + (size_t)(jl >> 7) * (k + 1) * LB + (jl & (LB - 1));
            const float* qp = queries + (size_t)cq * k;
            float d = 0.0f;
            for (int tb = 0; tb < k; tb += 32) {
                const float qv = (tb + lane < k) ? __ldg(qp + tb + lane) : 0.0f;
                const int te = min(32, k - tb);
#pragma unroll 8
                for (int tt = 0; tt < te; ++tt) {
                    const float qt = __shfl_sync(0xffffffffu, qv, tt);
                    const float e = qt - __ldg(col + (size_t)(tb + tt) * LB);
                    d = EXACT ? __fadd_rn(d, __fmul_rn(e, e)) : __fmaf_rn(e, e, d);
                }
            }
            u64 key = (d < inf_f()) ? pack_key(d, index_base + jl) : KEY_INIT;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const u64 o = __shfl_xor_sync(0xffffffffu, key, off);
                key = o < key ? o : key;
            }
            if (lane == 0 && key < KEY_INIT) atomicMin(keys + cq, key);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int tensor_kp(int k)
{
    const TensorGeom g = tensor_geom(k);
    return g.KB * 64 + g.KS * 16;
}

size_t tensor_image_bytes_per_block(int k)
{
    const TensorGeom g = tensor_geom(k);
    return image_bytes(T_BN, g.KB, g.KS);
}

size_t tensor_section_floats(int k, int n)
{
    if (k < 1 || k > TENSOR_MAX_K || n <= 0) return 0;
    const size_t nblocks = (size_t)((n + LB - 1) / LB);
    return (size_t)TENSOR_HDR_FLOATS + nblocks * tensor_image_bytes_per_block(k) / 4;
}

cudaError_t tensor_section_init(int k, int n, const float* d_blocks, float* d_section, const TensorCentre* fixed,
                                cudaStream_t st)
{
    if (tensor_section_floats(k, n) == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_section, 0, TENSOR_HDR_FLOATS * sizeof(float), st);  // mode word: 0 = split
    if (e != cudaSuccess) return e;
    TensorCentre c{};
    if (fixed) {
        c = *fixed;
    } else {
        const int nblocks = (n + LB - 1) / LB;
        const int stride = (nblocks + TENSOR_CENTRE_BLOCKS - 1) / TENSOR_CENTRE_BLOCKS;
        tensor_colsum_kernel<<<(nblocks + stride - 1) / stride, 128, 0, st>>>(d_blocks, n, k, stride, d_section,
                                                                              reinterpret_cast<unsigned*>(d_section) + THDR_MAX + 2);
    }
    tensor_centre_kernel<<<1, 512, 0, st>>>(d_section, k, c, fixed ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t tensor_image_build(int k, int cn, const float* d_blocks_part, float* d_hdr, int max_word, int flag_word,
                               const ImageDsts& dst, cudaStream_t st, int write_blocks)
{
    const int nb = write_blocks > 0 ? write_blocks : (cn + LB - 1) / LB;
    if (nb <= 0 || k < 1 || k > TENSOR_MAX_K) return cudaSuccess;
    const TensorGeom g = tensor_geom(k);  // slices of a shared index: always the default (split) layout
    tensor_ref_image_kernel<<<nb, 128, 0, st>>>(d_blocks_part, cn, k, g, d_hdr, max_word, flag_word, dst, nullptr, 0u, 0);
    return cudaGetLastError();
}

// ---- precision mode (TENSOR_PLAIN_MIN_K <= k <= 128) ----
// k <= TENSOR_SPLIT_MAX_K: split-precision BF16 or plain F16, decided per index by a probe; above: plain F16 (a section
// built in slices by several GPUs / processes keeps the default BF16 layout: its scale would have to be agreed first)
constexpr int TENSOR_PROBE_SAMPLES = 128;
bool tensor_has_modes(int k) { return k >= TENSOR_PLAIN_MIN_K && k <= 128; }

// sample s = reference number s * stride, copied out of the tiled-SoA blocks as an AoS query
__global__ void tensor_probe_gather_kernel(const float* __restrict__ blocks, const int n, const int k, const long long stride,
                                           float* __restrict__ out)
{
    const long long j = (long long)blockIdx.x * stride;
    if (j >= n) return;
    for (int t = threadIdx.x; t < k; t += blockDim.x)
        out[(size_t)blockIdx.x * k + t] = blocks[(size_t)(j >> 7) * (k + 1) * LB + (size_t)t * LB + (j & (LB - 1))];
}

// max |r - c|^2 over the strided block sample the centre came from (the images, and with them the exact
// maximum, do not exist yet when the mode is chosen; an estimate is enough for a heuristic): hdr[THDR_MAX + 4]
__global__ void __launch_bounds__(128)
tensor_rmax_sample_kernel(const float* __restrict__ blocks, const int n, const int k, const int stride, float* __restrict__ hdr)
{
    const long long b = (long long)blockIdx.x * stride;
    const long long j = b * LB + threadIdx.x;
    float rn = 0.0f;
    if (j < n) {
        const float* col = blocks + (size_t)b * (k + 1) * LB + threadIdx.x;
        for (int t = 0; t < k; ++t) {
            const float x = __ldg(col + (size_t)t * LB) - hdr[t];
            rn = __fmaf_rn(x, x, rn);
        }
    }
    unsigned bits = (rn < inf_f()) ? __float_as_uint(rn) : 0u;
    bits = __reduce_max_sync(0xffffffffu, bits);
    if ((threadIdx.x & 31) == 0 && bits) atomicMax(reinterpret_cast<unsigned*>(hdr) + THDR_MAX + 4, bits);
}

// One warp.  Fixes the F16 reference scale from the sampled radius and chooses the mode.  With a probe, for every
// sample: d1 = distance to its nearest OTHER reference (second entry of its 2-NN list; 0 for duplicated points),
// E = the F16 error bound at that point; (1 + 2E/d1)^(k/2) estimates how many references fall inside the plain band
// (locally uniform density in k dimensions -- an over-estimate for data of lower intrinsic dimension, which errs
// towards split precision).  F16 is chosen when that estimate is <= 32 for three quarters of the samples.
__global__ void tensor_mode_kernel(float* __restrict__ hdr, const float* __restrict__ samples, const u64* __restrict__ keys2,
                                   const int count, const int k, const int KP_plain, const int probe)
{
    const int lane = (int)threadIdx.x;
    const float rmax = sqrtf(__uint_as_float(reinterpret_cast<const unsigned*>(hdr)[THDR_MAX + 4]));  // sampled max |r'|
    const float s = tensor_f16_ref_scale(rmax);
    int ok = 0;
    for (int smp = lane; probe && smp < count; smp += 32) {
        float qn = 0.0f;
        for (int t = 0; t < k; ++t) {
            const float x = samples[(size_t)smp * k + t] - hdr[t];
            qn = __fmaf_rn(x, x, qn);
        }
        const float a = sqrtf(qn);
        const float tq = tensor_f16_query_scale(s * a, s * rmax);
        const float E = tq > 0.0f ? tensor_error_bound_f16(KP_plain, k, a, rmax, s, tq) : inf_f();
        const float d1 = __uint_as_float((unsigned)(keys2[(size_t)smp * 2 + 1] >> 32));
        const float est = (d1 > 0.0f && d1 < inf_f()) ? expf(0.5f * (float)k * log1pf(2.0f * E / d1)) : inf_f();
        ok += (est <= 32.0f) ? 1 : 0;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ok += __shfl_xor_sync(0xffffffffu, ok, off);
    if (lane == 0) {
        hdr[THDR_SCALE] = s;
        reinterpret_cast<unsigned*>(hdr)[THDR_MODE] = (rmax <= 1e15f && (!probe || 4 * ok >= 3 * count)) ? TMODE_F16 : TMODE_DEFAULT;
    }
}

cudaError_t tensor_index_build(int k, int n, const float* d_header, const float* d_blocks, float* d_section, cudaStream_t st)
{
    if (tensor_section_floats(k, n) == 0) return cudaSuccess;
    cudaError_t e = tensor_section_init(k, n, d_blocks, d_section, nullptr, st);
    if (e != cudaSuccess) return e;
    ImageDsts dst{};
    dst.p[0] = reinterpret_cast<unsigned char*>(d_section + TENSOR_HDR_FLOATS);
    dst.count = 1;
    const int nb = (n + LB - 1) / LB;
    const unsigned* mode_word = reinterpret_cast<const unsigned*>(d_section) + THDR_MODE;
    const bool probe = k <= TENSOR_SPLIT_MAX_K;  // below, split precision is the alternative and the data decides
    const bool modes = tensor_has_modes(k) && (!probe || (n >= 4 * TENSOR_PROBE_SAMPLES && d_header != nullptr));
    const TensorGeom gf = tensor_geom(k, true);
    if (modes) {
        const int stride = (nb + TENSOR_CENTRE_BLOCKS - 1) / TENSOR_CENTRE_BLOCKS;
        tensor_rmax_sample_kernel<<<(nb + stride - 1) / stride, 128, 0, st>>>(d_blocks, n, k, stride, d_section);
        if (probe) {
            // the 2 nearest references of TENSOR_PROBE_SAMPLES sample points (the exact FP32 top-K kernel)
            const int S = TENSOR_PROBE_SAMPLES;
            const int splits = topk_choose_splits(S, n, 148);
            const size_t off_keys = ((size_t)S * k * sizeof(float) + 255) & ~(size_t)255;
            const size_t off_scr = off_keys + (((size_t)S * 2 * sizeof(u64) + 255) & ~(size_t)255);
            unsigned char* tmp = nullptr;
            e = cudaMallocAsync((void**)&tmp, off_scr + topk_scratch_bytes(S, 2, splits), st);
            if (e != cudaSuccess) return e;
            float* samples = reinterpret_cast<float*>(tmp);
            u64* keys2 = reinterpret_cast<u64*>(tmp + off_keys);
            tensor_probe_gather_kernel<<<S, 32, 0, st>>>(d_blocks, n, k, (long long)n / S, samples);
            e = launch_keys_init(keys2, 2 * S, st);
            if (e == cudaSuccess) e = topk_search_launch(k, S, n, 2, samples, d_blocks, 0, keys2, reinterpret_cast<u64*>(tmp + off_scr), splits, false, st, nullptr);
            if (e == cudaSuccess) {
                tensor_mode_kernel<<<1, 32, 0, st>>>(d_section, samples, keys2, S, k, gf.KB * 64 + gf.KS * 16, 1);
                e = cudaGetLastError();
            }
            const cudaError_t fe = cudaFreeAsync(tmp, st);
            if (e != cudaSuccess) return e;
            if (fe != cudaSuccess) return fe;
        } else {
            tensor_mode_kernel<<<1, 32, 0, st>>>(d_section, nullptr, nullptr, 0, k, gf.KB * 64 + gf.KS * 16, 0);
        }
    }
    // the image in the layout of the chosen mode (both launches; the one that does not match exits at once)
    tensor_ref_image_kernel<<<nb, 128, 0, st>>>(d_blocks, n, k, tensor_geom(k, false), d_section, THDR_MAX, THDR_FLAGS, dst,
                                                modes ? mode_word : nullptr, TMODE_DEFAULT, 0);
    if (modes)
        tensor_ref_image_kernel<<<nb, 128, 0, st>>>(d_blocks, n, k, gf, d_section, THDR_MAX, THDR_FLAGS, dst, mode_word, TMODE_F16, 1);
    return cudaGetLastError();
}

// reference tiles per TMA stage / ring depth per operand geometry (stage = G tiles <= 36 KiB)
static int tensor_group(const TensorGeom& g) { return g.KB == 0 ? (g.KS == 1 ? 4 : 2) : 1; }
// must match the STAGES arguments of the NNS_SCREEN table in tensor_search()
static int tensor_stages(const TensorGeom& g)
{
    if (g.KB == 0) return 6;
    if (NNS_T_TS == 2) return g.KB == 2 ? 5 : 8;
    if (NNS_T_TS == 1) return g.KB == 2 ? 4 : 8;
    return g.KB == 2 ? 4 : (g.KS == 0 ? 8 : 6);
}

static size_t tensor_smem_bytes(const TensorGeom& g)
{
    const bool ts = (NNS_T_TS == 2 && g.KB > 0) || (NNS_T_TS == 1 && g.KB == 1);  // A lives in TMEM: no shared-memory A tile
    return (ts ? 0 : image_bytes(T_BM, g.KB, g.KS)) + (size_t)tensor_stages(g) * tensor_group(g) * image_bytes(T_BN, g.KB, g.KS) +
           (2 * T_MAX_STAGES + 8) * 8 + 16;
}

template <int KB, int KS, int STAGES, int G, int SUB, int NBUF, bool TS, int ISS, bool F16, int TEAMS = T_TEAMS, bool EN = false>
static cudaError_t tensor_screen_launch(dim3 grid, size_t smem, cudaStream_t st, const unsigned char* qimage, int m,
                                        const unsigned char* rimage, int ntiles, int tps, const float* band, unsigned* amin,
                                        const float* qscale, const CandBuf& cb, const unsigned* mode_word, unsigned my_mode)
{
    cudaError_t e = cudaFuncSetAttribute(tensor_screen_kernel<KB, KS, STAGES, G, SUB, NBUF, TS, ISS, F16, TEAMS, EN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // EN: the norm array follows the tile images of the whole section
    const unsigned short* rnorm = EN ? reinterpret_cast<const unsigned short*>(rimage + (size_t)ntiles * image_bytes(T_BN, KB, KS)) : nullptr;
    tensor_screen_kernel<KB, KS, STAGES, G, SUB, NBUF, TS, ISS, F16, TEAMS, EN><<<grid, 32 * (1 + ISS + TEAMS * T_TEAM_WARPS), smem, st>>>(
        qimage, m, rimage, ntiles, tps, band, amin, qscale, rnorm, cb, mode_word, my_mode);
    return cudaGetLastError();
}

// the screen kernel instantiated for operand geometry g: short contractions (k <= 9, or plain mid-k): SS form,
// 4 buffers of 64 references; longer ones: A in TMEM (NNS_T_TS), 64-reference units in 3 buffers (2 where A
// needs more than 128 columns).  F16: the same shapes with 16-bit operands and accumulators.
template <bool F16>
static cudaError_t tensor_screen_dispatch_t(const TensorGeom& g, dim3 grid, size_t smem, cudaStream_t st, const unsigned char* qimage,
                                            int m, const unsigned char* rimage, int ntiles, int tps, const float* band, unsigned* amin,
                                            const float* qscale, const CandBuf& cb, const unsigned* mode_word, unsigned my_mode)
{
#define NNS_SCREEN(KB_, KS_, ST_, G_, SUB_, NBUF_, TS_) \
    tensor_screen_launch<KB_, KS_, ST_, G_, SUB_, NBUF_, TS_, 1, F16>(grid, smem, st, qimage, m, rimage, ntiles, tps, band, amin, qscale, cb, mode_word, my_mode)
#define NNS_SCREEN_ISS(KB_, KS_, ST_, G_, SUB_, NBUF_, TS_, ISS_) \
    tensor_screen_launch<KB_, KS_, ST_, G_, SUB_, NBUF_, TS_, ISS_, F16>(grid, smem, st, qimage, m, rimage, ntiles, tps, band, amin, qscale, cb, mode_word, my_mode)
    // issuing threads: the FP32-accumulator epilogue needs 87 registers (one 64-column load), which allows 19 warps = 2 issuers;
    // the 16-bit one needs 60, which allows 21 warps = one issuer per accumulator buffer
    constexpr int ISS_ = NNS_T_SUB == 2 ? (F16 ? NNS_T_ISS_F16 : NNS_T_ISS) : 1;
    if constexpr (F16 && NNS_T_SUB == 2 && NNS_T_ISS_F16 == 3) {
        // A in TMEM, three buffers, one issuer per buffer: in the SS form an N = 64 MMA fetches 4 KiB of A and 2 KiB of B
        // from shared memory, 48 clk at 128 B/clk for 32 clk of tensor work (measured: 58 clk per MMA on C3)
        if (g.KB == 0 && g.KS == 1)
            return tensor_screen_launch<0, 1, 6, 4, 2, 3, true, 3, F16, NNS_T_TEAMS_F16>(grid, smem, st, qimage, m, rimage, ntiles, tps, band, amin, qscale, cb, mode_word, my_mode);
        if (g.KB == 0)
            return tensor_screen_launch<0, 2, 6, 2, 2, 3, true, 3, F16, NNS_T_TEAMS_F16>(grid, smem, st, qimage, m, rimage, ntiles, tps, band, amin, qscale, cb, mode_word, my_mode);
    } else {
        if (g.KB == 0 && g.KS == 1) return NNS_SCREEN_ISS(0, 1, 6, 4, NNS_T_SUB, 2 * NNS_T_SUB, false, ISS_);
        if (g.KB == 0) return NNS_SCREEN_ISS(0, 2, 6, 2, NNS_T_SUB, 2 * NNS_T_SUB, false, ISS_);
    }
    if constexpr (F16) {
        if (g.en) {  // k = 62..64 / 126..128: no norm columns, the epilogue adds |r'|^2
            if (g.KB == 1)
                return tensor_screen_launch<1, 0, 8, 1, 2, 3, true, 1, true, T_TEAMS, true>(grid, smem, st, qimage, m, rimage, ntiles, tps, band, amin, qscale, cb, mode_word, my_mode);
            return tensor_screen_launch<2, 0, 4, 1, 1, 2, false, 1, true, T_TEAMS, true>(grid, smem, st, qimage, m, rimage, ntiles, tps, band, amin, qscale, cb, mode_word, my_mode);
        }
    }
#if NNS_T_TS == 1
    // measured on B200 (profiles/r2_tune_ts.txt): A in TMEM + three 64-reference buffers is 3 % faster at 64
    // columns (C3 split: 434 vs 446 ms) and 3 % slower at 144 (C4: 229 vs 223 ms), where the tensor pipe is busy
    // 98 % of the time either way and twice as many MMA instructions only cost issue slots
    if (g.KB == 1 && g.KS == 0) return NNS_SCREEN(1, 0, 8, 1, 2, 3, true);
    if (g.KB == 1) return NNS_SCREEN(1, 1, 8, 1, 2, 3, true);
    if (g.KS == 0) return NNS_SCREEN(2, 0, 4, 1, 1, 2, false);
    return NNS_SCREEN(2, 1, 4, 1, 1, 2, false);
#elif NNS_T_TS == 2
    if (g.KB == 1 && g.KS == 0) return NNS_SCREEN(1, 0, 8, 1, 2, 3, true);
    if (g.KB == 1) return NNS_SCREEN(1, 1, 8, 1, 2, 3, true);
    if (g.KS == 0) return NNS_SCREEN(2, 0, 5, 1, 2, 3, true);
    return NNS_SCREEN(2, 1, 5, 1, 2, 2, true);
#else
    if (g.KB == 1 && g.KS == 0) return NNS_SCREEN(1, 0, 8, 1, 1, 2, false);
    if (g.KB == 1) return NNS_SCREEN(1, 1, 6, 1, 1, 2, false);
    if (g.KS == 0) return NNS_SCREEN(2, 0, 4, 1, 1, 2, false);
    return NNS_SCREEN(2, 1, 4, 1, 1, 2, false);
#endif
#undef NNS_SCREEN
#undef NNS_SCREEN_ISS
}

// One query batch through the screen: query image(s) + screen kernel(s).  For TENSOR_PLAIN_MIN_K <= k <= 128 the
// index may hold either the default BF16 images or the F16 ones (THDR_MODE): both variants of every kernel are
// launched and the one that does not match the header word exits at once -- no host round trip.
struct ScreenSetup {
    TensorGeom g, gf;          // default layout (sizes the scratch) / F16 plain layout
    bool modes, longk;
    const unsigned* mode_word;
    int rows;
    size_t smem, smem_f16;
};
static ScreenSetup screen_setup(int k, const float* d_section)
{
    ScreenSetup s{};
    s.g = tensor_geom(k);
    s.gf = tensor_geom(k, true);
    s.modes = tensor_has_modes(k);
    s.mode_word = s.modes ? reinterpret_cast<const unsigned*>(d_section) + THDR_MODE : nullptr;
    s.longk = s.g.KB > 2;
    s.rows = s.longk ? tensor_longk_rows(s.g.KB) : T_BM;
    // every CTA allocates all 512 TMEM columns: ask for enough shared memory that only one
    // CTA is resident per SM even at KP = 64
    s.smem = std::max(tensor_smem_bytes(s.g), (size_t)120 * 1024);
    s.smem_f16 = std::max(tensor_smem_bytes(s.gf), (size_t)120 * 1024);
    return s;
}
static cudaError_t screen_batch(const ScreenSetup& s, int k, const float* bq, int bm, int bs, int splits, const float* hdr,
                                unsigned char* qimage, float* band, unsigned* amin, float* qscale, const unsigned char* rimage,
                                int nblocks, int tps, const CandBuf& cb, cudaStream_t st, bool images_only = false, bool screen_only = false)
{
    cudaError_t e = cudaSuccess;
    if (!screen_only) {
        tensor_query_image_kernel<<<bs, s.rows, 0, st>>>(bq, bm, k, s.g, hdr, qimage, band, amin, qscale, s.mode_word, TMODE_DEFAULT, 0);
        if (s.modes) tensor_query_image_kernel<<<bs, s.rows, 0, st>>>(bq, bm, k, s.gf, hdr, qimage, band, amin, qscale, s.mode_word, TMODE_F16, 1);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess || images_only) return e;
    dim3 grid((unsigned)bs, (unsigned)splits);
    if (s.longk) return tensor_longk_launch(s.g.KB, grid, st, qimage, bm, rimage, nblocks, tps, band, amin, cb);
    e = tensor_screen_dispatch_t<false>(s.g, grid, s.smem, st, qimage, bm, rimage, nblocks, tps, band, amin, qscale, cb, s.mode_word, TMODE_DEFAULT);
    if (e == cudaSuccess && s.modes)
        e = tensor_screen_dispatch_t<true>(s.gf, grid, s.smem_f16, st, qimage, bm, rimage, nblocks, tps, band, amin, qscale, cb, s.mode_word, TMODE_F16);
    return e;
}

__global__ void tensor_status_init_kernel(unsigned* __restrict__ status, const unsigned cap, unsigned* __restrict__ common, const int nbatches,
                                          const unsigned* __restrict__ mode_word, const unsigned kp_split, const unsigned kp_plain)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { status[0] = 0u; status[1] = 0u; status[2] = cap; status[3] = (mode_word && *mode_word) ? (kp_plain | (*mode_word << 16)) : kp_split; }
    if (i < nbatches) common[i] = 0u;
}

// scratch budget of one search (candidate records dominate): NNS_B200_CAND_MB, default 2048
static size_t tensor_record_budget()
{
    static const size_t b = []() {
        const char* e = getenv("NNS_B200_CAND_MB");
        size_t mb = e ? (size_t)strtoull(e, nullptr, 0) : 2048;
        if (mb < 16) mb = 16;
        return (mb << 20) / sizeof(TensorCand);
    }();
    return b;
}

// Search m queries against the n references of the index section; accumulates into keys.
// The queries are processed in batches of at most 4 waves of CTAs, each batch = query image, screen,
// re-score on the same scratch: the scratch of a search is bounded (tensor_record_budget) however large m
// is, and a batch can afford up to 1024 candidate records per query -- clustered / duplicated data
// (BASELINE config C5: ~400 32-reference units per query inside the 2E band at n = 16.7 M) stays on
// the tensor cores instead of overflowing into the FP32 fallback.
// d_stats: [0] candidates emitted, [1] overflow flag (the caller launches the FP32 fallback kernel with
// it as its enable flag), [2] candidate capacity of one batch, [3] contraction length (BF16 columns) of the
// precision mode the index chose.
cudaError_t tensor_search(int k, int m, int n, const float* d_queries, const float* d_blocks, const float* d_section,
                          int index_base, u64* d_keys, bool exact, int num_sms, cudaStream_t st, cudaMemPool_t pool,
                          int* launches, unsigned* d_stats, bool tiny_candidate_buffer)
{
    const ScreenSetup ss = screen_setup(k, d_section);
    const TensorGeom g = ss.g;                      // default layout (split where it exists): sizes the scratch
    const TensorGeom gp = ss.gf;                    // F16 plain layout, used when the index header says so
    const bool modes = ss.modes;
    const unsigned* mode_word = ss.mode_word;
    const int nblocks = (n + LB - 1) / LB;
    const int rows = ss.rows;                       // query rows per strip / CTA (the K-loop kernel of tensor_longk.cu: 256 or 128)
    const int strips = (m + rows - 1) / rows;
    const float* hdr = d_section;
    const unsigned char* rimage = reinterpret_cast<const unsigned char*>(d_section + TENSOR_HDR_FLOATS);
    if (launches) *launches = 0;

    // query batches: at most 4 waves of CTAs each, balanced, whole waves where possible
    const int max_batch_ctas = 4 * num_sms;
    // reference splits (chosen for one batch of `bs` strips): minimise (waves of one CTA per SM) x (tiles
    // per CTA); every extra split costs each query one more seed candidate, so ties go to fewer splits
    auto choose_splits = [&](int bs, int smin) {
        double best = 1e300;
        int chosen = smin;
        const int smax = std::max(std::min(nblocks, 64), std::min(nblocks, 2 * smin));
        for (int sp = smin; sp <= smax; ++sp) {
            const int t = (nblocks + sp - 1) / sp;
            const int se = (nblocks + t - 1) / t;
            const double waves = (double)(((long long)bs * se + num_sms - 1) / num_sms);
            const double cost = waves * ((double)t + 24.0);  // + per-CTA prologue (A tile, TMEM alloc) in tile units
            if (cost < best * 0.97) { best = cost; chosen = se; }
        }
        return chosen;
    };
    int batch_strips = strips, nbatches = 1;
    if (strips > max_batch_ctas) {
        nbatches = (strips + max_batch_ctas - 1) / max_batch_ctas;
        batch_strips = (strips + nbatches - 1) / nbatches;
        batch_strips = std::min(max_batch_ctas, (batch_strips + num_sms - 1) / num_sms * num_sms);
        nbatches = (strips + batch_strips - 1) / batch_strips;
    }
    int splits = choose_splits(batch_strips, 1);
    // If the data overflows the candidate buffer, only the CTAs already running finish their share of
    // the (then useless) pass -- the rest see the flag and exit.  A job of fewer than four waves is
    // therefore cut into CTAs of at most 16384 tiles; such CTAs of one strip run concurrently, each from
    // an unconverged minimum, so they get first-split sized candidate regions below.
    bool short_ctas = false;
    if ((long long)batch_strips * splits < 4LL * num_sms && (nblocks + splits - 1) / splits > 16384) {
        splits = choose_splits(batch_strips, (nblocks + 16383) / 16384);
        short_ctas = true;
    }
    const int tps = (nblocks + splits - 1) / splits;
    splits = (nblocks + tps - 1) / tps;

    // Candidate capacity of a batch.  Every split of a strip emits its first tiles per query, then one
    // record per running-minimum improvement (~ln tiles) and per unit inside the band.  A CTA region
    // holds at least 64 / splits + 6 (40 for short CTAs) and at most 1024 records per query, as the
    // budget allows; the common spill region another 64 per query.
    CandBuf cb{};
    cb.n_ctas = (unsigned)batch_strips * (unsigned)splits;
    const size_t batch_queries = (size_t)batch_strips * rows;
    if (tiny_candidate_buffer) {  // test hook: forces the overflow -> FP32-kernel fallback
        cb.region_cap = 32;
        cb.common_cap = 32;
    } else {
        const size_t lo = short_ctas ? (size_t)rows * 40 : (((size_t)rows * 64 / splits + (size_t)rows * 6 + 31) & ~(size_t)31);
        const size_t hi = (size_t)rows * 1024;
        size_t common = std::min<size_t>(batch_queries * 64 + 65536, tensor_record_budget() / 4);
        size_t region = (tensor_record_budget() - common) / cb.n_ctas & ~(size_t)31;
        region = std::max(lo, std::min(hi, region));
        cb.region_cap = (unsigned)region;
        cb.common_cap = (unsigned)common;
    }
    const size_t cand_records = (size_t)cb.n_ctas * cb.region_cap + cb.common_cap;

    // stream-ordered scratch of one batch: query image, band, approx_min, per-CTA counts, candidates;
    // + one common-record counter per batch
    const size_t qimg_bytes = (size_t)batch_strips * image_bytes(rows, g.KB, g.KS);
    const size_t off_band = (qimg_bytes + 255) & ~(size_t)255;
    const size_t off_amin = off_band + ((batch_queries * 4 + 255) & ~(size_t)255);
    const size_t off_qs = off_amin + ((batch_queries * 4 + 255) & ~(size_t)255);
    const size_t off_cnt = off_qs + ((batch_queries * 4 + 255) & ~(size_t)255);
    const size_t off_ccnt = off_cnt + (((size_t)nbatches * 4 + 255) & ~(size_t)255);
    const size_t off_cand = off_ccnt + (((size_t)cb.n_ctas * 4 + 255) & ~(size_t)255);
    const size_t total = off_cand + cand_records * sizeof(TensorCand);
    unsigned char* scratch = nullptr;
    cudaError_t e = pool ? cudaMallocFromPoolAsync((void**)&scratch, total, pool, st) : cudaMallocAsync((void**)&scratch, total, st);
    if (e != cudaSuccess) return e;
    float* band = reinterpret_cast<float*>(scratch + off_band);
    unsigned* amin = reinterpret_cast<unsigned*>(scratch + off_amin);
    float* qscale = reinterpret_cast<float*>(scratch + off_qs);
    unsigned* common_counts = reinterpret_cast<unsigned*>(scratch + off_cnt);
    cb.cta_count = reinterpret_cast<unsigned*>(scratch + off_ccnt);
    cb.rec = reinterpret_cast<TensorCand*>(scratch + off_cand);
    cb.status = d_stats;

    tensor_status_init_kernel<<<(nbatches + 255) / 256, 256, 0, st>>>(d_stats, (unsigned)std::min<size_t>(cand_records, 0xffffffffu),
                                                                      common_counts, nbatches, mode_word,
                                                                      (unsigned)(g.KB * 64 + g.KS * 16), (unsigned)(gp.KB * 64 + gp.KS * 16));
    e = cudaGetLastError();
    int nl = 1;
    for (int b = 0; b < nbatches && e == cudaSuccess; ++b) {
        const int s0 = b * batch_strips;
        const int bs = std::min(batch_strips, strips - s0);
        const long long q0 = (long long)s0 * rows;
        const int bm = (int)std::min<long long>((long long)bs * rows, (long long)m - q0);
        const float* bq = d_queries + (size_t)q0 * k;
        cb.common_count = common_counts + b;
        cb.n_ctas = (unsigned)bs * (unsigned)splits;  // the common region follows the regions actually used
        e = screen_batch(ss, k, bq, bm, bs, splits, hdr, scratch, band, amin, qscale, rimage, nblocks, tps, cb, st);
        nl += modes ? 2 : 0;
        if (e != cudaSuccess) break;
        const int rgrid = num_sms * 8;
        if (exact)
            tensor_rescore_kernel<true><<<rgrid, 256, 0, st>>>(bq, k, d_blocks, index_base, cb, band, amin, d_keys + q0);
        else
            tensor_rescore_kernel<false><<<rgrid, 256, 0, st>>>(bq, k, d_blocks, index_base, cb, band, amin, d_keys + q0);
        e = cudaGetLastError();
        nl += 3;
    }
    if (launches) *launches = nl;
    cudaError_t e2 = cudaFreeAsync(scratch, st);
    return e != cudaSuccess ? e : e2;
}

// ---------------------------------------------------------------------------------------------
// K nearest neighbours through the screen
// ---------------------------------------------------------------------------------------------
// The 1-NN screen keeps a unit when its minimum is within the band of the RUNNING minimum.  For the K nearest
// the threshold must not move with the best reference, so it is fixed before the screen runs: an exact FP32
// K-nearest pass (topk_search.cu) over a SAMPLE of the reference blocks (every TOPK_SAMPLE_STRIDE-th block)
// leaves K true distances per query; the K-th, tau, is an upper bound of the final K-th distance, and a
// reference can only belong to the answer if its V0-form distance is <= tau, i.e. if its approximate score
// S~ <= tau - |q'|^2 + E.  The screen runs with that fixed threshold (CandBuf::fixed_threshold) over all blocks,
// the re-score evaluates the candidate units exactly, skips the sampled blocks (already in the lists) and
// appends every reference with d <= tau to a per-query list; topk_merge_var_kernel folds the lists into the
// sorted keys.  Expected list length = K * stride for any data the sample represents; a list or the
// candidate buffer that overflows raises the same device flag as in the 1-NN search and the FP32 K-nearest
// kernel, launched behind the flag over the non-sampled blocks, finishes the job -- no host round trip.
constexpr int TOPK_SAMPLE_STRIDE = 16;

// amin[q] = tau(q) - |q'|^2 (ordered encoding); band[q] stays the 2E the query-image kernel wrote
__global__ void tensor_topk_threshold_kernel(const float* __restrict__ queries, const int m, const int k, const float* __restrict__ hdr,
                                             const u64* __restrict__ keys, const int K, unsigned* __restrict__ approx_min)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    float qn = 0.0f;
    for (int t = 0; t < k; ++t) {
        const float x = __fsub_rn(__ldg(queries + (size_t)q * k + t), hdr[t]);
        qn = __fmaf_rn(x, x, qn);
    }
    const float tau = __uint_as_float((unsigned)(keys[(size_t)q * K + (K - 1)] >> 32));  // +INF while fewer than K are known
    const float thr = tau - qn;  // NaN (INF - INF, NaN queries) compares false everywhere: nothing is screened in, tau = INF decides below
    approx_min[q] = f2ord(thr == thr ? thr : inf_f());
}

template <bool EXACT>
__global__ void __launch_bounds__(256)
tensor_rescore_topk_kernel(const float* __restrict__ queries, const int k, const float* __restrict__ blocks, const int index_base,
                           const CandBuf cb, const u64* __restrict__ keys, const int K, u64* __restrict__ exact,
                           unsigned* __restrict__ exact_count, const unsigned cap, const int stride)
{
    const unsigned common = *cb.common_count;
    if (common > cb.common_cap || cb.status[1] != 0u) {
        if (blockIdx.x == 0 && threadIdx.x == 0) cb.status[1] = 1u;
        return;
    }
    const int lane = (int)(threadIdx.x & 31);
    const unsigned wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const unsigned gpr = cb.region_cap / 32;
    const unsigned region_groups = cb.n_ctas * gpr;
    const unsigned groups = region_groups + (common + 31) / 32;
    for (unsigned gi = wid; gi < groups; gi += nw) {
        size_t first;
        unsigned count;
        if (gi < region_groups) {
            const unsigned cta = gi / gpr, s0 = (gi - cta * gpr) * 32;
            const unsigned used = cb.cta_count[cta];
            if (s0 >= used) continue;
            first = (size_t)cta * cb.region_cap + s0;
            count = min(32u, used - s0);
        } else {
            const unsigned s0 = (gi - region_groups) * 32;
            first = (size_t)cb.n_ctas * cb.region_cap + s0;
            count = min(32u, common - s0);
        }
        TensorCand mine;
        mine.q = 0; mine.unit = 0; mine.smin = 0.0f;
        if ((unsigned)lane < count) mine = cb.rec[first + lane];
        for (unsigned src = 0; src < count; ++src) {
            const int cq = __shfl_sync(0xffffffffu, mine.q, (int)src);
            const int unit = __shfl_sync(0xffffffffu, mine.unit, (int)src);
            const int jl = unit * 32 + lane;
            if (stride > 1 && ((jl >> 7) % stride) == 0) continue;  // a sampled block: its references are in the lists already (uniform)
            const float* col = blocks + (size_t)(jl >> 7) * (k + 1) * LB + (jl & (LB - 1));
            const float* qp = queries + (size_t)cq * k;
            float d = 0.0f;
            for (int tb = 0; tb < k; tb += 32) {
                const float qv = (tb + lane < k) ? __ldg(qp + tb + lane) : 0.0f;
                const int te = min(32, k - tb);
#pragma unroll 8
                for (int tt = 0; tt < te; ++tt) {
                    const float qt = __shfl_sync(0xffffffffu, qv, tt);
                    const float e = qt - __ldg(col + (size_t)(tb + tt) * LB);
                    d = EXACT ? __fadd_rn(d, __fmul_rn(e, e)) : __fmaf_rn(e, e, d);
                }
            }
            const float tau = __uint_as_float((unsigned)(keys[(size_t)cq * K + (K - 1)] >> 32));
            const bool pass = d <= tau && d < inf_f();  // padding lanes are NaN
            const unsigned pm = __ballot_sync(0xffffffffu, pass);
            if (pm == 0u) continue;
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(exact_count + cq, (unsigned)__popc(pm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + (unsigned)__popc(pm) > cap) {
                if (lane == 0) cb.status[1] = 1u;  // list full: the FP32 kernel launched behind the flag takes over
                continue;
            }
            if (pass) exact[(size_t)cq * cap + base + (unsigned)__popc(pm & ((1u << lane) - 1u))] = pack_key(d, index_base + jl);
        }
    }
}

cudaError_t tensor_topk_search(int k, int m, int n, int K, const float* d_queries, const float* d_blocks, const float* d_section,
                               int index_base, u64* d_keys, bool exact, int num_sms, cudaStream_t st, cudaMemPool_t pool,
                               int* launches, unsigned* d_stats)
{
    const ScreenSetup ss = screen_setup(k, d_section);
    const TensorGeom g = ss.g, gp = ss.gf;
    const bool modes = ss.modes;
    const unsigned* mode_word = ss.mode_word;
    const int nblocks = (n + LB - 1) / LB;
    const int rows = ss.rows;
    const int strips = (m + rows - 1) / rows;
    const float* hdr = d_section;
    const unsigned char* rimage = reinterpret_cast<const unsigned char*>(d_section + TENSOR_HDR_FLOATS);
    const int stride = TOPK_SAMPLE_STRIDE;
    int nl = 0;

    // 1. exact K nearest over the block sample -> tau per query
    const int splits_s = topk_choose_splits(m, (nblocks / stride + 1) * LB, num_sms);
    u64* sample_scratch = nullptr;
    cudaError_t e = pool ? cudaMallocFromPoolAsync((void**)&sample_scratch, topk_scratch_bytes(m, K, splits_s), pool, st)
                         : cudaMallocAsync((void**)&sample_scratch, topk_scratch_bytes(m, K, splits_s), st);
    if (e != cudaSuccess) return e;
    int l2 = 0;
    e = topk_search_launch(k, m, n, K, d_queries, d_blocks, index_base, d_keys, sample_scratch, splits_s, exact, st, &l2, stride, 1, nullptr);
    nl += l2;
    cudaError_t fe = cudaFreeAsync(sample_scratch, st);
    if (e != cudaSuccess) return e;
    if (fe != cudaSuccess) return fe;

    // 2. the screen with the fixed threshold, in query batches (as tensor_search; at most 1 GiB of exact lists per batch)
    const size_t per_strip_exact = (size_t)rows * (size_t)(stride * K + (int)(6.0 * stride * sqrt((double)K)) + 64) * sizeof(u64);
    const int max_batch_ctas = (int)std::max<size_t>((size_t)num_sms, std::min<size_t>((size_t)4 * num_sms, ((size_t)1 << 30) / per_strip_exact));
    int batch_strips = strips, nbatches = 1;
    if (strips > max_batch_ctas) {
        nbatches = (strips + max_batch_ctas - 1) / max_batch_ctas;
        batch_strips = (strips + nbatches - 1) / nbatches;
        batch_strips = std::min(max_batch_ctas, (batch_strips + num_sms - 1) / num_sms * num_sms);
        nbatches = (strips + batch_strips - 1) / batch_strips;
    }
    int splits = 1;
    {
        double best = 1e300;
        for (int sp = 1; sp <= std::min(nblocks, 64); ++sp) {
            const int t = (nblocks + sp - 1) / sp;
            const int se = (nblocks + t - 1) / t;
            const double waves = (double)(((long long)batch_strips * se + num_sms - 1) / num_sms);
            const double cost = waves * ((double)t + 24.0);
            if (cost < best * 0.97) { best = cost; splits = se; }
        }
    }
    const int tps = (nblocks + splits - 1) / splits;
    splits = (nblocks + tps - 1) / tps;
    CandBuf cb{};
    cb.fixed_threshold = 1u;
    cb.n_ctas = (unsigned)batch_strips * (unsigned)splits;
    const size_t batch_queries = (size_t)batch_strips * rows;
    // exact candidates per query: the references below the K-th SAMPLED one number K (stride - 1) on average with a
    // standard deviation of ~stride sqrt(K) (negative binomial); mean + 6 sigma
    const unsigned cap = (unsigned)(stride * K + (int)(6.0 * stride * sqrt((double)K)) + 64);
    const size_t exact_bytes = batch_queries * cap * sizeof(u64);
    {
        const size_t budget = tensor_record_budget() / 2;  // the other half of the scratch budget holds the exact lists
        const size_t common = std::min<size_t>(batch_queries * 64 + 65536, budget / 4);
        size_t region = ((budget - common) / cb.n_ctas) & ~(size_t)31;
        region = std::max<size_t>((size_t)rows * 64, std::min<size_t>((size_t)rows * 1024, region));
        cb.region_cap = (unsigned)region;
        cb.common_cap = (unsigned)common;
    }
    const size_t cand_records = (size_t)cb.n_ctas * cb.region_cap + cb.common_cap;
    const size_t qimg_bytes = (size_t)batch_strips * image_bytes(rows, g.KB, g.KS);
    const size_t off_band = (qimg_bytes + 255) & ~(size_t)255;
    const size_t off_amin = off_band + ((batch_queries * 4 + 255) & ~(size_t)255);
    const size_t off_qs = off_amin + ((batch_queries * 4 + 255) & ~(size_t)255);
    const size_t off_cnt = off_qs + ((batch_queries * 4 + 255) & ~(size_t)255);
    const size_t off_ccnt = off_cnt + (((size_t)nbatches * 4 + 255) & ~(size_t)255);
    const size_t off_ecnt = off_ccnt + (((size_t)cb.n_ctas * 4 + 255) & ~(size_t)255);
    const size_t off_exact = off_ecnt + ((batch_queries * 4 + 255) & ~(size_t)255);
    const size_t off_cand = off_exact + ((exact_bytes + 255) & ~(size_t)255);
    const size_t total = off_cand + cand_records * sizeof(TensorCand);
    unsigned char* scratch = nullptr;
    e = pool ? cudaMallocFromPoolAsync((void**)&scratch, total, pool, st) : cudaMallocAsync((void**)&scratch, total, st);
    if (e != cudaSuccess) return e;
    float* band = reinterpret_cast<float*>(scratch + off_band);
    unsigned* amin = reinterpret_cast<unsigned*>(scratch + off_amin);
    float* qscale = reinterpret_cast<float*>(scratch + off_qs);
    unsigned* common_counts = reinterpret_cast<unsigned*>(scratch + off_cnt);
    unsigned* exact_count = reinterpret_cast<unsigned*>(scratch + off_ecnt);
    u64* exact_list = reinterpret_cast<u64*>(scratch + off_exact);
    cb.cta_count = reinterpret_cast<unsigned*>(scratch + off_ccnt);
    cb.rec = reinterpret_cast<TensorCand*>(scratch + off_cand);
    cb.status = d_stats;
    tensor_status_init_kernel<<<(nbatches + 255) / 256, 256, 0, st>>>(d_stats, (unsigned)std::min<size_t>(cand_records, 0xffffffffu), common_counts,
                                                                      nbatches, mode_word, (unsigned)(g.KB * 64 + g.KS * 16),
                                                                      (unsigned)(gp.KB * 64 + gp.KS * 16));
    e = cudaGetLastError();
    nl += 1;
    for (int b = 0; b < nbatches && e == cudaSuccess; ++b) {
        const int s0 = b * batch_strips;
        const int bs = std::min(batch_strips, strips - s0);
        const long long q0 = (long long)s0 * rows;
        const int bm = (int)std::min<long long>((long long)bs * rows, (long long)m - q0);
        const float* bq = d_queries + (size_t)q0 * k;
        u64* bkeys = d_keys + (size_t)q0 * K;
        cb.common_count = common_counts + b;
        cb.n_ctas = (unsigned)bs * (unsigned)splits;
        e = cudaMemsetAsync(exact_count, 0, (size_t)bm * 4, st);
        if (e != cudaSuccess) break;
        e = screen_batch(ss, k, bq, bm, bs, splits, hdr, scratch, band, amin, qscale, rimage, nblocks, tps, cb, st, true, false);
        if (e != cudaSuccess) break;
        tensor_topk_threshold_kernel<<<(bm + 255) / 256, 256, 0, st>>>(bq, bm, k, hdr, bkeys, K, amin);
        e = cudaGetLastError();
        if (e != cudaSuccess) break;
        e = screen_batch(ss, k, bq, bm, bs, splits, hdr, scratch, band, amin, qscale, rimage, nblocks, tps, cb, st, false, true);
        if (e != cudaSuccess) break;
        const int rgrid = num_sms * 8;
        if (exact)
            tensor_rescore_topk_kernel<true><<<rgrid, 256, 0, st>>>(bq, k, d_blocks, index_base, cb, bkeys, K, exact_list, exact_count, cap, stride);
        else
            tensor_rescore_topk_kernel<false><<<rgrid, 256, 0, st>>>(bq, k, d_blocks, index_base, cb, bkeys, K, exact_list, exact_count, cap, stride);
        e = cudaGetLastError();
        if (e != cudaSuccess) break;
        // the lists of this batch into its keys -- unless something overflowed (then nothing of the screen is used)
        e = topk_merge_var_launch(bkeys, exact_list, exact_count, bm, K, cap, d_stats + 1, st);
        nl += modes ? 7 : 5;
    }
    fe = cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) return e;
    if (fe != cudaSuccess) return fe;
    // 3. fallback behind the overflow flag: exact FP32 K nearest over the NON-sampled blocks (the sample is in the keys;
    //    batches merged before the overflow only hold true K-nearest candidates, which the merge keeps or replaces)
    const int splits_f = topk_choose_splits(m, n, num_sms);
    u64* fb_scratch = nullptr;
    e = pool ? cudaMallocFromPoolAsync((void**)&fb_scratch, topk_scratch_bytes(m, K, splits_f), pool, st)
             : cudaMallocAsync((void**)&fb_scratch, topk_scratch_bytes(m, K, splits_f), st);
    if (e != cudaSuccess) return e;
    e = topk_search_launch(k, m, n, K, d_queries, d_blocks, index_base, d_keys, fb_scratch, splits_f, exact, st, &l2, stride, 0,
                           reinterpret_cast<const int*>(d_stats + 1));
    nl += l2;
    fe = cudaFreeAsync(fb_scratch, st);
    if (launches) *launches = nl;
    return e != cudaSuccess ? e : fe;
}

}  // namespace nns

// Host-side evaluation of the screen's error bound: a pure function of (k, mode, |q'|, max |r'|, sampled max |r'|), exported
// so that the CPU emulation of the screen (tests/test_tensor_bound.py) is checked against the SHIPPED formula and
// geometry, not against a copy of them.  out4 = { E(q), reference scale s, query scale t (0 = the query cannot be
// screened), contraction columns }.  mode 0 = the default BF16 layout of k, 2 = F16.
extern "C" int nns_b200_tensor_bound(int k, int mode, float a, float rmax, float rmax_sampled, float* out4)
{
    using namespace nns;
    if (k < 1 || k > TENSOR_MAX_K || out4 == nullptr || (mode != 0 && mode != 2)) return 1;  // NNS_B200_ERR_INVALID
    if (mode == 2 && !tensor_has_modes(k)) return 1;
    const TensorGeom g = tensor_geom(k, mode == 2);
    const int KP = g.KB * 64 + g.KS * 16;
    float s = 1.0f, t = 1.0f, E;
    if (mode == 2) {
        s = tensor_f16_ref_scale(rmax_sampled);
        t = tensor_f16_query_scale(s * a, s * rmax);
        E = tensor_error_bound_f16(KP, k, a, rmax, s, t > 0.0f ? t : 1.0f, g.en != 0);
    } else {
        E = tensor_error_bound(g.split != 0, KP, a, rmax);
    }
    out4[0] = E; out4[1] = s; out4[2] = t; out4[3] = (float)KP;
    return 0;
}
