// nns_plan.h -- compile-time shape tables shared by the kernels and the host-side planner.
#pragma once
#ifdef __CUDACC__
#define NNS_HD __host__ __device__
#else
#define NNS_HD
#endif

namespace nns {

constexpr int LOWK_MAX_K = 32;       // register-blocked FP32 path covers k = 1..32
constexpr int LOWK_MAX_STAGES = 8;
constexpr int LOWK_BAR_BYTES = 128;  // mbarrier area at the start of dynamic shared memory
constexpr int WIDE_QT = 4;           // queries per CTA of the reference-parallel kernel
constexpr int WIDE_THREADS = 256;

// ---- device index layout (tiled SoA) ----
// float[INDEX_HEADER_FLOATS] header, then float[nblocks][k + 1][128]: rows 0..k-1 are the
// coordinates of 128 consecutive reference points, row k is |r|^2 (FP32); tail lanes are NaN.
// header[0] = bit pattern of max_j |r_j|^2 (written with atomicMax, so >= every row-k value).
constexpr int INDEX_HEADER_FLOATS = 32;
NNS_HD constexpr int index_block_floats(int k) { return (k + 1) * 128; }

// reference blocks (of 128 points) per shared-memory tile: tiles are <= ~17 KiB
NNS_HD constexpr int lowk_tb(int k) { return k <= 4 ? 8 : k <= 8 ? 4 : k <= 16 ? 2 : 1; }
NNS_HD constexpr int lowk_tile_bytes(int k) { return lowk_tb(k) * index_block_floats(k) * 4; }
// Queries held in registers per thread.  Two blockings are compiled per k; the planner prefers
// lowk_q_default (measured best on B200 for the screened kernel, tools/lowk_tune.cu ->
// profiles/r1_tune_*.txt), lowk_q_alt is selectable through the flags word.
NNS_HD constexpr int lowk_q_default(int k) { return k <= 8 ? 4 : 2; }
NNS_HD constexpr int lowk_q_alt(int k) { return k <= 4 ? 8 : k <= 8 ? 2 : k <= 16 ? 4 : 1; }
// code-generation choices: CTAs/SM the register budget is held to, quads per loop body, quads
// (4 references) screened per threshold check in the filter kernel
NNS_HD constexpr int lowk_minb(int k, int q) { return ((k <= 4 && q <= 4) || (k >= 9 && k * q <= 32)) ? 2 : 1; }
NNS_HD constexpr int lowk_unroll(int k) { return k <= 16 ? 2 : 1; }
NNS_HD constexpr int lowk_screen_quads(int k) { return k <= 8 ? 2 : 1; }

}  // namespace nns
