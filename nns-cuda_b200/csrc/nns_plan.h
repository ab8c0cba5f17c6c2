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

// reference blocks (of 128 points) per shared-memory tile: tiles are <= 16 KiB
NNS_HD constexpr int lowk_tb(int k) { return k <= 4 ? 8 : k <= 8 ? 4 : k <= 16 ? 2 : 1; }
NNS_HD constexpr int lowk_tile_bytes(int k) { return lowk_tb(k) * k * 128 * 4; }
// queries held in registers per thread: default and the alternative that is also compiled
NNS_HD constexpr int lowk_q_default(int k) { return k <= 4 ? 8 : k <= 16 ? 4 : 2; }
NNS_HD constexpr int lowk_q_alt(int k) { return k <= 4 ? 4 : k <= 16 ? 2 : 1; }

// code-generation choices per (k, q): CTAs/SM the register budget is held to, quads per loop
// body, software-pipelined argmin (tools/lowk_tune.cu measures the alternatives on a B200)
NNS_HD constexpr int lowk_minb(int k, int q) { return (k * q <= 16) ? 2 : 1; }
NNS_HD constexpr int lowk_unroll(int k) { return k <= 8 ? 2 : 1; }
NNS_HD constexpr bool lowk_pipe(int k) { return true; }

}  // namespace nns
