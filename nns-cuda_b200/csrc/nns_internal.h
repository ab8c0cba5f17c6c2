// nns_internal.h -- launch entry points shared between the kernel translation units and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include "nns_common.cuh"
#include "lowk_search.cuh"

namespace nns {

// index_build.cu
cudaError_t launch_index_build(int k, int n, const float* d_refs_aos, float* d_index, cudaStream_t st);
cudaError_t launch_keys_init(u64* d_keys, int m, cudaStream_t st);
cudaError_t launch_keys_unpack(const u64* d_keys, int m, int* d_idx, float* d_dist, cudaStream_t st);

// wide_search.cu
struct WideArgs {
    const float* queries;  // device AoS [m][k]
    int m, k;
    const float* index;    // tiled SoA
    int nblocks;
    int blocks_per_split;
    int index_base;
    u64* keys;
    int nqg;     // query groups of WIDE_QT (grid.x)
    int splits;  // grid.y
    cudaStream_t stream;
};
cudaError_t wide_launch(bool exact, const WideArgs& a);

// lowk_inst_N.cu
cudaError_t lowk_launch_range_0(int k, int q, bool exact, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_1(int k, int q, bool exact, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_2(int k, int q, bool exact, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_3(int k, int q, bool exact, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_4(int k, int q, bool exact, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_5(int k, int q, bool exact, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_6(int k, int q, bool exact, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_7(int k, int q, bool exact, const LowkArgs& a, int* occ);

}  // namespace nns
