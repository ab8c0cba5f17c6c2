// nns_internal.h -- launch entry points shared between the kernel translation units and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include "nns_common.cuh"
#include "lowk_search.cuh"

namespace nns {

// index_build.cu
cudaError_t launch_index_build(int k, int n, const float* d_refs_aos, float* d_header, float* d_blocks,
                               bool reset_header, cudaStream_t st);
cudaError_t launch_keys_init(u64* d_keys, int m, cudaStream_t st);
cudaError_t launch_keys_merge(u64* d_dst, const u64* d_src, int m, cudaStream_t st);
cudaError_t launch_keys_unpack(const u64* d_keys, int m, int* d_idx, float* d_dist, cudaStream_t st);

// wide_search.cu
struct WideArgs {
    const float* queries;  // device AoS [m][k]
    int m, k;
    const float* blocks;   // first reference block of the range: [nblocks][k+1][128]
    int nblocks;
    int blocks_per_split;
    int index_base;
    u64* keys;
    int nqg;     // query groups of WIDE_QT (grid.x)
    int splits;  // grid.y
    cudaStream_t stream;
    const int* enable = nullptr;  // optional device flag: the kernel exits immediately when it is 0
};
cudaError_t wide_launch(bool exact, const WideArgs& a);

// tensor_search.cu -- tcgen05 path for k <= TENSOR_MAX_K (split-precision BF16 up to TENSOR_SPLIT_MAX_K)
constexpr int TENSOR_MAX_K = 128;
constexpr int TENSOR_SPLIT_MAX_K = 42;  // 3k <= 128 contraction columns
constexpr int TENSOR_HDR_FLOATS = 256;  // [0..127] centre, [128] max |r'|^2 bits, [129] flags
int tensor_kp(int k);
size_t tensor_section_floats(int k, int n);
cudaError_t tensor_index_build(int k, int n, const float* d_refs_aos, float* d_section, cudaStream_t st);
cudaError_t tensor_search(int k, int m, int n, const float* d_queries, const float* d_blocks, const float* d_section,
                          int index_base, u64* d_keys, bool exact, int num_sms, cudaStream_t st, int* launches,
                          unsigned* d_stats, bool tiny_candidate_buffer);

// lowk_inst_N.cu (N = (k-1)/2)
cudaError_t lowk_launch_range_0(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_1(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_2(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_3(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_4(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_5(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_6(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_7(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_8(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_9(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_10(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_11(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_12(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_13(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_14(int k, int q, int mode, const LowkArgs& a, int* occ);
cudaError_t lowk_launch_range_15(int k, int q, int mode, const LowkArgs& a, int* occ);

}  // namespace nns
