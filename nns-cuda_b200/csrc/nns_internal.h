// nns_internal.h -- launch entry points shared between the kernel translation units and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include "nns_common.cuh"
#include "lowk_search.cuh"

namespace nns {

// ---- headers ----
// index header (INDEX_HEADER_FLOATS words): [0] max |r|^2 (bits); [HDR_PART_MAX + g] the same of GPU g's slice
// tensor section header (TENSOR_HDR_FLOATS words): [0..511] centre, [THDR_MAX] max |r'|^2 (bits),
// [THDR_FLAGS] bit 0: unusable; [THDR_PART_MAX + g], [THDR_PART_FLAGS + g] per-GPU partials (multi-GPU ingest)
constexpr int MAX_PEERS = 8;
constexpr int HDR_PART_MAX = 8;
constexpr int THDR_MAX = 512, THDR_FLAGS = 513, THDR_MODE = 515, THDR_SCALE = 517, THDR_PART_MAX = 520, THDR_PART_FLAGS = 536;
// [THDR_MODE] 0 = the default BF16 operand images (split precision up to TENSOR_SPLIT_MAX_K, plain above), 2 = plain F16 operands
// with F16 accumulators (TENSOR_PLAIN_MIN_K <= k <= 128, chosen per index by tensor_index_build); [THDR_SCALE] its reference scale
constexpr unsigned TMODE_DEFAULT = 0u, TMODE_F16 = 2u;
struct BlockDsts { float* p[MAX_PEERS]; int count; };            // first block of the part in every destination index
struct ImageDsts { unsigned char* p[MAX_PEERS]; int count; };    // first tile image of the part in every destination
struct HeaderPeers { float* header[MAX_PEERS]; float* section[MAX_PEERS]; int count; int self; };
struct TensorCentre { float c[512]; };                           // a caller-fixed centre, passed by value

// index_build.cu
cudaError_t launch_index_build(int k, int n, const float* d_refs_aos, float* d_header, float* d_blocks,
                               bool reset_header, cudaStream_t st);
// the same into several destinations (peer GPUs), max |r|^2 accumulated in header word `hmax_word`
cudaError_t launch_index_build_to(int k, int n, const float* d_refs_aos, float* d_header, int hmax_word,
                                  const BlockDsts& dst, bool reset_header, cudaStream_t st, int write_blocks = 0);
cudaError_t launch_header_publish(const HeaderPeers& hp, int g, cudaStream_t st);
cudaError_t launch_header_fold(float* d_header, float* d_section, int parts, cudaStream_t st);
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
constexpr int TENSOR_MAX_K = 509;    // 8 blocks of 64 columns incl. the 3 norm columns (tensor_longk.cu above 128)
#ifndef NNS_T_SPLIT_MAX
#define NNS_T_SPLIT_MAX 42
#endif
constexpr int TENSOR_SPLIT_MAX_K = NNS_T_SPLIT_MAX;  // 3k <= 128 contraction columns
constexpr int TENSOR_PLAIN_MIN_K = 10;  // below, the plain band is never selective enough (k = 3: 2E = 6e-3 vs d^2 ~ 4e-5)
constexpr int TENSOR_HDR_FLOATS = 1024;  // [0..127] centre, [128] max |r'|^2 bits, [129] flags
int tensor_kp(int k);
size_t tensor_section_floats(int k, int n);
size_t tensor_image_bytes_per_block(int k);
// section header: zeroed, centre = `fixed` or the mean of a strided sample of the n references in d_blocks
cudaError_t tensor_section_init(int k, int n, const float* d_blocks, float* d_section, const TensorCentre* fixed,
                                cudaStream_t st);
// BF16 operand images of the `cn` references in d_blocks_part (whole blocks) into every destination;
// max |r'|^2 / flags accumulate in words max_word / flag_word of the section header d_hdr
cudaError_t tensor_image_build(int k, int cn, const float* d_blocks_part, float* d_hdr, int max_word, int flag_word,
                               const ImageDsts& dst, cudaStream_t st, int write_blocks = 0);
// section_init (sampled centre) + image_build of the whole index
// d_header: the FP32 index header (max |r|^2 feeds the precision-mode probe); NULL = no probe, split layout
cudaError_t tensor_index_build(int k, int n, const float* d_header, const float* d_blocks, float* d_section, cudaStream_t st);
bool tensor_has_modes(int k);
// scratch comes from `pool` (stream-ordered); *launches = kernels launched
cudaError_t tensor_search(int k, int m, int n, const float* d_queries, const float* d_blocks, const float* d_section,
                          int index_base, u64* d_keys, bool exact, int num_sms, cudaStream_t st, cudaMemPool_t pool,
                          int* launches, unsigned* d_stats, bool tiny_candidate_buffer);

// topk_search.cu -- K nearest neighbours, K <= TOPK_MAX_K (extension; FP32 V0-form distances, any k <= TOPK_MAX_DIMS)
constexpr int TOPK_MAX_K = 32;
constexpr int TOPK_MAX_DIMS = 1024;
int topk_choose_splits(int m, int n, int num_sms);
size_t topk_scratch_bytes(int m, int K, int splits);
// stride > 1 splits the reference blocks into a sample (b % stride == 0; only_sampled = 1) and the rest (0);
// enable != NULL: a conditional launch that exits at once when *enable == 0
cudaError_t topk_search_launch(int k, int m, int n, int K, const float* d_queries, const float* d_blocks, int index_base,
                               u64* d_keys, u64* d_scratch, int splits, bool exact, cudaStream_t st, int* launches,
                               int stride = 1, int only_sampled = 0, const int* enable = nullptr);
cudaError_t topk_merge_var_launch(u64* d_keys, const u64* d_exact, const unsigned* d_count, int m, int K, unsigned cap,
                                  const unsigned* skip_if_set, cudaStream_t st);
// K nearest neighbours through the tcgen05 screen (tensor_search.cu): exact FP32 top-K over a block sample gives
// every query a distance threshold, the screen keeps the 32-reference units that can hold anything below it,
// their exact distances are merged into the sorted key lists; *launches = kernels launched
cudaError_t tensor_topk_search(int k, int m, int n, int K, const float* d_queries, const float* d_blocks, const float* d_section,
                               int index_base, u64* d_keys, bool exact, int num_sms, cudaStream_t st, cudaMemPool_t pool,
                               int* launches, unsigned* d_stats);
cudaError_t topk_unpack_launch(const u64* d_keys, int m, int K, int* d_idx, float* d_dist, cudaStream_t st);

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
