/*
 * nns_b200.h -- C ABI of the B200-native brute-force nearest-neighbour engine.
 *
 * Drop-in boundary for the brute-force path of sty-hhh/NNS-CUDA.  Every entry point cites the
 * reference interface it replaces (file:line into the reference tree).  Plain pointers and
 * sizes only; `void *stream` is a cudaStream_t (NULL = the legacy default stream).  There is
 * no CPU fallback behind any of these calls: without a usable sm_100 GPU they fail with
 * NNS_B200_ERR_CUDA (the drop-in symbol prints the error and exits, like the reference).
 *
 * Semantics shared by all search calls (the reference's V0, core.cu:31-52):
 *   for each query i the result is the 0-based index j of the reference point with the
 *   smallest squared L2 distance sum_t (s[i*k+t] - r[j*k+t])^2 accumulated in FP32 over
 *   ascending t; exact ties resolve to the LOWEST index; a distance that is NaN never wins;
 *   if no distance is < +INF (including n == 0) the result is 0.
 */
#ifndef NNS_B200_H
#define NNS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NNS_B200_VERSION 200

/* status codes of the int-returning entry points */
#define NNS_B200_OK 0
#define NNS_B200_ERR_INVALID 1     /* bad argument (negative size, k <= 0, NULL pointer ...) */
#define NNS_B200_ERR_CUDA 2        /* a CUDA runtime call failed; see nns_b200_last_error()  */
#define NNS_B200_ERR_UNSUPPORTED 3 /* shape outside what the kernels cover                    */
#define NNS_B200_ERR_NOMEM 4       /* host or device allocation failed                        */

/* Reference blocks: the device-side reference set ("index") is a tiled structure-of-arrays:
 * a 32-float header (header[0] = max |r|^2), then float[num_blocks][k + 1][NNS_B200_REF_BLOCK]
 * -- k coordinate rows and one |r|^2 row per block of 128 points -- padded with NaN to a whole
 * number of blocks.  It replaces the reference's SoA transpose `rr_d` (core.cu:293-306, 362,
 * 367-370). */
#define NNS_B200_REF_BLOCK 128

/* Packed (distance, index) key: (float_bits(dist) << 32) | (uint32)index.  Distances are
 * >= +0 so the unsigned (and the signed 64-bit) order of keys is (dist, index) lexicographic:
 * a plain integer MIN implements "smaller distance, then lower index" exactly.  Replaces the
 * reference's per-block partial index lists + host re-reduction (core.cu:669-696, 821-852). */
#define NNS_B200_KEY_INIT 0x7F80000000000000ull /* (+INF, index 0) */

/* flags for nns_b200_search_keys / nns_b200_search_device */
#define NNS_B200_FLAG_V0_ROUNDING 1u /* round mul and add separately (bit-exact V0 distances;
                                        1.5x the FP32 work) instead of contracting to FMA   */
#define NNS_B200_FLAG_FORCE_LOWK 2u  /* testing: force the register-blocked low-k kernel    */
#define NNS_B200_FLAG_FORCE_WIDE 4u  /* testing: force the reference-parallel generic kernel */
#define NNS_B200_FLAG_FORCE_TENSOR 8u /* testing: force the tcgen05 screen + exact re-score (k <= 128) */
#define NNS_B200_FLAG_EXACT_FORM 16u /* evaluate every pair in V0's subtract-square form (2k FP32
                                        lane-slots per pair) instead of screening pairs with the
                                        norm expansion and evaluating only the survivors that way;
                                        both return the same indices                          */
#define NNS_B200_FLAG_TEST_TINY_CANDIDATES (1u << 28) /* testing: 32-record candidate regions in the
                                        tcgen05 path, forcing its overflow fallback              */
/* bits 8-15 / 16-23 / 24-27: tuning overrides (queries per thread, warps per CTA, ring stages) */

/* ---- the drop-in symbol ------------------------------------------------------------------
 * Replaces vN::cudaCall (core.cu:23-29; same signature in all 14 namespaces) as consumed by
 * the function pointer `func` (main.cu:7) and called at main.cu:74.  Host pointers in,
 * row-major s_points[m][k] / r_points[n][k] (core.cu:41, main.cu:27-34), never written.
 * *results is malloc'd here (core.cu:31) and owned by the caller (free()).  No return code:
 * on a CUDA failure it prints "Error: file:line, code:N, reason: ..." and exit(1)s exactly
 * like the reference's CHECK macro (utils.h:16-26).  Synchronous; thread-safe; restores the
 * caller's current device. */
void nns_b200_cudaCall(int k, int m, int n, float *s_points, float *r_points, int **results);

/* Same work as nns_b200_cudaCall with a status code and a caller-provided int[m] result. */
int nns_b200_search_host(int k, int m, int n, const float *s_points, const float *r_points,
                         int *results);

/* nns_b200_search_host that also returns, per query, the FP32 squared distance to the reported
 * neighbour (+INF when no reference has a distance below +INF, the case in which V0 reports index
 * 0).  An extension: the reference's callback returns indices only (core.cu:52). */
int nns_b200_search_host_dist(int k, int m, int n, const float *s_points, const float *r_points,
                              int *results, float *distances);

/* Single-process multi-GPU search over host arrays: replaces v8/v9::cudaCall's OpenMP
 * fan-out + host merge (core.cu:761-853, 965-1057).  shard_mode 0 = query-sharded (each GPU
 * gets a query slice and all references), 1 = reference-sharded (contiguous reference
 * slices, core.cu:781-791; partial results merged by an integer MIN over packed keys, so
 * results are identical for every num_gpus).  num_gpus <= 0 means all visible devices. */
int nns_b200_search_multi(int k, int m, int n, const float *s_points, const float *r_points,
                          int *results, int num_gpus, int shard_mode);

/* ---- device-resident index handle: build once, query many -----------------------------------
 * The reference uploads and transposes the reference set inside every call (core.cu:351-370 and its
 * copies in v4..v9); a handle keeps the tiled-SoA index (and, once a search is planned onto the
 * tcgen05 path, its BF16 operand images) resident in the HBM of `device` (< 0: the current device).
 * r_points / s_points are host arrays exactly as in the drop-in symbol (pageable or pinned);
 * distances may be NULL.  Searches on one handle are serialised; handles are independent. */
typedef struct nns_b200_index nns_b200_index_t;
int nns_b200_index_create(int k, int n, const float *r_points, int device, nns_b200_index_t **out);
int nns_b200_index_search(nns_b200_index_t *index, int m, const float *s_points, int *results,
                          float *distances);
int nns_b200_index_size(const nns_b200_index_t *index, int *k, int *n);
int nns_b200_index_destroy(nns_b200_index_t *index);

/* ---- exact search through a bucketed KD-tree (k <= 32) -----------------------------------------
 * What the reference's v10 / v11 were meant to be (core.cu:1059-1163: CPU KD-tree; core.cu:1289-1451: GPU
 * KD-tree whose kernel body is commented out and which therefore returns zeros): the tree is built on the
 * host by median splits in an implicit heap layout like the reference's (core.cu:1072-1114), its leaves
 * are 128-point blocks in the engine's tiled-SoA layout, and the search runs on the GPU, one warp per
 * query, with exactly the brute-force kernels' FP32 distance arithmetic -- the answers are V0's
 * (lowest index on exact ties, index 0 when no distance is < +INF), not approximations.  For low k and many
 * references it visits a few leaves per query instead of all n points.  distances may be NULL. */
typedef struct nns_b200_tree nns_b200_tree_t;
int nns_b200_tree_create(int k, int n, const float *r_points, int device, nns_b200_tree_t **out);
int nns_b200_tree_search(nns_b200_tree_t *tree, int m, const float *s_points, int *results,
                         float *distances);
int nns_b200_tree_destroy(nns_b200_tree_t *tree);

/* ---- lifetime ----------------------------------------------------------------------------
 * Replaces the load-time WarmUP static initialiser (core.cu:1900-1933): nothing touches the
 * GPU at load; nns_b200_init(device) creates the per-device state eagerly (device < 0 = the
 * current device), otherwise it is created lazily by the first call.  nns_b200_shutdown()
 * releases every cached device/pinned buffer. */
int nns_b200_init(int device);
int nns_b200_shutdown(void);
int nns_b200_version(void);
int nns_b200_device_sms(int device); /* SM count of the device (< 0: current); -1 on failure */
const char *nns_b200_last_error(void); /* thread-local text of the last failure */

/* ---- device-resident building blocks (inputs already in HBM) ------------------------------ */

/* number of floats of the tiled-SoA index for n reference points of k dims */
size_t nns_b200_index_floats(int k, int n);

/* AoS float[n][k] (device) -> tiled SoA index (device, nns_b200_index_floats(k,n) floats,
 * 16-byte aligned).  Replaces v4::mat_inv_kernel (core.cu:293-306). */
int nns_b200_index_build(int k, int n, const float *d_refs_aos, float *d_index, void *stream);

/* An index that several GPUs build together (one process per GPU, e.g. under torch.distributed):
 * each builds the slice [j0, j0 + cn) of the n_total references -- `part_blocks` blocks starting at block
 * j0 / 128, of which those past cn points are padding -- straight into its copy of the index, with the
 * centre of the tcgen05 operand images fixed by the caller (nns_b200_sample_centre on the host array,
 * so that every slice uses the same one); the slices' byte ranges (nns_b200_index_part_ranges: out6 =
 * { blocks offset, blocks bytes, image offset, image bytes, index-header word offset, section-header
 * word offset }, bytes from d_index; the two header words are 4 bytes each, and the section's flag word
 * lies 64 bytes after its partial-max word) are then exchanged (all-gather over NVLink), and
 * nns_b200_index_finish folds the per-part maxima.  Replaces v8/v9's "every GPU uploads and transposes
 * everything it needs" (core.cu:778-818, 982-1022). */
int nns_b200_sample_centre(int k, int n, const float *r_points, float *centre_out);
int nns_b200_index_build_part(int k, int n_total, int j0, int cn, int part_blocks,
                              const float *d_refs_aos_part, float *d_index, const float *centre,
                              int part, void *stream);
int nns_b200_index_part_ranges(int k, int n_total, int j0, int part_blocks, int part, size_t *out6);
int nns_b200_index_finish(int k, int n_total, float *d_index, int parts, void *stream);

/* keys[i] = NNS_B200_KEY_INIT for i < m */
int nns_b200_keys_init(uint64_t *d_keys, int m, void *stream);

/* The hot path.  For each query i: keys[i] = min(keys[i], key(best distance, index_base + j))
 * over the n references held in d_index.  d_queries is AoS float[m][k] on the device.
 * Calling it for several reference shards (each with its own index_base) accumulates the
 * global minimum.  Replaces v3/v4/v7/v8/v9::cudaCallKernel (core.cu:215-257, 307-349, 589-633,
 * 716-760, 872-964) and the second-stage reduce (core.cu:669-696). */
int nns_b200_search_keys(int k, int m, int n, const float *d_queries, const float *d_index,
                         int index_base, uint64_t *d_keys, unsigned flags, void *stream);

/* d_idx[i] = low 32 bits of keys[i]; d_dist[i] = the FP32 squared distance (may be NULL). */
int nns_b200_keys_unpack(const uint64_t *d_keys, int m, int *d_idx, float *d_dist, void *stream);

/* ---- K nearest neighbours (extension: the reference returns one index per query, core.cu:52) ------
 * For each query the K <= 32 references with the smallest FP32 squared distance (V0's form,
 * core.cu:38-43), ordered by (distance, index): ties go to the lower index, so K = 1 is the 1-NN
 * answer.  References at distance NaN / +INF are never reported; missing neighbours (n < K) are
 * index -1, distance +INF.
 * nns_b200_topk_keys: d_keys is uint64[m][K], ascending packed keys, initialised with
 * nns_b200_keys_init(d_keys, m * K); calls for several reference shards accumulate like
 * nns_b200_search_keys does.  nns_b200_topk_unpack writes int[m][K] (+ float[m][K]).
 * nns_b200_search_topk_host: host arrays in, host int[m][K] (+ float[m][K], may be NULL) out. */
int nns_b200_topk_keys(int k, int m, int n, int K, const float *d_queries, const float *d_index,
                       int index_base, uint64_t *d_keys, unsigned flags, void *stream);
int nns_b200_topk_unpack(const uint64_t *d_keys, int m, int K, int *d_idx, float *d_dist, void *stream);
int nns_b200_search_topk_host(int k, int m, int n, int K, const float *s_points,
                              const float *r_points, int *indices, float *distances);

/* Convenience: index_build + keys_init + search_keys + keys_unpack on device arrays.
 * d_workspace must hold nns_b200_workspace_bytes(k, m, n) bytes (256-byte aligned). */
size_t nns_b200_workspace_bytes(int k, int m, int n);
int nns_b200_search_device(int k, int m, int n, const float *d_queries, const float *d_refs_aos,
                           int *d_idx, void *d_workspace, size_t workspace_bytes, unsigned flags,
                           void *stream);

/* ---- host-side planning (pure function, no GPU needed; exported for tests) ----------------
 * Fills plan[0..7] = { path (0 = low-k, 1 = wide, 2 = tensor), queries per thread,
 * consumer warps per CTA, pipeline stages, query blocks, reference splits, reference blocks
 * per split, dynamic shared memory bytes } for a device with num_sms SMs.  Path 2 (tcgen05) reports
 * 256-query strips as its query blocks; its reference splits are chosen at launch. */
int nns_b200_plan(int k, int m, int n, unsigned flags, int num_sms, int *plan);

/* diagnostics of the last tensor-path search on the current device (synchronises the device):
 * out4 = { (query, 32-reference unit) candidates emitted by the tcgen05 screen, 1 if the candidate
 * buffer overflowed and the FP32 kernel launched behind it redid the search, candidate capacity of a
 * query batch, contraction length (16-bit columns) of the operand images the index holds | precision mode << 16
 * (0 = BF16 operands with FP32 accumulators, 2 = F16 operands with F16 accumulators) } */
int nns_b200_tensor_stats(unsigned *out4);

/* host-side evaluation of the tcgen05 screen's error bound E(q) (pure function, no GPU needed; exported so that the
 * tests' CPU emulation of the screen is checked against the shipped formula): a = |q - c|, rmax = max |r - c|,
 * rmax_sampled = the sampled radius the F16 scale is derived from; mode 0 = the default BF16 operand layout of k,
 * 2 = F16 operands and accumulators (10 <= k <= 128).  out4 = { E, reference scale s, query scale t (0 = the query
 * cannot be screened), contraction columns }. */
int nns_b200_tensor_bound(int k, int mode, float a, float rmax, float rmax_sampled, float *out4);

/* number of kernels of this library launched by this process so far (bench.py's gpu_launches) */
unsigned long long nns_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NNS_B200_H */
