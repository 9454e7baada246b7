/*
 * fedmlp_b200.h — C ABI of libfedmlp_b200.so
 *
 * Drop-in boundary for the FedMLP per-round hot path (tag + prototypes + loss + FedAvg) on
 * NVIDIA B200 (sm_100a).  The reference (szbonaldo/FedMLP) is pure Python/PyTorch and has no
 * FFI of its own; every entry point below therefore cites the reference
 * *Python* block it replaces.  Host code (the fedmlp_b200 package, ctypes) keeps the reference's call
 * surface and forwards raw device pointers + the caller's CUDA stream to these functions.
 *
 * Conventions (all entry points):
 *   - return 0 on success, a NEGATIVE fmlp_status for argument errors, or a POSITIVE
 *     cudaError_t value if a CUDA runtime call / launch failed.  No C++ exception crosses.
 *   - no allocation, no host synchronisation, no hidden streams: every kernel is launched on
 *     `stream` on the caller's current device; workspaces are caller-provided
 *     (query the size with the matching *_ws_bytes function).
 *   - pointers named *_dev / unqualified tensors are DEVICE pointers (borrowed);
 *     pointers documented "host" are small host arrays that are copied into the kernel's
 *     parameter block at launch time (so a launch is self-contained and graph-capturable).
 *   - matrices are row-major and dense unless a leading dimension is given.
 *   - "segment" = one simulated federated client whose rows are stored contiguously in a
 *     batched [N_total, D] matrix; `seg_rows` is the host prefix array [S+1] of row offsets.
 *     A single client is S = 1, seg_rows = {0, N}.
 *   - class sets are bit masks (bit c = class c); at most FMLP_MAX_CLASSES classes.
 */
#ifndef FEDMLP_B200_H_
#define FEDMLP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMLP_ABI_VERSION 4 /* 4: + fmlp_agg_tails_local_f32, fmlp_set/get_tuning, fmlp_host_copy_many; 3: fmlp_tag_sim_f32 takes a workspace (class-vector table); 2: allreduce n_chunks, pool_tag, eval, adam */
#define FMLP_MAX_CLASSES 32   /* class bit masks are uint32_t                        */
#define FMLP_MAX_SEGMENTS 64  /* segments (clients) per launch                        */
#define FMLP_MAX_CLIENTS 64   /* client buffers folded per fedavg launch              */

typedef enum fmlp_status {
    FMLP_OK = 0,
    FMLP_ERR_BAD_ARG = -1,     /* null pointer, negative size, K/S/C out of range       */
    FMLP_ERR_UNSUPPORTED = -2, /* shape/alignment the kernels do not handle             */
    FMLP_ERR_WORKSPACE = -3    /* workspace too small                                   */
} fmlp_status;

typedef void* fmlp_stream_t; /* cudaStream_t */

/* ------------------------------------------------------------------ misc */
int fmlp_abi_version(void);
/* Human-readable text for a return code of any function in this header. */
const char* fmlp_status_string(int code);
/* Number of SMs of the current device (grid sizing is derived from it); <0 on error. */
int fmlp_sm_count(void);
/* Kernel launches issued by this library since it was loaded (process-wide, monotonic). */
unsigned long long fmlp_launch_count(void);

/* Host utility (no CUDA call): n independent copies dsts[i] <- srcs[i] of nbytes[i] bytes, split over n_threads host
 * threads by bytes.  FedAvg's CPU-state_dict path (the reference hands over `net.cpu()` weights in stage 2,
 * utils/local_training.py:1251, main.py:196) packs a client's 727 pageable tensors into one pinned buffer with it. */
int fmlp_host_copy_many(const void* const* srcs, void* const* dsts, const int64_t* nbytes, int64_t n, int n_threads);

/* Scheduling knobs for rounds that run several of these kernels concurrently on different streams (process-wide;
 * value -1 = unset: the environment variable of the same meaning, then the built-in default, applies).
 *   FMLP_TUNE_PROTO_PAD_SMEM_KB   unused dynamic shared memory requested per prototype-accumulate CTA (0..56 KB):
 *                                 limits how many of them share an SM with a CTA of the similarity kernel.  Measured
 *                                 on B200 (profiles/r02_exp_coresidency.txt): 40 KB -> at most one next to a
 *                                 similarity CTA, round 0.131 -> 0.124 ms at BASELINE configs[1].   env FMLP_PROTO_PAD_SMEM_KB
 *   FMLP_TUNE_SIM_REQUEST_SMEM_KB the similarity kernel requests at least this much shared memory per CTA (0..227).
 *                                 env FMLP_SIM_REQUEST_SMEM_KB
 *   FMLP_TUNE_SIM_SMEM_BUDGET_KB  shared memory the similarity kernel's ring may use (64..227, default 200). env FMLP_SIM_SMEM_KB
 *   FMLP_TUNE_SELECT_CLUSTER      CTAs per (segment, class) item of the selection kernel: 1, 2, 4 or 8 (one thread-block
 *                                 cluster per item); 0 / unset = as few as keep a CTA's keys in registers. env FMLP_SELECT_CLUSTER */
#define FMLP_TUNE_PROTO_PAD_SMEM_KB 0
#define FMLP_TUNE_SIM_REQUEST_SMEM_KB 1
#define FMLP_TUNE_SIM_SMEM_BUDGET_KB 2
#define FMLP_TUNE_SELECT_CLUSTER 3
#define FMLP_TUNE_COUNT 4
int fmlp_set_tuning(int knob, int value);
int fmlp_get_tuning(int knob);

/* ------------------------------------------------------------------ K1: FedAvg
 * Replaces utils/FedAvg.py:7-14 `FedAvg` (and its twin `Fed_w` :16-23):
 *     out[j] = ((((w_0[j]*n_0) + w_1[j]*n_1) + ...) + w_{K-1}[j]*n_{K-1}) / sum(n)
 * evaluated per element in fp32 in exactly that order (separate round-to-nearest multiply and
 * add, IEEE divide), so the result is bit-identical to the reference run on CPU tensors.
 */
enum {
    FMLP_FEDAVG_DIVIDE = 1,     /* apply the final `/ divisor`                          */
    FMLP_FEDAVG_ACCUMULATE = 2  /* start from out[j] instead of w_0[j]*n_0 (chains      */
                                /* launches when K > FMLP_MAX_CLIENTS, still in order)  */
};

/* Flat form: every client is one contiguous fp32 buffer of P elements.
 *   srcs    host array [K] of device pointers          weights  host array [K]          */
int fmlp_fedavg_flat_f32(const float* const* srcs, const float* weights, int K, int64_t P,
                         float divisor, int flags, float* out, fmlp_stream_t stream);

/* int64 buffers (BatchNorm `num_batches_tracked`): reference multiplies/adds in int64 when the
 * weights are integers and the true division then yields float32 (FedAvg.py:13).
 * weights_integral != 0: acc(int64) = sum w_i[j]*(int64)n_i ; out = (float)acc / (float)divisor
 * weights_integral == 0: acc(float) = (float)w_0[j]*(float)n_0 + ... ; out = acc / (float)divisor */
int fmlp_fedavg_flat_i64(const int64_t* const* srcs, const double* weights, int K, int64_t J,
                         double divisor, int weights_integral, int flags, float* out,
                         fmlp_stream_t stream);

/* Multi-tensor form: averages T separately-allocated tensors per client in ONE launch, reading
 * the clients' state_dict storage in place (no packing copy).
 *   src_table_dev  device array [T*K]: src_table_dev[t*K + i] = tensor t of client i
 *   dst_table_dev  device array [T]  : output tensor t
 *   numel_dev      device array [T]
 *   chunk_tensor_dev / chunk_start_dev   device arrays [n_chunks]: chunk c covers elements
 *        [chunk_start, chunk_start + FMLP_FEDAVG_CHUNK) of tensor chunk_tensor (clipped to numel)
 *   weights        host array [K]                                                      */
#define FMLP_FEDAVG_CHUNK 2048
int fmlp_fedavg_multi_f32(const float* const* src_table_dev, float* const* dst_table_dev,
                          const int64_t* numel_dev, const int32_t* chunk_tensor_dev,
                          const int64_t* chunk_start_dev, int64_t n_chunks, int T,
                          const float* weights, int K, float divisor, int flags,
                          fmlp_stream_t stream);
/* Same for T int64 tensors -> float32 outputs (one thread per element; tensors are scalars in
 * practice).  elem_tensor_dev/elem_index_dev enumerate all J elements.                   */
int fmlp_fedavg_multi_i64(const int64_t* const* src_table_dev, float* const* dst_table_dev,
                          const int32_t* elem_tensor_dev, const int64_t* elem_index_dev,
                          int64_t J, int T, const double* weights, int K, double divisor,
                          int weights_integral, int flags, fmlp_stream_t stream);

/* Multi-GPU FedAvg in ONE kernel per rank: K-way weighted fold of this rank's client buffers fused
 * with a two-shot all-reduce over NVLink peer memory (reduce-scatter by peer stores that overlap
 * the HBM-bound fold, fixed-order slice reduction, all-gather by peer stores).  There is no
 * reference counterpart (the reference is single-process, main.py:130,196); semantics = the global
 * weighted mean of utils/FedAvg.py:7-14 over the clients of all ranks.
 *   srcs / weights   host arrays [K]: this rank's client buffers (P floats each, 16-B aligned) and
 *                    their weights PRE-NORMALISED by the global sum (n_k / sum over all ranks)
 *   stage_ptrs       host array [world]: every rank's inbox   (symmetric memory, world*slice_len floats)
 *   result_ptrs      host array [world]: every rank's result buffer (>= world*slice_len floats)
 *   flag_ptrs        host array [world]: every rank's flag words (FMLP_AR_FLAG_WORDS x uint32,
 *                    zero-initialised once, never touched by the host afterwards)
 *   slice_len        floats per rank over all chunks, multiple of 4*n_chunks, world*slice_len >= P;
 *                    P % 4 == 0
 *   n_chunks         1..FMLP_AR_MAX_CHUNKS pipeline chunks: the all-gather of chunk c-1 overlaps the
 *                    fold + reduce-scatter of chunk c+1
 *   epoch_dev        device uint32 of THIS rank, zero-initialised once; the kernel uses *epoch_dev+1
 *                    as the call's epoch and stores it back (so CUDA-graph replays stay in step)
 * Collective: every rank must launch it (same P, slice_len, n_chunks); the kernel returns when
 * result_ptrs[rank] is complete. */
#define FMLP_AR_MAX_CHUNKS 16
#define FMLP_AR_FLAG_WORDS 512
int fmlp_fedavg_allreduce_f32(const float* const* srcs, const float* weights, int K, int64_t P,
                              float* const* stage_ptrs, float* const* result_ptrs,
                              uint32_t* const* flag_ptrs, int64_t slice_len, int n_chunks, int rank,
                              int world, uint32_t* epoch_dev, fmlp_stream_t stream);

/* Round-2 form of the multi-GPU aggregation: a NON-cooperative work-queue kernel (starts on whatever
 * SM slots are free, shares the SMs with the tagging kernels of another stream) that reduces through
 * the NVSwitch (multimem.ld_reduce / multimem.st on a multicast mapping of the symmetric buffers) or,
 * without multicast, by peer loads in rank order + peer stores.  Replaces main.py:218-234 across ranks:
 * FedAvg of the flat parameters, plus the small tails riding the same exchange:
 *   fp32 vector  [ P parameters | T tail floats (sum_k w_k * tail_k, e.g. the 2C*D prototypes) ]
 *   fp64 tail    [ M scalars ] summed over ranks (class weight sums, n*t sums, int64 counter sums)
 *   srcs / tail_srcs / weights   host arrays [K]: client buffers (P floats), tail vectors (T floats, or NULL
 *                    when T == 0) and weights PRE-NORMALISED by the global sum
 *   tail_f64         device [M] doubles: this rank's partial sums (fmlp_agg_tail_pack_f64), NULL iff M == 0
 *   partial_ptrs / result_ptrs   host arrays [world]: every rank's symmetric buffers of
 *                    fmlp_fedavg_allreduce_q_buffer_floats(P, T, M) floats
 *   flag_ptrs        host array [world]: FMLP_AR_FLAG_WORDS x uint32 per rank, zero-initialised once
 *   mc_partial / mc_result       multicast addresses of the two buffers, or both NULL (peer-to-peer path)
 *   max_ctas         0 = one CTA per SM; smaller values leave SMs to concurrent kernels
 *   fold_iters / red_iters   float4 per thread of a fold / reduce work item (0 = defaults); the same values
 *                    must be used for every call on one set of flag words
 * The result buffer of every rank holds [ mean parameters | tail sums | fp64 sums ] on return
 * (bit-identical on all ranks).  Collective: every rank launches it with the same shapes.            */
size_t fmlp_fedavg_allreduce_q_buffer_floats(int64_t P, int64_t T, int M);
int fmlp_fedavg_allreduce_q_f32(const float* const* srcs, const float* const* tail_srcs, const float* weights,
                                int K, int64_t P, int64_t T, const double* tail_f64, int M,
                                float* const* partial_ptrs, float* const* result_ptrs,
                                uint32_t* const* flag_ptrs, float* mc_partial, float* mc_result, int n_chunks,
                                int rank, int world, uint32_t* epoch_dev, int max_ctas, int fold_iters,
                                int red_iters, fmlp_stream_t stream);
/* This rank's fp64 partial sums for the exchange above, M = 3C + J (layout in fedavg_allreduce_q.cu):
 * class weight sums over the annotating clients (FedAvg_proto, utils/FedAvg.py:72-93), n*t sums and weight
 * sums over the clients that miss the class (FedAvg_tao, :51-70; t = tcnt / rows as at
 * utils/local_training.py:1000,1249), and the weighted int64 BatchNorm counter sums (FedAvg.py:9-13).
 *   tcnt device [S][C] int32 (or NULL); weights / rows / act / neg host [S]; counters host [S] device ptrs */
int fmlp_agg_tail_pack_f64(const int32_t* tcnt, int S, int C, const double* weights, const int64_t* rows,
                           const uint32_t* act, const uint32_t* neg, const int64_t* const* counters, int J,
                           double* out, fmlp_stream_t stream);
/* After the exchange: prototypes = tail sums * (total / class weight) (0 * inf = NaN where nobody annotates
 * the class, like the reference's 0/0), tao = num / den (1.0 for an empty list), counters as float32.   */
int fmlp_agg_finalize_f32(const float* proto_sum, const double* tail, int C, int D, int J, double total_weight,
                          float* proto_out, double* tao_out, float* counters_out, fmlp_stream_t stream);

/* Prototype aggregation, replaces utils/FedAvg.py:72-93 `FedAvg_proto`:
 *   out[2c+j] = (sum over clients i in act(c), in list order, of protos[i][2c+j]*n_i) / sum n_i
 * protos  device [K][2C][D] (stacked client prototypes);  weights host [K] (double);
 * class_clients  host [C] bit masks over clients (bit i set = client i annotates class c),
 *                K <= 64.  Empty act(c) gives 0/0 = NaN exactly like the reference.        */
int fmlp_proto_avg_f32(const float* protos, int K, int C, int D, int rows_per_class,
                       const double* weights, const uint64_t* class_clients, float* out,
                       fmlp_stream_t stream);
/* rows_per_class = 2: the layout above (FedAvg_proto).  rows_per_class = 1: protos [K][C][D], one row
 * per class = utils/FedAvg.py:95-103 `FedAvg_rela`.                                            */

/* All the small tails of the single-GPU server aggregation (main.py:218-234) in ONE launch: FedAvg_proto
 * (utils/FedAvg.py:72-93, as fmlp_proto_avg_f32 with rows_per_class = 2), FedAvg_tao over the clients that miss
 * each class (:51-70, IEEE double in the reference's order; t_k[c] = tcnt[k][c] / rows[k] as at
 * utils/local_training.py:1000,1249; 1.0 for an empty list) and the int64 BatchNorm counters
 * (FedAvg.py:9-13: sum_k n_k * counter_k, then the float32 divide by total_weight).
 *   protos device [K][2C][D]; weights / rows / neg (missing-class masks) host [K]; class_clients host [C];
 *   tcnt device [K][C] int32; counters host [K] device pointers to J int64 each (NULL iff J == 0);
 *   proto_out device [2C][D]; tao_out device [C] double (NULL = skip); counters_out device [J] float32.   */
int fmlp_agg_tails_local_f32(const float* protos, int K, int C, int D, const double* weights,
                             const uint64_t* class_clients, float* proto_out, const int32_t* tcnt,
                             const int64_t* rows, const uint32_t* neg, const int64_t* const* counters, int J,
                             double total_weight, double* tao_out, float* counters_out, fmlp_stream_t stream);

/* Difficulty aggregation, replaces utils/FedAvg.py:51-70 `FedAvg_tao` (float64, bit-identical:
 * same operation order, IEEE double mul/add/div):
 *   class_clients != NULL: out[c] = sum_{i in list(c)} t[i][c]*w_i / sum w_i, 1.0 for an empty list
 *   class_clients == NULL: out[c] = sum_i t[i][c]*w_i / total_weight
 * t_dev device [K][C] double; weights host [K]; out_dev device [C] double.                     */
int fmlp_tao_avg_f64(const double* t_dev, int K, int C, const double* weights,
                     const uint64_t* class_clients, double total_weight, double* out_dev,
                     fmlp_stream_t stream);

/* Model distance, replaces utils/FedNoRo.py:106-115 / utils/FedAvg.py:42-49 `model_dist`:
 *   out[0] = sum over the T float tensors, in table order, of || a_t - b_t ||_2
 * (int64 tensors are skipped by the caller, as FedNoRo.py:110-111 does).  Tables as in
 * fmlp_fedavg_multi_f32 plus tensor_chunk0_dev [T+1] = index of every tensor's first chunk.   */
size_t fmlp_model_dist_ws_bytes(int64_t n_chunks, int T);
int fmlp_model_dist_f32(const float* const* a_table_dev, const float* const* b_table_dev,
                        const int64_t* numel_dev, const int32_t* chunk_tensor_dev,
                        const int64_t* chunk_start_dev, const int64_t* tensor_chunk0_dev,
                        int64_t n_chunks, int T, float* out, void* ws, size_t ws_bytes,
                        fmlp_stream_t stream);

/* ------------------------------------------------------------------ K2: class prototypes
 * Replaces utils/local_training.py:973-1000 (stage 1) and :1208-1249 (stage 2):
 * for every active class c of the segment, rows with labels[n,c]==0 are summed into row 2c and
 * rows with labels[n,c]==1 into row 2c+1 (other label values contribute to neither), rows are
 * counted, and for every class in the segment's `tcount` mask the number of confident
 * predictions #{n : p<L or p>U}, p = sigmoid(logits[n,c]), is counted (:995-996,1239).
 *
 *   feat      [N_total, D] fp32, leading dimension ld_feat (elements), D % 4 == 0, 16-B aligned
 *   labels    [N_total, C] fp32 (0/1)        logits  [N_total, C] fp32 or NULL (no t counts)
 *   logits_are_probs != 0: `logits` already holds sigmoid outputs
 *   seg_rows  host [S+1];  seg_active / seg_tcount  host [S] class masks
 *   proto     [S, 2C, D] fp32 out: means (rows of non-active classes = 0)
 *   cnt       [S, 2C] int32 out           tcnt  [S, C] int32 out (may be NULL iff logits NULL)
 *   guard_empty != 0: an empty (class,label) group leaves its row 0 (stage 2, :1241-1248);
 *                == 0: it is 0/0 = NaN (stage 1, :997-999)
 * Deterministic: per-chunk partial sums are combined in a fixed order.                    */
size_t fmlp_proto_ws_bytes(int64_t n_total, int D, int C, int S);
int fmlp_proto_build_f32(const float* feat, int64_t ld_feat, int D, const float* labels,
                         const float* logits, int logits_are_probs, int C, int S,
                         const int64_t* seg_rows, const uint32_t* seg_active,
                         const uint32_t* seg_tcount, float L, float U, int guard_empty,
                         float* proto, int32_t* cnt, int32_t* tcnt, void* ws, size_t ws_bytes,
                         fmlp_stream_t stream);

/* ------------------------------------------------------------------ K3: tagging similarity
 * Replaces utils/local_training.py:1052-1058 + CosineSimilarityFast.forward :1417-1435:
 *   sim[c][n] = cos(f_n, P[2c]) - cos(f_n, P[2c+1]),   cos(f,p) = (f.p) * (1 / (|f|*|p|))
 * for every class c in the row's segment `missing` mask, in ONE pass over feat.
 *   proto   [2C, D] fp32 (the server-aggregated prototypes, shared by all segments)
 *   sim     [C, N_total] fp32, class-major; rows of classes outside a segment's mask are
 *           left untouched
 *   mode    FMLP_SIM_PAIR: two cosines then subtract (op order of the reference)
 *           FMLP_SIM_FOLDED: one dot against q_c = P[2c]/|P[2c]| - P[2c+1]/|P[2c+1]|
 *           (half the FMAs; differs from PAIR by O(1e-7), inside the north_star waiver)
 *   ws      fmlp_tag_sim_ws_bytes(C, D) bytes, 16-byte aligned: the class-vector table a small
 *           pre-kernel builds once per call (prototype norms, folded vectors) and every CTA of
 *           the streaming kernel copies; the two launches are chained by programmatic
 *           dependent launches (the pre-kernel to whatever precedes it in the stream, the
 *           streaming kernel to the pre-kernel; every global read sits behind a
 *           griddepcontrol.wait, so the call is stream-ordered like any other)
 * The streaming kernel is persistent (one CTA per SM: 8 compute warps fed by 2-D TMA boxes +
 * 2 epilogue warps) and takes 160-227 KB of shared memory per CTA.                           */
enum { FMLP_SIM_PAIR = 0, FMLP_SIM_FOLDED = 1 };
size_t fmlp_tag_sim_ws_bytes(int C, int D);
int fmlp_tag_sim_f32(const float* feat, int64_t ld_feat, int D, const float* proto, int C,
                     int S, const int64_t* seg_rows, const uint32_t* seg_missing, float* sim,
                     int64_t ld_sim, int mode, void* ws, size_t ws_bytes, fmlp_stream_t stream);

/* ------------------------------------------------------------------ K3 fused into feature extraction
 * (SURVEY §8f.1).  The reference's tagging pass keeps, per batch, only the pooled feature of the
 * backbone's last feature map, `features, _ = net(images1)` (utils/local_training.py:1033-1036;
 * the model tail is flatten(adaptive_avg_pool2d(relu(fmap), 1))), concatenates them over the
 * dataset (:1037, quadratic torch.cat) and scores the result (:1052-1058).  fmlp_pool_tag_f32 does
 * pooling and scoring in one pass over the feature map of a batch:
 *     feat[b, d] = (1/HW) * sum_hw act(fmap[b, d, hw])      act = relu if `relu` != 0, else identity
 *     sim[c][b]  = cos(feat_b, P[2c]) - cos(feat_b, P[2c+1])   for every class c in `classes`
 *   fmap     [B, D, HW] (FMLP_FMAP_NCHW) or [B, HW, D] (FMLP_FMAP_NHWC, channels_last), dense, 16-B
 *            aligned, D % 4 == 0, HW <= 256
 *   table    class vectors built once per round by fmlp_sim_table_build_f32 from the aggregated
 *            prototypes [2C, D] with the SAME `classes` and `mode`; fmlp_sim_table_bytes(C, D) bytes
 *   feat     [B, ld_feat] out (what the classifier consumes) or NULL
 *   sim      [C, ld_sim] class-major, already offset to the batch's first column; column b is
 *            written for the classes in `classes`; NULL iff classes == 0 (pooling only)
 *   mode     FMLP_SIM_PAIR (<= 16 classes per launch) or FMLP_SIM_FOLDED                       */
enum { FMLP_FMAP_NCHW = 0, FMLP_FMAP_NHWC = 1 };
size_t fmlp_sim_table_bytes(int C, int D);
int fmlp_sim_table_build_f32(const float* proto, int C, int D, uint32_t classes, int mode,
                             float* table, fmlp_stream_t stream);
int fmlp_pool_tag_f32(const float* fmap, int layout, int B, int D, int HW, int relu,
                      const float* table, int C, uint32_t classes, int mode, float* feat,
                      int64_t ld_feat, float* sim, int64_t ld_sim, fmlp_stream_t stream);

/* ------------------------------------------------------------------ K3b: selection
 * Replaces utils/local_training.py:1061-1112 + utils/utils.py:24-35 (max_m_indices /
 * min_n_indices).  Per (segment, missing class), over the candidate rows (tag == 0):
 *   clean = {sim >= 0}, noise = {sim < 0}  (NaN in neither)
 *   m = (int)(clean_frac*|clean|), k = (int)(noise_frac*|noise|)     (double arithmetic)
 *   pick the m largest sims of `clean` and the k smallest of `noise`; equal sims are taken in
 *   increasing row order (Python's stable sort); results are emitted in rank order.
 * The picked rows are marked in `tag` (1 = clean/confident-negative, 2 = noise/confident-
 * positive), which is also what makes them non-candidates next round (:1197-1204).
 *   tag       [C, N_total] uint8, class-major, in/out
 *   counts    [S, C, 4] int32 out: n_clean, n_noise, m, k   (zeros for non-missing classes)
 *   remaining [S, C] int32 out or NULL: rows that are still candidates (tag == 0) after this
 *             selection = the number of `distill` entries of that (client, class), :1467-1468
 *   sel       [S, C, 2, cap] int32 out: global row numbers, side 0 = clean, 1 = noise
 *   cap       per-(segment,class,side) capacity; must be >= the largest possible m or k
 * ws: fmlp_tag_select_ws_bytes(S, C, cap).
 * Launch shape: one 1024-thread CTA per (segment, class) item while the largest segment has at most 16,384 rows
 * (keys in registers); beyond that one thread-block cluster of 2 / 4 / 8 CTAs per item (histograms merged through
 * distributed shared memory).  FMLP_TUNE_SELECT_CLUSTER forces the split.  Programmatic dependent launch behind
 * the preceding kernel of the stream (all reads sit behind griddepcontrol.wait).                */
size_t fmlp_tag_select_ws_bytes(int S, int C, int64_t cap);
int fmlp_tag_select(const float* sim, int64_t ld_sim, uint8_t* tag, int64_t ld_tag, int C,
                    int S, const int64_t* seg_rows, const uint32_t* seg_missing,
                    double clean_frac, double noise_frac, int32_t* counts, int32_t* remaining,
                    int32_t* sel, int64_t cap, void* ws, size_t ws_bytes, fmlp_stream_t stream);

/* ------------------------------------------------------------------ K3c: label / mask fill
 * Replaces DatasetSplit_pseudo.__getitem__ (utils/local_training.py:1456-1477) and
 * `sup_cls = ~distill_cls` (:1173), for all rows at once:
 *   active class  : y = labels_in, distill = 0
 *   missing class : y = (tag == 2), distill = (tag == 0)
 *   other classes : y = 0, distill = 0
 *   sup = 1 - distill.      Any of y / distill / sup may be NULL.                          */
int fmlp_mask_fill(const float* labels_in, const uint8_t* tag, int64_t ld_tag, int C, int S,
                   const int64_t* seg_rows, const uint32_t* seg_active,
                   const uint32_t* seg_missing, float* y, float* distill, float* sup,
                   fmlp_stream_t stream);

/* ------------------------------------------------------------------ K4: fused losses
 * Stage 1, replaces utils/local_training.py:933-963 (+ utils/FedNoRo.py:16-22):
 *   p_i = sigmoid(z_i)
 *   loss = sum_{n, c in active} 0.5*(BCE(p1,y)+BCE(p2,y)) / (bs*A)
 *        + sum_{n, c in missing} 0.5*((p1-p3)^2+(p2-p4)^2) / (bs*M)
 *   BCE(p,y) = -(y*max(log p,-100) + (1-y)*max(log(1-p),-100))    (F.binary_cross_entropy)
 *   dz = autograd of the above: BCE backward (p-y)/max((1-p)*p,1e-12) then sigmoid backward.
 * `bs` is args.batch_size (NOT the number of rows B; :956-959).  A/M are popcounts of the masks.
 * z1,z2 student logits (two views), z3,z4 frozen global model logits, all [B, C].
 * loss: device float[1] out; dz1, dz2: [B, C] out (gradient of loss w.r.t. z1, z2).         */
size_t fmlp_loss_ws_bytes(int64_t B, int C);
int fmlp_loss_stage1_f32(const float* z1, const float* z2, const float* z3, const float* z4,
                         const float* y, int64_t B, int C, uint32_t active, uint32_t missing,
                         int bs, float* loss, float* dz1, float* dz2, void* ws,
                         size_t ws_bytes, fmlp_stream_t stream);

/* Stage 2, replaces utils/local_training.py:1171-1188:
 *   variant FMLP_LOSS2_SUP      : loss = sum(BCE(p,y)*sup) / sum(sup)               (:1188, live)
 *   variant FMLP_LOSS2_SUP_DIS  : loss = (sum(BCE*sup) + sum((p-pg)^2*distill))
 *                                        / (sum(sup) + sum(distill))        (:1187, commented)
 *   sup = 1 - distill (:1173), p = sigmoid(z), pg = sigmoid(zg).  zg may be NULL for _SUP.  */
enum { FMLP_LOSS2_SUP = 0, FMLP_LOSS2_SUP_DIS = 1 };
int fmlp_loss_stage2_f32(const float* z, const float* zg, const float* y, const float* distill,
                         int64_t B, int C, int variant, float* loss, float* dz, void* ws,
                         size_t ws_bytes, fmlp_stream_t stream);

/* Segmented form: S clients' rows stored back to back, each with its own denominator and its own
 * scalar loss, in one launch.  loss: device float[S].
 * seg_class_distill: device int32 [S, C] = number of distill==1 entries per (client, class) (the
 * `remaining` output of fmlp_tag_select), or NULL to have the kernel count them itself (one more
 * pass over `distill` and one more grid barrier).                                               */
int fmlp_loss_stage2_seg_f32(const float* z, const float* zg, const float* y, const float* distill,
                             int C, int S, const int64_t* seg_rows, int variant,
                             const int32_t* seg_class_distill, float* loss, float* dz, void* ws,
                             size_t ws_bytes, fmlp_stream_t stream);

/* Label / mask fill fused into the segmented stage-2 loss (utils/local_training.py:1456-1477 + :1171-1188) in ONE
 * ordinary launch: y / distill / sup (any may be NULL) are written as by-products, the denominators come from
 * `seg_class_distill` = fmlp_tag_select's `remaining` counts [S][C], the per-client losses are reduced by the CTA
 * that finishes last.  `ws`: fmlp_loss_ws_bytes bytes that were ZERO when first used (the kernel leaves its
 * arrival counter zero); chained to the preceding fmlp_tag_select by a programmatic dependent launch.       */
int fmlp_fill_loss_stage2_f32(const float* labels, const uint8_t* tag, int64_t ld_tag, const float* z,
                              const float* zg, int C, int S, const int64_t* seg_rows,
                              const uint32_t* seg_active, const uint32_t* seg_missing, int variant,
                              const int32_t* seg_class_distill, float* y, float* distill, float* sup,
                              float* loss, float* dz, void* ws, size_t ws_bytes, fmlp_stream_t stream);

/* ------------------------------------------------------------------ fused Adam
 * One torch.optim.Adam step (amsgrad=False; the optimizer the reference re-creates every round,
 * utils/local_training.py:912,1149) over the PARAMETER runs of flat fp32 buffers p/g/m/v that
 * share one layout.  chunk_start_dev / chunk_len_dev (device arrays [n_chunks], len <= 2048)
 * enumerate those runs; `step` is the 1-based step count; zero_grad != 0 also clears g.         */
int fmlp_adam_step_f32(float* p, float* g, float* m, float* v, const int64_t* chunk_start_dev,
                       const int32_t* chunk_len_dev, int64_t n_chunks, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int64_t step, int zero_grad,
                       fmlp_stream_t stream);

/* ------------------------------------------------------------------ evaluation metrics
 * Replaces the metric half of utils/evaluations.py:15-73 `globaltest` / :89-140 `classtest`: per class,
 * sklearn average_precision_score and roc_curve + auc of the probabilities, and the integer sums behind
 * utils/multilabel_metrixs.py (BACC, Recall, Precision, F1Measure, Hamming_Loss) on preds = probs > threshold.
 *   scores  [N, C] fp32 logits (scores_are_probs == 0: p = sigmoid(z) in fp32) or probabilities
 *   labels  [N, C] fp32 0/1
 *   class_counts  device int32 [C][8] out: n_pos, n_neg, n_pred, tp, tn, mismatches, 0, 0
 *   class_ap_auc  device double [C][2] out: AP_c, AUC_c (NaN for a class without positives / negatives)
 * Exact (no sort: P_c x N pairwise counts), deterministic; ws: fmlp_eval_ws_bytes(N, C).          */
size_t fmlp_eval_ws_bytes(int64_t N, int C);
int fmlp_eval_multilabel_f32(const float* scores, const float* labels, int64_t N, int C,
                             int scores_are_probs, float threshold, int32_t* class_counts,
                             double* class_ap_auc, void* ws, size_t ws_bytes, fmlp_stream_t stream);

/* dz[i] *= *scale_dev  (upstream gradient of the scalar loss, read from device memory). */
int fmlp_scale_f32(float* x, int64_t n, const float* scale_dev, fmlp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FEDMLP_B200_H_ */
