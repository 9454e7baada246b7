// eval.cu — SURVEY §8(f).4: on-device evaluation metrics of the global model.
//
// Replaces the post-loop part of utils/evaluations.py:15-73 `globaltest` (and :89-140 `classtest`):
// the reference moves all probabilities to the host and calls, per class, sklearn's
// average_precision_score and roc_curve + auc (sort based), plus the count metrics of
// utils/multilabel_metrixs.py (BACC, Recall, Precision, F1Measure, Hamming_Loss) on
// preds = probs > 0.5.  Here everything stays on the device and nothing is sorted: both ranking
// metrics are sums over the POSITIVE samples of a class of quantities that only need counts,
//     AP_c  = (1/P) * sum_{i pos} TP(p_i) / (TP(p_i) + FP(p_i)),   TP(t) = #{pos j: p_j >= t}, FP(t) = #{neg j: p_j >= t}
//     AUC_c = (1/(P*Nn)) * sum_{i pos} ( #{neg j: p_j < p_i} + 0.5 * #{neg j: p_j == p_i} )
// (sklearn's step-wise AP over distinct thresholds and the trapezoidal ROC area with ties, written per
// positive sample), so P_c x N comparisons per class replace the sort: ~0.9 G compares for the 25,596 x 14
// ChestX-ray14 test set, a few tens of microseconds of integer work.
//
//   eval_prepare_kernel   one CTA per class: p = sigmoid(z) (fp32, as torch.sigmoid), pred = p > th,
//                         the count metrics' integer sums, and an order-preserving split of the class's
//                         probabilities into a positive and a negative list (block scan, deterministic)
//   eval_pair_kernel      thread <-> positive sample; both lists stream through shared memory
//                         (broadcast LDS.128), three counters per thread, float64 terms out
//   eval_finalize_kernel  one CTA per class: fixed-order float64 sums -> AP_c, AUC_c
// The last arithmetic (means over classes, divisions with numpy's zero-division behaviour) is done by
// the host mirror on C numbers, literally as the reference does it.
#include "common.cuh"

namespace fmlp {

constexpr int kEvalPrepThreads = 1024;
constexpr int kEvalPairThreads = 256;
constexpr int kEvalTile = 2048;       // list elements staged per shared-memory tile
constexpr int kEvalCounts = 8;        // int32 per class: n_pos, n_neg, n_pred, tp, tn, mismatches, pad, pad

struct EvalWs {
    float* pos_p;      // [C][N] probabilities of the class's positives, original order
    float* neg_p;      // [C][N]
    double* term_ap;   // [C][N] per positive (list order)
    double* term_auc;  // [C][N]
};

__device__ __forceinline__ int block_excl_scan_1024(int flag, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned b = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_warp[warp] = __popc(b);
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll 8
    for (int w = 0; w < kEvalPrepThreads / 32; ++w) {
        const int v = s_warp[w];
        before += (w < warp) ? v : 0;
        tot += v;
    }
    __syncthreads();
    total = tot;
    return before + __popc(b & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(kEvalPrepThreads) eval_prepare_kernel(const float* __restrict__ scores,
                                                                        const float* __restrict__ labels, int64_t N, int C,
                                                                        int scores_are_probs, float th, EvalWs ws,
                                                                        int32_t* __restrict__ counts) {
    __shared__ int s_warp[kEvalPrepThreads / 32];
    __shared__ int s_red[6][kEvalPrepThreads / 32];
    const int c = blockIdx.x;
    float* pos_p = ws.pos_p + (int64_t)c * N;
    float* neg_p = ws.neg_p + (int64_t)c * N;
    int base_pos = 0, base_neg = 0;
    int n_pred = 0, tp = 0, tn = 0, mism = 0;
    for (int64_t n0 = 0; n0 < N; n0 += kEvalPrepThreads) {
        const int64_t n = n0 + threadIdx.x;
        const bool ok = n < N;
        float p = 0.f;
        bool y = false;
        if (ok) {
            const float z = scores[n * C + c];
            p = scores_are_probs ? z : sigmoid_ref(z);
            y = labels[n * C + c] != 0.f;          // np.logical_and / == on 0/1 labels
            const bool pred = p > th;              // preds = probs > accuracy_th  (:27-28)
            n_pred += pred;
            tp += (y && pred);
            tn += (!y && !pred);
            mism += (y != pred);
        }
        int tot_pos, tot_neg;
        const int slot_pos = block_excl_scan_1024(ok && y, s_warp, tot_pos);
        const int slot_neg = block_excl_scan_1024(ok && !y, s_warp, tot_neg);
        if (ok) {
            if (y) pos_p[base_pos + slot_pos] = p; else neg_p[base_neg + slot_neg] = p;
        }
        base_pos += tot_pos;
        base_neg += tot_neg;
    }
    int v[4] = {n_pred, tp, tn, mism};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = warp_sum_i(v[k]);
        if ((threadIdx.x & 31) == 0) s_red[k][threadIdx.x >> 5] = r;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        int r = 0;
        for (int w = 0; w < kEvalPrepThreads / 32; ++w) r += s_red[threadIdx.x][w];
        counts[c * kEvalCounts + 2 + threadIdx.x] = r;
    }
    if (threadIdx.x == 0) { counts[c * kEvalCounts + 0] = base_pos; counts[c * kEvalCounts + 1] = base_neg; }
}

// grid (ceil(N / 256), C); CTAs past the class's positive count exit.
__global__ void __launch_bounds__(kEvalPairThreads) eval_pair_kernel(int64_t N, EvalWs ws, const int32_t* __restrict__ counts) {
    __shared__ __align__(16) float s_tile[kEvalTile];
    const int c = blockIdx.y;
    const int n_pos = counts[c * kEvalCounts + 0], n_neg = counts[c * kEvalCounts + 1];
    const int q0 = blockIdx.x * kEvalPairThreads;
    if (q0 >= n_pos) return;
    const float* pos_p = ws.pos_p + (int64_t)c * N;
    const float* neg_p = ws.neg_p + (int64_t)c * N;
    const int q = q0 + threadIdx.x;
    const bool mine = q < n_pos;
    const float pi = mine ? pos_p[q] : 2.f;    // 2 > every probability: idle lanes count nothing
    int ge_pos = 0, ge_neg = 0, gt_neg = 0;
    // positives: TP(p_i)
    for (int j0 = 0; j0 < n_pos; j0 += kEvalTile) {
        __syncthreads();
        for (int j = threadIdx.x; j < kEvalTile; j += kEvalPairThreads) s_tile[j] = (j0 + j < n_pos) ? pos_p[j0 + j] : -1.f;
        __syncthreads();
        const int lim = min(kEvalTile, (n_pos - j0 + 3) & ~3);
        for (int j = 0; j < lim; j += 4) {
            const float4 v = *reinterpret_cast<const float4*>(s_tile + j);
            ge_pos += (v.x >= pi) + (v.y >= pi) + (v.z >= pi) + (v.w >= pi);
        }
    }
    // negatives: FP(p_i) and the strict count for the tie correction of the ROC area
    for (int j0 = 0; j0 < n_neg; j0 += kEvalTile) {
        __syncthreads();
        for (int j = threadIdx.x; j < kEvalTile; j += kEvalPairThreads) s_tile[j] = (j0 + j < n_neg) ? neg_p[j0 + j] : -1.f;
        __syncthreads();
        const int lim = min(kEvalTile, (n_neg - j0 + 3) & ~3);
        for (int j = 0; j < lim; j += 4) {
            const float4 v = *reinterpret_cast<const float4*>(s_tile + j);
            ge_neg += (v.x >= pi) + (v.y >= pi) + (v.z >= pi) + (v.w >= pi);
            gt_neg += (v.x > pi) + (v.y > pi) + (v.z > pi) + (v.w > pi);
        }
    }
    if (mine) {
        ws.term_ap[(int64_t)c * N + q] = (double)ge_pos / (double)(ge_pos + ge_neg);
        ws.term_auc[(int64_t)c * N + q] = (double)(n_neg - ge_neg) + 0.5 * (double)(ge_neg - gt_neg);
    }
}

__global__ void __launch_bounds__(256) eval_finalize_kernel(int64_t N, EvalWs ws, const int32_t* __restrict__ counts,
                                                            double* __restrict__ ap_auc) {
    __shared__ double s_a[256], s_b[256];
    const int c = blockIdx.x;
    const int n_pos = counts[c * kEvalCounts + 0], n_neg = counts[c * kEvalCounts + 1];
    double a = 0.0, b = 0.0;
    for (int q = threadIdx.x; q < n_pos; q += 256) {      // fixed assignment, fixed tree: deterministic
        a += ws.term_ap[(int64_t)c * N + q];
        b += ws.term_auc[(int64_t)c * N + q];
    }
    s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) { s_a[threadIdx.x] += s_a[threadIdx.x + w]; s_b[threadIdx.x] += s_b[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // no positives: 0/0 = NaN for both (sklearn warns and returns nan / 0 there; the host mirror says so)
        ap_auc[2 * c] = s_a[0] / (double)n_pos;
        ap_auc[2 * c + 1] = s_b[0] / ((double)n_pos * (double)n_neg);
    }
}

static size_t eval_align(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace fmlp

using namespace fmlp;

extern "C" size_t fmlp_eval_ws_bytes(int64_t N, int C) {
    if (N < 0 || C < 1) return 0;
    const size_t cn = (size_t)C * (size_t)(N > 0 ? N : 1);
    return 2 * eval_align(cn * sizeof(float)) + 2 * eval_align(cn * sizeof(double));
}

extern "C" int fmlp_eval_multilabel_f32(const float* scores, const float* labels, int64_t N, int C,
                                        int scores_are_probs, float threshold, int32_t* class_counts,
                                        double* class_ap_auc, void* ws, size_t ws_bytes, fmlp_stream_t stream) {
    if (!scores || !labels || !class_counts || !class_ap_auc || N < 1 || N > 0x7fffffff / 2 || C < 1 || C > FMLP_MAX_CLASSES)
        return FMLP_ERR_BAD_ARG;
    if (!ws || ws_bytes < fmlp_eval_ws_bytes(N, C)) return FMLP_ERR_WORKSPACE;
    const size_t cn = (size_t)C * (size_t)N;
    char* p = reinterpret_cast<char*>(ws);
    EvalWs w;
    w.pos_p = reinterpret_cast<float*>(p); p += eval_align(cn * sizeof(float));
    w.neg_p = reinterpret_cast<float*>(p); p += eval_align(cn * sizeof(float));
    w.term_ap = reinterpret_cast<double*>(p); p += eval_align(cn * sizeof(double));
    w.term_auc = reinterpret_cast<double*>(p);
    cudaStream_t st = (cudaStream_t)stream;
    eval_prepare_kernel<<<C, kEvalPrepThreads, 0, st>>>(scores, labels, N, C, scores_are_probs, threshold, w, class_counts);
    int rc = launch_status();
    if (rc != FMLP_OK) return rc;
    const unsigned gx = (unsigned)((N + kEvalPairThreads - 1) / kEvalPairThreads);
    eval_pair_kernel<<<dim3(gx, (unsigned)C), kEvalPairThreads, 0, st>>>(N, w, class_counts);
    rc = launch_status();
    if (rc != FMLP_OK) return rc;
    eval_finalize_kernel<<<C, 256, 0, st>>>(N, w, class_counts, class_ap_auc);
    return launch_status();
}
