// fedavg.cu — K1: weighted averaging of K client parameter buffers (server aggregation).
//
// Replaces utils/FedAvg.py:7-14 `FedAvg` / :16-23 `Fed_w` / :72-93 `FedAvg_proto` of the
// reference.  Pure HBM streaming: (K+1)*4*P algorithmic bytes, ~2K+1 flops per element, so the
// only thing that matters is keeping enough 128-bit loads in flight on every SM.  Because the
// arithmetic is free, the fold is done in the reference's exact order with separately rounded
// multiply / add and an IEEE divide -> bit-identical to the reference evaluated on CPU tensors.
#include "common.cuh"

namespace fmlp {

struct FedAvgArgs {
    const float* src[FMLP_MAX_CLIENTS];
    float w[FMLP_MAX_CLIENTS];
    float* out;
    int64_t P;
    float divisor;
    int K;
    int flags;
};

constexpr int kFedAvgThreads = 256;
constexpr int kFedAvgUnroll = 8;  // client loads in flight per thread (8 x 16 B)

// acc <- acc + s*w with the reference's rounding (no FMA contraction).
__device__ __forceinline__ void fold(float4& acc, const float4& s, float w) {
    acc.x = __fadd_rn(acc.x, __fmul_rn(s.x, w));
    acc.y = __fadd_rn(acc.y, __fmul_rn(s.y, w));
    acc.z = __fadd_rn(acc.z, __fmul_rn(s.z, w));
    acc.w = __fadd_rn(acc.w, __fmul_rn(s.w, w));
}

// One float4 column of the K-way fold.  `first` = index of the first client to fold
// (0 when accumulating onto `acc`, 1 when acc was initialised from client 0).
template <typename SrcAt>
__device__ __forceinline__ void fold_clients(float4& acc, int first, int K, const float* w, SrcAt src_at) {
    int i = first;
    for (; i + kFedAvgUnroll <= K; i += kFedAvgUnroll) {
        float4 v[kFedAvgUnroll];
#pragma unroll
        for (int u = 0; u < kFedAvgUnroll; ++u) v[u] = ld_stream_f4(src_at(i + u));
#pragma unroll
        for (int u = 0; u < kFedAvgUnroll; ++u) fold(acc, v[u], w[i + u]);
    }
    if (i < K) {
        float4 v[kFedAvgUnroll];
#pragma unroll
        for (int u = 0; u < kFedAvgUnroll; ++u)
            if (i + u < K) v[u] = ld_stream_f4(src_at(i + u));
#pragma unroll
        for (int u = 0; u < kFedAvgUnroll; ++u)
            if (i + u < K) fold(acc, v[u], w[i + u]);
    }
}

template <bool kVec>
__global__ void __launch_bounds__(kFedAvgThreads, 4)
fedavg_flat_kernel(const __grid_constant__ FedAvgArgs a) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const bool accumulate = (a.flags & FMLP_FEDAVG_ACCUMULATE) != 0;
    const bool divide = (a.flags & FMLP_FEDAVG_DIVIDE) != 0;

    if (kVec) {
        const int64_t nvec = a.P >> 2;
        for (int64_t v = tid; v < nvec; v += nthreads) {
            const int64_t e = v << 2;
            float4 acc;
            int first;
            if (accumulate) {
                acc = *reinterpret_cast<const float4*>(a.out + e);
                first = 0;
            } else {
                float4 s0 = ld_stream_f4(a.src[0] + e);
                acc = make_float4(__fmul_rn(s0.x, a.w[0]), __fmul_rn(s0.y, a.w[0]),
                                  __fmul_rn(s0.z, a.w[0]), __fmul_rn(s0.w, a.w[0]));
                first = 1;
            }
            fold_clients(acc, first, a.K, a.w, [&](int i) { return a.src[i] + e; });
            if (divide) {
                acc.x = __fdiv_rn(acc.x, a.divisor);
                acc.y = __fdiv_rn(acc.y, a.divisor);
                acc.z = __fdiv_rn(acc.z, a.divisor);
                acc.w = __fdiv_rn(acc.w, a.divisor);
            }
            st_stream_f4(a.out + e, acc);
        }
    }
    // scalar part: everything when !kVec, else the P % 4 tail
    const int64_t s_begin = kVec ? (a.P & ~(int64_t)3) : 0;
    for (int64_t e = s_begin + tid; e < a.P; e += nthreads) {
        float acc;
        int first;
        if (accumulate) { acc = a.out[e]; first = 0; }
        else { acc = __fmul_rn(a.src[0][e], a.w[0]); first = 1; }
        for (int i = first; i < a.K; ++i) acc = __fadd_rn(acc, __fmul_rn(a.src[i][e], a.w[i]));
        if (divide) acc = __fdiv_rn(acc, a.divisor);
        a.out[e] = acc;
    }
}

// ---------------------------------------------------------------- multi-tensor form
struct FedAvgMultiArgs {
    const float* const* src_table;  // [T*K]
    float* const* dst_table;        // [T]
    const int64_t* numel;           // [T]
    const int32_t* chunk_tensor;    // [n_chunks]
    const int64_t* chunk_start;     // [n_chunks]
    int64_t n_chunks;
    float w[FMLP_MAX_CLIENTS];
    float divisor;
    int K;
    int flags;
};

constexpr int kMultiThreads = 256;  // 2 float4 per thread per chunk of 2048

__global__ void __launch_bounds__(kMultiThreads, 4)
fedavg_multi_kernel(const __grid_constant__ FedAvgMultiArgs a) {
    __shared__ const float* s_src[FMLP_MAX_CLIENTS];
    __shared__ int s_vec_ok;
    const bool accumulate = (a.flags & FMLP_FEDAVG_ACCUMULATE) != 0;
    const bool divide = (a.flags & FMLP_FEDAVG_DIVIDE) != 0;

    for (int64_t c = blockIdx.x; c < a.n_chunks; c += gridDim.x) {
        const int t = a.chunk_tensor[c];
        const int64_t start = a.chunk_start[c];
        const int64_t numel = a.numel[t];
        float* dst = a.dst_table[t];
        __syncthreads();  // previous iteration done with s_src
        if (threadIdx.x == 0) s_vec_ok = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
        __syncthreads();
        if (threadIdx.x < a.K) {
            const float* p = a.src_table[(int64_t)t * a.K + threadIdx.x];
            s_src[threadIdx.x] = p;
            if (reinterpret_cast<uintptr_t>(p) & 15u) s_vec_ok = 0;  // benign race: all writers store 0
        }
        __syncthreads();
        const int64_t len = min((int64_t)FMLP_FEDAVG_CHUNK, numel - start);
        const bool vec_ok = s_vec_ok != 0;  // chunk starts are multiples of 2048 -> keep alignment
        const int64_t nvec = vec_ok ? (len >> 2) : 0;
        for (int64_t v = threadIdx.x; v < nvec; v += kMultiThreads) {
            const int64_t e = start + (v << 2);
            float4 acc;
            int first;
            if (accumulate) { acc = *reinterpret_cast<const float4*>(dst + e); first = 0; }
            else {
                float4 s0 = ld_stream_f4(s_src[0] + e);
                acc = make_float4(__fmul_rn(s0.x, a.w[0]), __fmul_rn(s0.y, a.w[0]),
                                  __fmul_rn(s0.z, a.w[0]), __fmul_rn(s0.w, a.w[0]));
                first = 1;
            }
            fold_clients(acc, first, a.K, a.w, [&](int i) { return s_src[i] + e; });
            if (divide) {
                acc.x = __fdiv_rn(acc.x, a.divisor);
                acc.y = __fdiv_rn(acc.y, a.divisor);
                acc.z = __fdiv_rn(acc.z, a.divisor);
                acc.w = __fdiv_rn(acc.w, a.divisor);
            }
            *reinterpret_cast<float4*>(dst + e) = acc;
        }
        for (int64_t e = start + (nvec << 2) + threadIdx.x; e < start + len; e += kMultiThreads) {
            float acc;
            int first;
            if (accumulate) { acc = dst[e]; first = 0; }
            else { acc = __fmul_rn(s_src[0][e], a.w[0]); first = 1; }
            for (int i = first; i < a.K; ++i) acc = __fadd_rn(acc, __fmul_rn(s_src[i][e], a.w[i]));
            if (divide) acc = __fdiv_rn(acc, a.divisor);
            dst[e] = acc;
        }
    }
}

// ---------------------------------------------------------------- int64 counters
struct FedAvgI64Args {
    const int64_t* src[FMLP_MAX_CLIENTS];
    double w[FMLP_MAX_CLIENTS];
    float* out;
    int64_t J;
    double divisor;
    int K;
    int weights_integral;
    int flags;
};

// The reference keeps int64 through the multiply/adds when the weights are Python ints and
// only the final true division promotes to float32: (float)acc / (float)divisor.  With float
// weights `int64 * float` promotes immediately, so the fold is the fp32 one.
// ACCUMULATE is only meaningful with float weights (the int64 partial cannot live in `out`);
// the host wrapper never chains integral launches.
__device__ __forceinline__ float fold_i64(const FedAvgI64Args& a, int64_t e, float prev,
                                          const int64_t* const* src) {
    const bool accumulate = (a.flags & FMLP_FEDAVG_ACCUMULATE) != 0;
    const bool divide = (a.flags & FMLP_FEDAVG_DIVIDE) != 0;
    float r;
    if (a.weights_integral) {
        long long acc = 0;
        for (int i = 0; i < a.K; ++i) acc += src[i][e] * (long long)a.w[i];
        r = (float)acc;
        if (divide) r = __fdiv_rn(r, (float)a.divisor);
    } else {
        float acc;
        int first;
        if (accumulate) { acc = prev; first = 0; }
        else { acc = __fmul_rn((float)src[0][e], (float)a.w[0]); first = 1; }
        for (int i = first; i < a.K; ++i) acc = __fadd_rn(acc, __fmul_rn((float)src[i][e], (float)a.w[i]));
        r = acc;
        if (divide) r = __fdiv_rn(r, (float)a.divisor);
    }
    return r;
}

__global__ void __launch_bounds__(256)
fedavg_flat_i64_kernel(const __grid_constant__ FedAvgI64Args a) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = tid; e < a.J; e += nthreads) {
        const float prev = (a.flags & FMLP_FEDAVG_ACCUMULATE) ? a.out[e] : 0.f;
        a.out[e] = fold_i64(a, e, prev, a.src);
    }
}

struct FedAvgMultiI64Args {
    const int64_t* const* src_table;  // [T*K]
    float* const* dst_table;          // [T]
    const int32_t* elem_tensor;       // [J]
    const int64_t* elem_index;        // [J]
    int64_t J;
    FedAvgI64Args f;  // w, divisor, K, flags (src/out unused)
};

__global__ void __launch_bounds__(256)
fedavg_multi_i64_kernel(const __grid_constant__ FedAvgMultiI64Args a) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const bool accumulate = (a.f.flags & FMLP_FEDAVG_ACCUMULATE) != 0;
    const bool divide = (a.f.flags & FMLP_FEDAVG_DIVIDE) != 0;
    for (int64_t j = tid; j < a.J; j += nthreads) {
        const int t = a.elem_tensor[j];
        const int64_t e = a.elem_index[j];
        const int64_t* const* src = a.src_table + (int64_t)t * a.f.K;
        float* dst = a.dst_table[t];
        float r;
        if (a.f.weights_integral) {
            long long acc = 0;
            for (int i = 0; i < a.f.K; ++i) acc += src[i][e] * (long long)a.f.w[i];
            r = (float)acc;
        } else {
            float acc;
            int first;
            if (accumulate) { acc = dst[e]; first = 0; }
            else { acc = __fmul_rn((float)src[0][e], (float)a.f.w[0]); first = 1; }
            for (int i = first; i < a.f.K; ++i)
                acc = __fadd_rn(acc, __fmul_rn((float)src[i][e], (float)a.f.w[i]));
            r = acc;
        }
        if (divide) r = __fdiv_rn(r, (float)a.f.divisor);
        dst[e] = r;
    }
}

// ---------------------------------------------------------------- prototype aggregation
struct ProtoAvgArgs {
    const float* protos;  // [K][2C][D]
    float* out;           // [2C][D]
    float w[FMLP_MAX_CLIENTS];
    float class_div[FMLP_MAX_CLASSES];  // fp32(sum of the ORIGINAL weights over act(c))
    uint64_t class_clients[FMLP_MAX_CLASSES];
    int K, C, D, rpc;  // rpc = rows per class: 2 for FedAvg_proto, 1 for FedAvg_rela
};

// grid.x = 2C rows; FedAvg.py:79-86: acc = P_i*w_i + acc over act(c) in list (= ascending
// client) order starting from zeros, then / np.sum(weights[act]).  The divisor is formed by
// the host (entry point below) from the original weights in double and rounded once to fp32,
// which is what torch does with a numpy scalar divisor.
__device__ __forceinline__ void proto_avg_row(const ProtoAvgArgs& a, int row) {
    const int c = row / a.rpc;
    const uint64_t members = a.class_clients[c];
    const float divisor = a.class_div[c];
    const int64_t row_off = (int64_t)row * a.D;
    const int64_t client_stride = (int64_t)a.rpc * a.C * a.D;
    for (int d = threadIdx.x; d < a.D; d += blockDim.x) {
        float acc = 0.f;
        // the members' values are loaded eight at a time before they are folded (in client order): the fold is
        // a dependent chain, the loads need not be (r02 ncu: 7.2 us for 10 x 1024 outputs with one load per step)
        for (int i0 = 0; i0 < a.K; i0 += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u;
                v[u] = (i < a.K && ((members >> i) & 1ull)) ? a.protos[i * client_stride + row_off + d] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u;
                if (i < a.K && ((members >> i) & 1ull)) acc = __fadd_rn(__fmul_rn(v[u], a.w[i]), acc);
            }
        }
        a.out[row_off + d] = __fdiv_rn(acc, divisor);
    }
}

__global__ void __launch_bounds__(256) proto_avg_kernel(const __grid_constant__ ProtoAvgArgs a) { proto_avg_row(a, blockIdx.x); }

// ---- all the small tails of the single-GPU aggregation in ONE launch (main.py:218-234) --------
// CTAs [0, 2C): FedAvg_proto rows (bit-exact, as above).  CTA 2C: FedAvg_tao (utils/FedAvg.py:51-70, IEEE
// double in the reference's order, t_k[c] = count / N_k as at local_training.py:1000,1249) and the int64
// BatchNorm counters (FedAvg.py:9-13: weighted sum, then the float32 divide).  Round 1/early round 2 ran
// these as three dependent launches (proto_avg -> tail pack -> finalize, 16 us of launch latency at the
// end of the aggregation chain).
struct AggTailsArgs {
    ProtoAvgArgs p;
    double w[FMLP_MAX_CLIENTS];
    int64_t rows[FMLP_MAX_CLIENTS];
    uint32_t neg[FMLP_MAX_CLIENTS];
    const int64_t* counters[FMLP_MAX_CLIENTS];
    const int32_t* tcnt;   // [K][C]
    double* tao_out;       // [C]
    float* counters_out;   // [J]
    double total;
    int J;
};

__global__ void __launch_bounds__(256) agg_tails_local_kernel(const __grid_constant__ AggTailsArgs a) {
    const int n_rows = a.p.rpc * a.p.C;
    if ((int)blockIdx.x < n_rows) { proto_avg_row(a.p, blockIdx.x); return; }
    const int K = a.p.K, C = a.p.C;
    if (a.tao_out && (int)threadIdx.x < C) {
        const int c = threadIdx.x;
        double acc = 0.0, wsum = 0.0;
        bool any = false;
        for (int k0 = 0; k0 < K; k0 += 8) {
            int cnt[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) cnt[u] = (a.tcnt && k0 + u < K) ? a.tcnt[(k0 + u) * C + c] : 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int k = k0 + u;
                if (k < K && ((a.neg[k] >> c) & 1u)) {
                    const double t = __ddiv_rn((double)cnt[u], (double)a.rows[k]);
                    acc = __dadd_rn(acc, __dmul_rn(t, a.w[k]));
                    wsum = __dadd_rn(wsum, a.w[k]);
                    any = true;
                }
            }
        }
        a.tao_out[c] = any ? __ddiv_rn(acc, wsum) : 1.0;
    }
    if (a.counters_out) {
        for (int j = threadIdx.x; j < a.J; j += blockDim.x) {
            double s = 0.0;
            for (int k0 = 0; k0 < K; k0 += 8) {
                long long v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = (k0 + u < K) ? a.counters[k0 + u][j] : 0ll;
#pragma unroll
                for (int u = 0; u < 8; ++u) if (k0 + u < K) s += (double)v[u] * a.w[k0 + u];
            }
            a.counters_out[j] = (float)(int64_t)s / (float)a.total;
        }
    }
}

// ---------------------------------------------------------------- tao aggregation (float64)
struct TaoAvgArgs {
    const double* t;   // [K][C]
    double* out;       // [C]
    double w[FMLP_MAX_CLIENTS];
    uint64_t class_clients[FMLP_MAX_CLASSES];
    int K, C, use_lists;
};

// utils/FedAvg.py:51-70 in the reference's operation order, in IEEE double without contraction:
//   lists given : t_avg[c] = (sum_{i in list(c), ascending} t_i[c]*w_i) / (sum w_i), 1.0 for an empty list
//   no lists    : t_avg    = (sum_i t_i*w_i) / sum(w)   (the divisor is passed in w[K])
__global__ void tao_avg_kernel(const __grid_constant__ TaoAvgArgs a, double total_w) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.C) return;
    double acc = 0.0, wsum = 0.0;
    bool any = false;
    for (int i = 0; i < a.K; ++i) {
        if (a.use_lists && !((a.class_clients[c] >> i) & 1ull)) continue;
        acc = __dadd_rn(acc, __dmul_rn(a.t[(int64_t)i * a.C + c], a.w[i]));
        wsum = __dadd_rn(wsum, a.w[i]);
        any = true;
    }
    if (a.use_lists) a.out[c] = any ? __ddiv_rn(acc, wsum) : 1.0;
    else a.out[c] = __ddiv_rn(acc, total_w);
}

// ---------------------------------------------------------------- model_dist
// utils/FedNoRo.py:106-115 (and utils/FedAvg.py:42-49): sum over the float tensors, in key order,
// of ||w1[k] - w2[k]||_2.  Deterministic: per-chunk sums of squares (fixed tree inside the CTA),
// per-tensor chunk sums in chunk order, then ONE thread adds the per-tensor norms in key order
// exactly like the reference's `dist_total += dist` loop.
struct ModelDistArgs {
    const float* const* a_table;    // [T]
    const float* const* b_table;    // [T]
    const int64_t* numel;           // [T]
    const int32_t* chunk_tensor;    // [n_chunks]
    const int64_t* chunk_start;     // [n_chunks]
    const int64_t* tensor_chunk0;   // [T+1] first chunk of every tensor
    float* partial;                 // ws [n_chunks + T]
    float* out;                     // [1]
    int64_t n_chunks;
    int T;
};

__global__ void __launch_bounds__(256) model_dist_partial_kernel(const __grid_constant__ ModelDistArgs a) {
    __shared__ float s_red[8];
    for (int64_t c = blockIdx.x; c < a.n_chunks; c += gridDim.x) {
        const int t = a.chunk_tensor[c];
        const int64_t start = a.chunk_start[c];
        const int64_t len = min((int64_t)FMLP_FEDAVG_CHUNK, a.numel[t] - start);
        const float* x = a.a_table[t] + start;
        const float* y = a.b_table[t] + start;
        float ss = 0.f;
        for (int64_t e = threadIdx.x; e < len; e += 256) { const float d = __fsub_rn(x[e], y[e]); ss = fmaf(d, d, ss); }
        ss = warp_sum(ss);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += s_red[w];
            a.partial[c] = tot;
        }
    }
}

__global__ void __launch_bounds__(256) model_dist_finalize_kernel(const __grid_constant__ ModelDistArgs a) {
    float* norms = a.partial + a.n_chunks;
    for (int t = threadIdx.x; t < a.T; t += 256) {
        float ss = 0.f;
        for (int64_t c = a.tensor_chunk0[t]; c < a.tensor_chunk0[t + 1]; ++c) ss += a.partial[c];
        norms[t] = sqrtf(ss);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float total = 0.f;
        for (int t = 0; t < a.T; ++t) total = __fadd_rn(total, norms[t]);
        a.out[0] = total;
    }
}

}  // namespace fmlp

using namespace fmlp;

static int check_k(int K) { return (K >= 1 && K <= FMLP_MAX_CLIENTS) ? FMLP_OK : FMLP_ERR_BAD_ARG; }

extern "C" int fmlp_fedavg_flat_f32(const float* const* srcs, const float* weights, int K,
                                    int64_t P, float divisor, int flags, float* out,
                                    fmlp_stream_t stream) {
    if (!srcs || !weights || !out || P < 0 || check_k(K) != FMLP_OK) return FMLP_ERR_BAD_ARG;
    if (P == 0) return FMLP_OK;
    FedAvgArgs a;
    bool vec = aligned16(out);
    for (int i = 0; i < K; ++i) {
        if (!srcs[i]) return FMLP_ERR_BAD_ARG;
        a.src[i] = srcs[i];
        a.w[i] = weights[i];
        vec = vec && aligned16(srcs[i]);
    }
    for (int i = K; i < FMLP_MAX_CLIENTS; ++i) { a.src[i] = nullptr; a.w[i] = 0.f; }
    a.out = out; a.P = P; a.divisor = divisor; a.K = K; a.flags = flags;
    const int sms = sm_count();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    const int64_t work = vec ? ((P + 3) >> 2) : P;
    int64_t blocks = (work + kFedAvgThreads - 1) / kFedAvgThreads;
    const int64_t cap = (int64_t)sms * 4;  // persistent: 4 resident CTAs per SM
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (vec) fedavg_flat_kernel<true><<<(unsigned)blocks, kFedAvgThreads, 0, st>>>(a);
    else fedavg_flat_kernel<false><<<(unsigned)blocks, kFedAvgThreads, 0, st>>>(a);
    return launch_status();
}

extern "C" int fmlp_fedavg_flat_i64(const int64_t* const* srcs, const double* weights, int K,
                                    int64_t J, double divisor, int weights_integral, int flags,
                                    float* out, fmlp_stream_t stream) {
    if (!srcs || !weights || !out || J < 0 || check_k(K) != FMLP_OK) return FMLP_ERR_BAD_ARG;
    if (weights_integral && (flags & FMLP_FEDAVG_ACCUMULATE)) return FMLP_ERR_UNSUPPORTED;
    if (J == 0) return FMLP_OK;
    FedAvgI64Args a;
    for (int i = 0; i < K; ++i) {
        if (!srcs[i]) return FMLP_ERR_BAD_ARG;
        a.src[i] = srcs[i];
        a.w[i] = weights[i];
    }
    for (int i = K; i < FMLP_MAX_CLIENTS; ++i) { a.src[i] = nullptr; a.w[i] = 0.0; }
    a.out = out; a.J = J; a.divisor = divisor; a.K = K; a.weights_integral = weights_integral;
    a.flags = flags;
    int64_t blocks = (J + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    fedavg_flat_i64_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    return launch_status();
}

extern "C" int fmlp_fedavg_multi_f32(const float* const* src_table_dev, float* const* dst_table_dev,
                                     const int64_t* numel_dev, const int32_t* chunk_tensor_dev,
                                     const int64_t* chunk_start_dev, int64_t n_chunks, int T,
                                     const float* weights, int K, float divisor, int flags,
                                     fmlp_stream_t stream) {
    if (!src_table_dev || !dst_table_dev || !numel_dev || !chunk_tensor_dev || !chunk_start_dev ||
        !weights || T < 0 || n_chunks < 0 || check_k(K) != FMLP_OK)
        return FMLP_ERR_BAD_ARG;
    if (n_chunks == 0 || T == 0) return FMLP_OK;
    FedAvgMultiArgs a;
    a.src_table = src_table_dev; a.dst_table = dst_table_dev; a.numel = numel_dev;
    a.chunk_tensor = chunk_tensor_dev; a.chunk_start = chunk_start_dev; a.n_chunks = n_chunks;
    for (int i = 0; i < FMLP_MAX_CLIENTS; ++i) a.w[i] = i < K ? weights[i] : 0.f;
    a.divisor = divisor; a.K = K; a.flags = flags;
    const int sms = sm_count();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    int64_t blocks = n_chunks;
    const int64_t cap = (int64_t)sms * 4;
    if (blocks > cap) blocks = cap;
    fedavg_multi_kernel<<<(unsigned)blocks, kMultiThreads, 0, (cudaStream_t)stream>>>(a);
    return launch_status();
}

extern "C" int fmlp_fedavg_multi_i64(const int64_t* const* src_table_dev, float* const* dst_table_dev,
                                     const int32_t* elem_tensor_dev, const int64_t* elem_index_dev,
                                     int64_t J, int T, const double* weights, int K, double divisor,
                                     int weights_integral, int flags, fmlp_stream_t stream) {
    if (!src_table_dev || !dst_table_dev || !elem_tensor_dev || !elem_index_dev || !weights ||
        T < 0 || J < 0 || check_k(K) != FMLP_OK)
        return FMLP_ERR_BAD_ARG;
    if (weights_integral && (flags & FMLP_FEDAVG_ACCUMULATE)) return FMLP_ERR_UNSUPPORTED;
    if (J == 0) return FMLP_OK;
    FedAvgMultiI64Args a;
    a.src_table = src_table_dev; a.dst_table = dst_table_dev; a.elem_tensor = elem_tensor_dev;
    a.elem_index = elem_index_dev; a.J = J;
    for (int i = 0; i < FMLP_MAX_CLIENTS; ++i) { a.f.src[i] = nullptr; a.f.w[i] = i < K ? weights[i] : 0.0; }
    a.f.out = nullptr; a.f.J = J; a.f.divisor = divisor; a.f.K = K;
    a.f.weights_integral = weights_integral; a.f.flags = flags;
    int64_t blocks = (J + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    fedavg_multi_i64_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    return launch_status();
}

extern "C" int fmlp_proto_avg_f32(const float* protos, int K, int C, int D, int rows_per_class,
                                  const double* weights, const uint64_t* class_clients, float* out,
                                  fmlp_stream_t stream) {
    if (!protos || !weights || !class_clients || !out || C < 1 || C > FMLP_MAX_CLASSES || D < 1 ||
        (rows_per_class != 1 && rows_per_class != 2) || check_k(K) != FMLP_OK)
        return FMLP_ERR_BAD_ARG;
    ProtoAvgArgs a;
    a.protos = protos; a.out = out; a.K = K; a.C = C; a.D = D; a.rpc = rows_per_class;
    for (int i = 0; i < FMLP_MAX_CLIENTS; ++i) a.w[i] = i < K ? (float)weights[i] : 0.f;
    for (int c = 0; c < FMLP_MAX_CLASSES; ++c) {
        a.class_clients[c] = c < C ? class_clients[c] : 0ull;
        double wsum = 0.0;
        for (int i = 0; i < K; ++i)
            if (c < C && ((class_clients[c] >> i) & 1ull)) wsum += weights[i];
        a.class_div[c] = (float)wsum;
    }
    proto_avg_kernel<<<rows_per_class * C, 256, 0, (cudaStream_t)stream>>>(a);
    return launch_status();
}

extern "C" int fmlp_agg_tails_local_f32(const float* protos, int K, int C, int D, const double* weights,
                                        const uint64_t* class_clients, float* proto_out, const int32_t* tcnt,
                                        const int64_t* rows, const uint32_t* neg, const int64_t* const* counters, int J,
                                        double total_weight, double* tao_out, float* counters_out, fmlp_stream_t stream) {
    if (!protos || !weights || !class_clients || !proto_out || C < 1 || C > FMLP_MAX_CLASSES || D < 1 || J < 0 ||
        check_k(K) != FMLP_OK || !(total_weight > 0.0))
        return FMLP_ERR_BAD_ARG;
    if (tao_out && (!rows || !neg)) return FMLP_ERR_BAD_ARG;
    if (J > 0 && (!counters || !counters_out)) return FMLP_ERR_BAD_ARG;
    AggTailsArgs a;
    a.p.protos = protos; a.p.out = proto_out; a.p.K = K; a.p.C = C; a.p.D = D; a.p.rpc = 2;
    for (int i = 0; i < FMLP_MAX_CLIENTS; ++i) {
        a.p.w[i] = i < K ? (float)weights[i] : 0.f;
        a.w[i] = i < K ? weights[i] : 0.0;
        a.rows[i] = (i < K && rows && rows[i] > 0) ? rows[i] : 1;
        a.neg[i] = (i < K && neg) ? neg[i] : 0u;
        a.counters[i] = (i < K && J > 0) ? counters[i] : nullptr;
        if (i < K && J > 0 && !counters[i]) return FMLP_ERR_BAD_ARG;
    }
    for (int c = 0; c < FMLP_MAX_CLASSES; ++c) {
        a.p.class_clients[c] = c < C ? class_clients[c] : 0ull;
        double wsum = 0.0;
        for (int i = 0; i < K; ++i)
            if (c < C && ((class_clients[c] >> i) & 1ull)) wsum += weights[i];
        a.p.class_div[c] = (float)wsum;
    }
    a.tcnt = tcnt; a.tao_out = tao_out; a.counters_out = J > 0 ? counters_out : nullptr; a.total = total_weight; a.J = J;
    const int tail_cta = (tao_out || J > 0) ? 1 : 0;
    agg_tails_local_kernel<<<2 * C + tail_cta, 256, 0, (cudaStream_t)stream>>>(a);
    return launch_status();
}

extern "C" size_t fmlp_model_dist_ws_bytes(int64_t n_chunks, int T) {
    return (size_t)((n_chunks > 0 ? n_chunks : 0) + (T > 0 ? T : 0) + 1) * sizeof(float);
}

extern "C" int fmlp_model_dist_f32(const float* const* a_table_dev, const float* const* b_table_dev,
                                   const int64_t* numel_dev, const int32_t* chunk_tensor_dev,
                                   const int64_t* chunk_start_dev, const int64_t* tensor_chunk0_dev,
                                   int64_t n_chunks, int T, float* out, void* ws, size_t ws_bytes,
                                   fmlp_stream_t stream) {
    if (!a_table_dev || !b_table_dev || !numel_dev || !chunk_tensor_dev || !chunk_start_dev || !tensor_chunk0_dev ||
        !out || !ws || n_chunks < 0 || T < 0)
        return FMLP_ERR_BAD_ARG;
    if (ws_bytes < fmlp_model_dist_ws_bytes(n_chunks, T)) return FMLP_ERR_WORKSPACE;
    ModelDistArgs a;
    a.a_table = a_table_dev; a.b_table = b_table_dev; a.numel = numel_dev; a.chunk_tensor = chunk_tensor_dev;
    a.chunk_start = chunk_start_dev; a.tensor_chunk0 = tensor_chunk0_dev; a.partial = (float*)ws; a.out = out;
    a.n_chunks = n_chunks; a.T = T;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_chunks > 0) {
        const int sms = sm_count();
        if (sms <= 0) return (int)cudaErrorInvalidDevice;
        int64_t blocks = n_chunks < (int64_t)sms * 8 ? n_chunks : (int64_t)sms * 8;
        model_dist_partial_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
        int rc = launch_status();
        if (rc != FMLP_OK) return rc;
    }
    model_dist_finalize_kernel<<<1, 256, 0, st>>>(a);
    return launch_status();
}

extern "C" int fmlp_tao_avg_f64(const double* t_dev, int K, int C, const double* weights,
                                const uint64_t* class_clients, double total_weight, double* out_dev,
                                fmlp_stream_t stream) {
    if (!t_dev || !weights || !out_dev || C < 1 || C > FMLP_MAX_CLASSES || check_k(K) != FMLP_OK) return FMLP_ERR_BAD_ARG;
    TaoAvgArgs a;
    a.t = t_dev; a.out = out_dev; a.K = K; a.C = C; a.use_lists = class_clients ? 1 : 0;
    for (int i = 0; i < FMLP_MAX_CLIENTS; ++i) a.w[i] = i < K ? weights[i] : 0.0;
    for (int c = 0; c < FMLP_MAX_CLASSES; ++c) a.class_clients[c] = (class_clients && c < C) ? class_clients[c] : 0ull;
    tao_avg_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a, total_weight);
    return launch_status();
}
