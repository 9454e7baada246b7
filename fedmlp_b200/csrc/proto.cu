// proto.cu — K2: class-prototype construction (label-masked segmented feature means) and the
// per-class confident-prediction counts behind the difficulty statistic `t`.
//
// Replaces utils/local_training.py:973-1000 (end of stage 1) and :1208-1249 (every stage-2
// round).  The reference walks the data in batches of 128 rows and, per active class, does
// where/gather/sum/.cpu() twice plus one .item() per missing class (3+ host syncs per batch).
// Here the [N, D] features are streamed once: a CTA owns the full width of a chunk of rows
// (thread <-> float4 column, so every row is one coalesced request per warp), keeps the
// label-0 / label-1 running sums of up to 4 active classes in registers, and emits one partial
// per chunk; a second tiny kernel adds the partials of a segment in chunk order (deterministic,
// no float atomics) and applies the guarded divide.  Algorithmic bytes: 4*N*D + 8*N*C + 8*C*D.
#include "common.cuh"

namespace fmlp {

constexpr int kProtoMaxActive = 4;  // active classes accumulated per pass (annotation_num is 1)
constexpr int kProtoUnroll = 8;     // rows in flight per thread

// Chunk height: large enough that the partial traffic (2*NA*D*4 B per chunk, written and read
// once) stays ~3 % of the feature bytes, small enough that >= 2 CTAs per SM exist.
static inline int proto_chunk_rows(int64_t n_total) {
    if (n_total >= 37888) return 128;  // 256 rows * 148 SMs
    if (n_total >= 14208) return 64;
    return 32;
}

struct ProtoArgs {
    const float* feat;
    const float* labels;
    const float* logits;
    float* partial;   // [n_items][NA][2][D]
    int32_t* cnt;     // [S][2C]
    int32_t* tcnt;    // [S][C]
    int64_t ld_feat;
    int64_t item_base[FMLP_MAX_SEGMENTS + 1];  // prefix of chunk counts per segment
    SegTable seg;     // mask_a = active classes (restricted to this pass), mask_b = tcount classes
    float L, U;
    int D, C, chunk_rows, logits_are_probs, count_labels;
};

template <int NA>
__global__ void __launch_bounds__(512) proto_accum_kernel(const __grid_constant__ ProtoArgs a) {
    __shared__ int s_t[FMLP_MAX_CLASSES];
    const int col = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
    const bool col_ok = col < a.D;
    const int64_t n_items = a.item_base[a.seg.S];

    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        // segment of this item (item_base is a prefix array like seg.rows)
        const int s = find_segment(a.item_base, a.seg.S, item);
        const int64_t r_begin = a.seg.rows[s] + (item - a.item_base[s]) * a.chunk_rows;
        const int64_t r_end = min(r_begin + a.chunk_rows, a.seg.rows[s + 1]);
        const uint32_t active = a.seg.mask_a[s];
        int cls[NA];
        {
            uint32_t m = active;
#pragma unroll
            for (int i = 0; i < NA; ++i) {
                if (m) { cls[i] = __ffs(m) - 1; m &= m - 1; } else cls[i] = -1;
            }
        }
        float4 acc[NA][2];
        int n_lab[NA][2];
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            acc[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
            acc[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            n_lab[i][0] = 0; n_lab[i][1] = 0;
        }

        for (int64_t r0 = r_begin; r0 < r_end; r0 += kProtoUnroll) {
            float4 f[kProtoUnroll];
            float y[kProtoUnroll][NA];
#pragma unroll
            for (int u = 0; u < kProtoUnroll; ++u) {
                const int64_t row = r0 + u;
                const bool ok = row < r_end;
                f[u] = (ok && col_ok) ? ld_stream_f4(a.feat + row * a.ld_feat + col)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < NA; ++i)
                    y[u][i] = (ok && cls[i] >= 0) ? __ldg(a.labels + row * a.C + cls[i]) : -1.f;
            }
#pragma unroll
            for (int u = 0; u < kProtoUnroll; ++u) {
#pragma unroll
                for (int i = 0; i < NA; ++i) {
                    // labels == 0 / labels == 1 exactly as torch.where(labels[:, cls] == k) (:985-986)
                    if (y[u][i] == 0.f) {
                        acc[i][0].x += f[u].x; acc[i][0].y += f[u].y; acc[i][0].z += f[u].z; acc[i][0].w += f[u].w;
                        n_lab[i][0]++;
                    } else if (y[u][i] == 1.f) {
                        acc[i][1].x += f[u].x; acc[i][1].y += f[u].y; acc[i][1].z += f[u].z; acc[i][1].w += f[u].w;
                        n_lab[i][1]++;
                    }
                }
            }
        }
        if (col_ok) {
#pragma unroll
            for (int i = 0; i < NA; ++i) {
                float* dst = a.partial + ((item * NA + i) * 2) * (int64_t)a.D + col;
                *reinterpret_cast<float4*>(dst) = acc[i][0];
                *reinterpret_cast<float4*>(dst + a.D) = acc[i][1];
            }
        }
        if (blockIdx.y == 0) {
            if (a.count_labels && threadIdx.x == 0) {
#pragma unroll
                for (int i = 0; i < NA; ++i)
                    if (cls[i] >= 0) {
                        if (n_lab[i][0]) atomicAdd(a.cnt + (int64_t)s * 2 * a.C + 2 * cls[i], n_lab[i][0]);
                        if (n_lab[i][1]) atomicAdd(a.cnt + (int64_t)s * 2 * a.C + 2 * cls[i] + 1, n_lab[i][1]);
                    }
            }
            // confident-prediction counts (:995-996, :1239): #{p < L or p > U}
            const uint32_t tmask = a.seg.mask_b[s];
            if (a.logits != nullptr && a.tcnt != nullptr && tmask != 0u) {
                __syncthreads();
                if (threadIdx.x < FMLP_MAX_CLASSES) s_t[threadIdx.x] = 0;
                __syncthreads();
                const int64_t n_el = (r_end - r_begin) * a.C;
                const float* z = a.logits + r_begin * a.C;
                for (int64_t e = threadIdx.x; e < n_el; e += blockDim.x) {
                    const int c = (int)(e % a.C);
                    if ((tmask >> c) & 1u) {
                        const float v = z[e];
                        const float p = a.logits_are_probs ? v : sigmoid_ref(v);
                        if (p < a.L || p > a.U) atomicAdd(&s_t[c], 1);
                    }
                }
                __syncthreads();
                if (threadIdx.x < a.C && s_t[threadIdx.x] != 0)
                    atomicAdd(a.tcnt + (int64_t)s * a.C + threadIdx.x, s_t[threadIdx.x]);
            }
        }
    }
}

struct ProtoFinArgs {
    const float* partial;
    const int32_t* cnt;  // [S][2C]
    float* proto;        // [S][2C][D]
    int64_t item_base[FMLP_MAX_SEGMENTS + 1];
    SegTable seg;        // mask_a = active classes of THIS pass; mask_b = all active classes
    int D, C, NA, guard_empty, first_pass;
};

// grid = (S * 2C, ceil(D / 256)).  Row 2c+y of segment s: sum the chunk partials in chunk
// order, divide by the row count (tensor / python int -> fp32 divide, :997-999,1241-1248).
__global__ void __launch_bounds__(256) proto_finalize_kernel(const __grid_constant__ ProtoFinArgs a) {
    const int s = blockIdx.x / (2 * a.C);
    const int row = blockIdx.x - s * 2 * a.C;
    const int c = row >> 1, y = row & 1;
    const int d = blockIdx.y * blockDim.x + threadIdx.x;
    if (d >= a.D) return;
    float* out = a.proto + ((int64_t)s * 2 * a.C + row) * a.D + d;
    const uint32_t pass_active = a.seg.mask_a[s];
    if (!((pass_active >> c) & 1u)) {
        // rows of classes that are not active on this client stay zero (proto = torch.zeros, :973)
        if (a.first_pass && !((a.seg.mask_b[s] >> c) & 1u)) *out = 0.f;
        return;
    }
    const int slot = __popc(pass_active & ((1u << c) - 1u));  // position among this pass's classes
    const int64_t i0 = a.item_base[s], i1 = a.item_base[s + 1];
    float acc = 0.f;
    for (int64_t it = i0; it < i1; ++it)
        acc += a.partial[((it * a.NA + slot) * 2 + y) * (int64_t)a.D + d];
    const int n = a.cnt[(int64_t)s * 2 * a.C + row];
    if (n == 0 && a.guard_empty) *out = acc;  // == 0
    else *out = __fdiv_rn(acc, (float)n);
}

__global__ void zero_i32_kernel(int32_t* p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0;
}

}  // namespace fmlp

using namespace fmlp;

static int64_t proto_items(const int64_t* seg_rows, int S, int chunk, int64_t* item_base) {
    int64_t n = 0;
    for (int s = 0; s < S; ++s) {
        if (item_base) item_base[s] = n;
        n += (seg_rows[s + 1] - seg_rows[s] + chunk - 1) / chunk;
    }
    if (item_base) for (int s = S; s <= FMLP_MAX_SEGMENTS; ++s) item_base[s] = n;
    return n;
}

extern "C" size_t fmlp_proto_ws_bytes(int64_t n_total, int D, int C, int S) {
    (void)C;
    if (n_total < 0 || D < 1 || S < 1) return 0;
    const int chunk = proto_chunk_rows(n_total);
    const int64_t items = n_total / chunk + S;  // upper bound on sum of ceil(N_s / chunk)
    return (size_t)items * kProtoMaxActive * 2 * (size_t)D * sizeof(float);
}

extern "C" int fmlp_proto_build_f32(const float* feat, int64_t ld_feat, int D, const float* labels,
                                    const float* logits, int logits_are_probs, int C, int S,
                                    const int64_t* seg_rows, const uint32_t* seg_active,
                                    const uint32_t* seg_tcount, float L, float U, int guard_empty,
                                    float* proto, int32_t* cnt, int32_t* tcnt, void* ws,
                                    size_t ws_bytes, fmlp_stream_t stream) {
    if (!feat || !labels || !seg_active || !proto || !cnt || !ws || C < 1 || C > FMLP_MAX_CLASSES || D < 4)
        return FMLP_ERR_BAD_ARG;
    if (logits && !tcnt) return FMLP_ERR_BAD_ARG;
    if ((D & 3) || (ld_feat & 3) || ld_feat < D || !aligned16(feat) || !aligned16(ws)) return FMLP_ERR_UNSUPPORTED;
    const uint32_t cmask = C < 32 ? ((1u << C) - 1u) : 0xffffffffu;
    cudaStream_t st = (cudaStream_t)stream;

    ProtoArgs a;
    uint32_t act_all[FMLP_MAX_SEGMENTS], act_left[FMLP_MAX_SEGMENTS], tc[FMLP_MAX_SEGMENTS];
    if (S < 1 || S > FMLP_MAX_SEGMENTS || !seg_rows) return FMLP_ERR_BAD_ARG;
    for (int s = 0; s < S; ++s) {
        act_all[s] = seg_active[s] & cmask;
        act_left[s] = act_all[s];
        tc[s] = (seg_tcount ? seg_tcount[s] : 0u) & cmask;
    }
    const int64_t n_total = seg_rows[S];
    const int chunk = proto_chunk_rows(n_total);
    const int64_t n_items = proto_items(seg_rows, S, chunk, a.item_base);
    if (ws_bytes < (size_t)n_items * kProtoMaxActive * 2 * (size_t)D * sizeof(float)) return FMLP_ERR_WORKSPACE;

    // counters are accumulated with integer atomics -> clear them first
    const int64_t n_cnt = (int64_t)S * 2 * C;
    zero_i32_kernel<<<1, 256, 0, st>>>(cnt, n_cnt);
    if (tcnt) zero_i32_kernel<<<1, 256, 0, st>>>(tcnt, (int64_t)S * C);

    a.feat = feat; a.labels = labels; a.logits = logits; a.partial = (float*)ws; a.cnt = cnt; a.tcnt = tcnt;
    a.ld_feat = ld_feat; a.L = L; a.U = U; a.D = D; a.C = C; a.chunk_rows = chunk;
    a.logits_are_probs = logits_are_probs; a.count_labels = 1;

    const int sms = sm_count();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    const int nvec = D / 4;
    int threads = ((nvec + 31) / 32) * 32;
    if (threads > 512) threads = 512;
    const int gy = (nvec + threads - 1) / threads;

    bool first = true;
    for (;;) {
        // take up to kProtoMaxActive active classes per segment for this pass
        uint32_t pass[FMLP_MAX_SEGMENTS];
        int na_max = 0;
        for (int s = 0; s < S; ++s) {
            uint32_t m = act_left[s], take = 0;
            int n = 0;
            while (m && n < kProtoMaxActive) { uint32_t b = m & (~m + 1u); take |= b; m ^= b; ++n; }
            pass[s] = take;
            act_left[s] &= ~take;
            if (n > na_max) na_max = n;
        }
        if (na_max == 0 && !first) break;
        int rc = fill_seg_table(a.seg, S, seg_rows, pass, first ? tc : nullptr);
        if (rc != FMLP_OK) return rc;
        const int NA = na_max <= 1 ? 1 : (na_max == 2 ? 2 : 4);
        if (n_items > 0) {
            int64_t gx = n_items;
            const int64_t cap = (int64_t)sms * 8;
            if (gx > cap) gx = cap;
            dim3 grid((unsigned)gx, (unsigned)gy);
            if (NA == 1) proto_accum_kernel<1><<<grid, threads, 0, st>>>(a);
            else if (NA == 2) proto_accum_kernel<2><<<grid, threads, 0, st>>>(a);
            else proto_accum_kernel<4><<<grid, threads, 0, st>>>(a);
            rc = launch_status();
            if (rc != FMLP_OK) return rc;
        }
        ProtoFinArgs f;
        f.partial = (const float*)ws; f.cnt = cnt; f.proto = proto;
        for (int s = 0; s <= FMLP_MAX_SEGMENTS; ++s) f.item_base[s] = a.item_base[s];
        rc = fill_seg_table(f.seg, S, seg_rows, pass, act_all);
        if (rc != FMLP_OK) return rc;
        f.D = D; f.C = C; f.NA = NA; f.guard_empty = guard_empty; f.first_pass = first ? 1 : 0;
        dim3 fgrid((unsigned)(S * 2 * C), (unsigned)((D + 255) / 256));
        proto_finalize_kernel<<<fgrid, 256, 0, st>>>(f);
        rc = launch_status();
        if (rc != FMLP_OK) return rc;
        first = false;
        bool more = false;
        for (int s = 0; s < S; ++s) more = more || act_left[s] != 0;
        if (!more) break;
    }
    return FMLP_OK;
}
