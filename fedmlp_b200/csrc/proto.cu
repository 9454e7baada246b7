// proto.cu — K2: class-prototype construction (label-masked segmented feature means) and the
// per-class confident-prediction counts behind the difficulty statistic `t`.
//
// Replaces utils/local_training.py:973-1000 (end of stage 1) and :1208-1249 (every stage-2
// round).  The reference walks the data in batches of 128 rows and, per active class, does
// where/gather/sum/.cpu() twice plus one .item() per missing class (3+ host syncs per batch).
// Here the [N, D] features are streamed once.  Thread <-> float4 column, so a CTA reads whole
// rows (one coalesced 512-B request per warp per row) and keeps the label-0 / label-1 running
// sums of up to 4 active classes in registers.
//
// Work split (r01 ncu: the first version launched 1.46 waves of chunk-CTAs and lost 30 % to the
// tail): the grid is exactly the number of co-resident CTAs and CTA b owns the contiguous row
// range [b*rows_per_cta, (b+1)*rows_per_cta) -> one balanced wave.  Where a range crosses a
// segment (client) boundary the CTA flushes and starts a new partial; the partial of (CTA b,
// segment s) lives in slot b + s (injective, both grow along the rows).  Row loads are ping-pong
// buffered (4 + 4 rows in flight per thread) because B200 HBM needs ~100 KB in flight per SM.
// Label and t counts are per-slot partials too, so nothing is zeroed and nothing is atomic on
// global memory; a second small kernel adds the slots of each segment in CTA order
// (deterministic) and applies the guarded divide.  Algorithmic bytes: 4*N*D + 8*N*C + 8*C*D.
#include <cstdlib>

#include "common.cuh"

namespace fmlp {

constexpr int kProtoMaxActive = 4;  // active classes accumulated per pass (annotation_num is 1)
constexpr int kProtoUnroll = 4;     // rows per batch; two batches in flight (ping-pong)
constexpr int kProtoMaxCtasPerSm = 8;

struct ProtoArgs {
    const float* feat;
    const float* labels;
    const float* logits;
    float* partial;    // [slots][NA][2][D]
    int32_t* pcount;   // [slots][2*kProtoMaxActive + FMLP_MAX_CLASSES]
    int64_t ld_feat;
    int64_t rows_per_cta;
    SegTable seg;      // mask_a = active classes (restricted to this pass), mask_b = tcount classes
    float L, U;
    int D, C, logits_are_probs, do_t;
};
constexpr int kProtoCountStride = 2 * kProtoMaxActive + FMLP_MAX_CLASSES;

template <int NA>
__global__ void __launch_bounds__(512, NA == 1 ? 2 : 1) proto_accum_kernel(const __grid_constant__ ProtoArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the finalize kernel is a programmatic dependent launch
    __shared__ int s_t[FMLP_MAX_CLASSES];
    const int col = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
    const bool col_ok = col < a.D;
    const int S = a.seg.S;
    const int64_t n_total = a.seg.rows[S];
    const int64_t cta_begin = (int64_t)blockIdx.x * a.rows_per_cta;
    const int64_t cta_end = min(n_total, cta_begin + a.rows_per_cta);
    if (cta_begin >= n_total) return;
    int s = find_segment(a.seg.rows, S, cta_begin);
    int64_t row = cta_begin;

    while (row < cta_end) {
        while (s < S - 1 && a.seg.rows[s + 1] <= row) ++s;  // skip empty segments
        const int64_t r_end = min(a.seg.rows[s + 1], cta_end);
        const int64_t r_begin = row;
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
        float4 fa[kProtoUnroll], fb[kProtoUnroll];
        float ya[kProtoUnroll][NA], yb[kProtoUnroll][NA];

        auto load = [&](int64_t r0, float4* f, float (*y)[NA]) {
#pragma unroll
            for (int u = 0; u < kProtoUnroll; ++u) {
                const int64_t r = r0 + u;
                const bool ok = r < r_end;
                f[u] = (ok && col_ok) ? ld_stream_f4(a.feat + r * a.ld_feat + col) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < NA; ++i)
                    y[u][i] = (ok && cls[i] >= 0) ? __ldg(a.labels + r * a.C + cls[i]) : -1.f;
            }
        };
        auto consume = [&](const float4* f, const float (*y)[NA]) {
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
        };
        // Fast path (r02 ncu: the checked loop below issued 61 warp instructions per row, two thirds of them
        // address arithmetic and bounds predicates, and was issue-limited at 0.77 of the HBM roof): whole
        // double batches with running pointers and no bounds checks; the checked loop only mops up the
        // last < 2*kProtoUnroll rows of the range.
        int64_t r_done = r_begin;
        bool all_cls = col_ok;
#pragma unroll
        for (int i = 0; i < NA; ++i) all_cls = all_cls && cls[i] >= 0;
        const int64_t n_fast = ((r_end - r_begin) / (2 * kProtoUnroll)) * (2 * kProtoUnroll);
        if (all_cls && n_fast > 0) {
            const int ldi = (int)a.ld_feat, Ci = a.C;          // element strides (the entry point checks they fit)
            const float* fp = a.feat + r_begin * a.ld_feat + col;
            const float* lp = a.labels + r_begin * a.C;
            auto load_fast = [&](float4* f, float (*y)[NA]) {
#pragma unroll
                for (int u = 0; u < kProtoUnroll; ++u) {
                    f[u] = ld_stream_f4(fp);
#pragma unroll
                    for (int i = 0; i < NA; ++i) y[u][i] = __ldg(lp + cls[i]);
                    fp += ldi;
                    lp += Ci;
                }
            };
            load_fast(fa, ya);
            for (int64_t r = 0; r < n_fast; r += 2 * kProtoUnroll) {
                load_fast(fb, yb);
                consume(fa, ya);
                if (r + 2 * kProtoUnroll < n_fast) load_fast(fa, ya);
                consume(fb, yb);
            }
            r_done = r_begin + n_fast;
        }
        if (r_done < r_end) {
            load(r_done, fa, ya);
            for (int64_t r0 = r_done; r0 < r_end; r0 += 2 * kProtoUnroll) {
                load(r0 + kProtoUnroll, fb, yb);
                consume(fa, ya);
                load(r0 + 2 * kProtoUnroll, fa, ya);
                consume(fb, yb);
            }
        }

        const int64_t slot = (int64_t)blockIdx.x + s;
        if (col_ok) {
#pragma unroll
            for (int i = 0; i < NA; ++i) {
                float* dst = a.partial + ((slot * NA + i) * 2) * (int64_t)a.D + col;
                *reinterpret_cast<float4*>(dst) = acc[i][0];
                *reinterpret_cast<float4*>(dst + a.D) = acc[i][1];
            }
        }
        if (blockIdx.y == 0) {
            int32_t* pc = a.pcount + slot * kProtoCountStride;
            if (threadIdx.x == 0) {
#pragma unroll
                for (int i = 0; i < NA; ++i) { pc[2 * i] = n_lab[i][0]; pc[2 * i + 1] = n_lab[i][1]; }
            }
            // confident-prediction counts (:995-996, :1239): #{p < L or p > U}
            if (a.do_t) {
                const uint32_t tmask = a.seg.mask_b[s];
                __syncthreads();
                if (threadIdx.x < FMLP_MAX_CLASSES) s_t[threadIdx.x] = 0;
                __syncthreads();
                if (a.logits != nullptr && tmask != 0u) {
                    const int n_el = (int)(r_end - r_begin) * a.C;        // one CTA's row range: far below 2^31
                    const float* z = a.logits + r_begin * a.C;
                    for (int e = threadIdx.x; e < n_el; e += blockDim.x) {
                        const int c = e % a.C;
                        if ((tmask >> c) & 1u) {
                            const float v = z[e];
                            const float p = a.logits_are_probs ? v : sigmoid_ref(v);
                            if (p < a.L || p > a.U) atomicAdd(&s_t[c], 1);
                        }
                    }
                }
                __syncthreads();
                if (threadIdx.x < a.C) pc[2 * kProtoMaxActive + threadIdx.x] = s_t[threadIdx.x];
            }
        }
        row = r_end;
        if (row < cta_end) ++s;
    }
}

constexpr int kFinGroups = 16, kFinCols = 16;   // finalize CTA: 16 slot groups x 16 float4 columns

struct ProtoFinArgs {
    const float* partial;
    const int32_t* pcount;
    float* proto;        // [S][2C][D]
    int32_t* cnt;        // [S][2C]
    int32_t* tcnt;       // [S][C] or null
    int64_t rows_per_cta;
    SegTable seg;        // mask_a = active classes of THIS pass; mask_b = all active classes
    int D, C, NA, guard_empty, first_pass, has_t;
};

// grid.y = ceil(D / 64), 256 threads = 16 slot groups x 16 float4 columns.  Row 2c+y of
// segment s: the slot partials are added in a FIXED order (group g takes slots g, g+16, ... with
// 4 loads in flight, then the sixteen group sums are added in group order) -> deterministic, and
// the L2 round trips overlap instead of forming one 70-deep dependent chain (r01 ncu: 33 us).
// Round 2: 16 groups x 64 columns per CTA instead of 4 x 256 — a single client of 85,000 rows has ~440 slots
// and only 2 rows x D/256 CTAs with work, each walking 111 slots per group (~28 dependent batches).
// Then the row is divided by its count (tensor / python int -> fp32 divide, :997-999,1241-1248).
// grid.x = S * NA * 2: CTA (s, k, y) owns row 2c+y of segment s, c = the k-th class of this pass that is active on
// the segment (it leaves at once when the segment has fewer) — only rows with work get CTAs (a grid over all 2C rows
// launched 3,584 CTAs at C = 14, 16 of every 224 with anything to add).  The rows that stay zero, their counts and the
// t counts are the job of the (s, 0, 0) CTAs of column block 0.
__global__ void __launch_bounds__(256) proto_finalize_kernel(const __grid_constant__ ProtoFinArgs a) {
    asm volatile("griddepcontrol.wait;" ::: "memory");                // slot partials of the accumulate kernel
    __shared__ int s_n[8];
    __shared__ float4 s_part[kFinGroups][kFinCols];
    const int per_seg = 2 * a.NA;
    const int s = blockIdx.x / per_seg;
    const int ky = blockIdx.x - s * per_seg;
    const int slot_i = ky >> 1, y = ky & 1;                        // position among this pass's classes, label value
    const int grp = threadIdx.x / kFinCols;                        // slot group 0..kFinGroups-1
    const int cw = threadIdx.x % kFinCols;
    const int d = (blockIdx.y * kFinCols + cw) * 4;                // first of this thread's 4 columns
    const int64_t seg_lo = a.seg.rows[s], seg_hi = a.seg.rows[s + 1];
    const bool nonempty = seg_hi > seg_lo;
    const int64_t b0 = nonempty ? seg_lo / a.rows_per_cta : 0;
    const int64_t b1 = nonempty ? (seg_hi - 1) / a.rows_per_cta : -1;
    const uint32_t pass_active = a.seg.mask_a[s];

    if (ky == 0 && a.first_pass) {
        // ---- housekeeping of segment s (first pass only), spread over the column blocks ---------------
        const uint32_t active_any = a.seg.mask_b[s];
        // rows of classes that are not active on this client stay zero (proto = torch.zeros, :973): every column
        // block clears its own 64 columns of those rows (16 rows x 16 float4 per sweep)
        if (d < a.D) {
            for (int row = grp; row < 2 * a.C; row += kFinGroups) {
                if ((active_any >> (row >> 1)) & 1u) continue;
                *reinterpret_cast<float4*>(a.proto + ((int64_t)s * 2 * a.C + row) * a.D + d) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (blockIdx.y == 0) {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            if (threadIdx.x < 2 * a.C && !((active_any >> (threadIdx.x >> 1)) & 1u)) a.cnt[(int64_t)s * 2 * a.C + threadIdx.x] = 0;
            // t counts: one warp per class adds the slots
            if (a.tcnt != nullptr) {
                for (int c = warp; c < a.C; c += 8) {
                    int t = 0;
                    if (a.has_t)
                        for (int64_t b = b0 + lane; b <= b1; b += 32)
                            t += a.pcount[(b + s) * kProtoCountStride + 2 * kProtoMaxActive + c];
                    t = warp_sum_i(t);
                    if (lane == 0) a.tcnt[(int64_t)s * a.C + c] = t;
                }
            }
        }
    }

    // the slot_i-th class of this pass on this segment
    uint32_t m = pass_active;
    for (int i = 0; i < slot_i && m; ++i) m &= m - 1;
    if (m == 0u) return;                                           // fewer active classes here than NA
    const int c = __ffs(m) - 1;
    const int row = 2 * c + y;
    // row count: the slots are independent loads -> spread them over the CTA, then add
    int n = 0;
    for (int64_t b = b0 + threadIdx.x; b <= b1; b += blockDim.x) n += a.pcount[(b + s) * kProtoCountStride + 2 * slot_i + y];
    n = warp_sum_i(n);
    if ((threadIdx.x & 31) == 0) s_n[threadIdx.x >> 5] = n;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool col_ok = d < a.D;
    if (col_ok) {
        const float* p0 = a.partial + ((int64_t)slot_i * 2 + y) * (int64_t)a.D + d;
        const int64_t slot_stride = (int64_t)a.NA * 2 * a.D;
        int64_t b = b0 + grp;
        for (; b + 3 * kFinGroups <= b1; b += 4 * kFinGroups) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(p0 + (b + kFinGroups * u + s) * slot_stride);
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        for (; b <= b1; b += kFinGroups) {
            const float4 v = *reinterpret_cast<const float4*>(p0 + (b + s) * slot_stride);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    s_part[grp][cw] = acc;
    __syncthreads();
    n = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) n += s_n[w];
    if (blockIdx.y == 0 && threadIdx.x == 0) a.cnt[(int64_t)s * 2 * a.C + row] = n;
    if (grp != 0 || !col_ok) return;
    float4 t = s_part[0][cw];
#pragma unroll
    for (int g = 1; g < kFinGroups; ++g) {
        const float4 v = s_part[g][cw];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    if (!(n == 0 && a.guard_empty)) {
        const float fn = (float)n;
        t.x = __fdiv_rn(t.x, fn); t.y = __fdiv_rn(t.y, fn); t.z = __fdiv_rn(t.z, fn); t.w = __fdiv_rn(t.w, fn);
    }
    *reinterpret_cast<float4*>(a.proto + ((int64_t)s * 2 * a.C + row) * a.D + d) = t;
}

template <int NA>
static int proto_ctas_per_sm(int threads) {
    static int cached_threads = 0, cached = 0;
    if (cached_threads != threads) {
        int b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, proto_accum_kernel<NA>, threads, 0) != cudaSuccess) b = 1;
        cached = b < 1 ? 1 : (b > kProtoMaxCtasPerSm ? kProtoMaxCtasPerSm : b);
        cached_threads = threads;
    }
    return cached;
}

}  // namespace fmlp

using namespace fmlp;

static size_t proto_slot_bytes(int D) {
    return (size_t)kProtoMaxActive * 2 * (size_t)D * sizeof(float);
}

extern "C" size_t fmlp_proto_ws_bytes(int64_t n_total, int D, int C, int S) {
    (void)C; (void)n_total;
    if (D < 1 || S < 1) return 0;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const size_t slots = (size_t)sms * kProtoMaxCtasPerSm + (size_t)S;
    return slots * (proto_slot_bytes(D) + kProtoCountStride * sizeof(int32_t));
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
    if (ld_feat > 0x7fffffffll) return FMLP_ERR_UNSUPPORTED;      // the fast loop advances by 32-bit element strides
    if (S < 1 || S > FMLP_MAX_SEGMENTS || !seg_rows) return FMLP_ERR_BAD_ARG;
    const uint32_t cmask = C < 32 ? ((1u << C) - 1u) : 0xffffffffu;
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = sm_count();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    const size_t max_slots = (size_t)sms * kProtoMaxCtasPerSm + (size_t)S;
    if (ws_bytes < max_slots * (proto_slot_bytes(D) + kProtoCountStride * sizeof(int32_t))) return FMLP_ERR_WORKSPACE;

    uint32_t act_all[FMLP_MAX_SEGMENTS], act_left[FMLP_MAX_SEGMENTS], tc[FMLP_MAX_SEGMENTS];
    for (int s = 0; s < S; ++s) {
        act_all[s] = seg_active[s] & cmask;
        act_left[s] = act_all[s];
        tc[s] = (seg_tcount ? seg_tcount[s] : 0u) & cmask;
    }
    const int64_t n_total = seg_rows[S];

    ProtoArgs a;
    a.feat = feat; a.labels = labels; a.logits = logits; a.partial = (float*)ws;
    a.pcount = (int32_t*)((char*)ws + max_slots * proto_slot_bytes(D));
    a.ld_feat = ld_feat; a.L = L; a.U = U; a.D = D; a.C = C; a.logits_are_probs = logits_are_probs;

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
        int rc = fill_seg_table(a.seg, S, seg_rows, pass, first ? tc : nullptr);
        if (rc != FMLP_OK) return rc;
        a.do_t = (first && logits != nullptr && tcnt != nullptr) ? 1 : 0;
        const int NA = na_max <= 1 ? 1 : (na_max == 2 ? 2 : 4);
        // one balanced wave: as many CTAs as can be co-resident, each with a contiguous row range
        int per_sm = NA == 1 ? proto_ctas_per_sm<1>(threads) : (NA == 2 ? proto_ctas_per_sm<2>(threads) : proto_ctas_per_sm<4>(threads));
        int64_t gx = (int64_t)sms * per_sm / gy;
        if (gx < 1) gx = 1;
        const int64_t min_rows = 2 * kProtoUnroll;
        if (gx > (n_total + min_rows - 1) / min_rows) gx = (n_total + min_rows - 1) / min_rows;
        if (gx < 1) gx = 1;
        int64_t rpc = (n_total + gx - 1) / gx;
        rpc = (rpc + kProtoUnroll - 1) / kProtoUnroll * kProtoUnroll;
        if (rpc < kProtoUnroll) rpc = kProtoUnroll;
        a.rows_per_cta = rpc;
        if (n_total > 0) {
            gx = (n_total + rpc - 1) / rpc;
            dim3 grid((unsigned)gx, (unsigned)gy);
            // scheduling knob (include/fedmlp_b200.h): unused dynamic shared memory per CTA steers co-residency with
            // the similarity kernel of a concurrent stream
            const size_t dsm = (size_t)tuning_value(FMLP_TUNE_PROTO_PAD_SMEM_KB, "FMLP_PROTO_PAD_SMEM_KB", 0, 56, 0) * 1024;
            if (dsm > 48u * 1024u) {
                static bool raised = false;
                if (!raised) {
                    cudaFuncSetAttribute(proto_accum_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024);
                    cudaFuncSetAttribute(proto_accum_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024);
                    cudaFuncSetAttribute(proto_accum_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024);
                    raised = true;
                }
            }
            if (NA == 1) proto_accum_kernel<1><<<grid, threads, dsm, st>>>(a);
            else if (NA == 2) proto_accum_kernel<2><<<grid, threads, dsm, st>>>(a);
            else proto_accum_kernel<4><<<grid, threads, dsm, st>>>(a);
            rc = launch_status();
            if (rc != FMLP_OK) return rc;
        }
        ProtoFinArgs f;
        f.partial = a.partial; f.pcount = a.pcount; f.proto = proto; f.cnt = cnt; f.tcnt = tcnt;
        f.rows_per_cta = rpc;
        rc = fill_seg_table(f.seg, S, seg_rows, pass, act_all);
        if (rc != FMLP_OK) return rc;
        f.D = D; f.C = C; f.NA = NA; f.guard_empty = guard_empty; f.first_pass = first ? 1 : 0; f.has_t = a.do_t;
        dim3 fgrid((unsigned)(S * 2 * NA), (unsigned)((D + 4 * kFinCols - 1) / (4 * kFinCols)));
        {
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cfg.gridDim = fgrid; cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
            cudaError_t e = cudaLaunchKernelEx(&cfg, proto_finalize_kernel, f);
            rc = e == cudaSuccess ? launch_status() : (int)e;
        }
        if (rc != FMLP_OK) return rc;
        first = false;
        bool more = false;
        for (int s = 0; s < S; ++s) more = more || act_left[s] != 0;
        if (!more) break;
    }
    return FMLP_OK;
}
