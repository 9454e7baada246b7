// loss.cu — K4: FedMLP training losses, forward value + gradient w.r.t. the logits in one launch.
//
// Replaces the op chains at utils/local_training.py:933-963 (stage 1) and :1171-1188 (stage 2)
// together with utils/FedNoRo.py:16-22 (LogitAdjust_Multilabel == F.binary_cross_entropy on
// probabilities) and the autograd graph torch builds behind them (~30 forward + ~40 backward
// element-wise launches on [32, C] tensors).  Everything is evaluated with the formulas ATen
// uses, in ATen's operation order, so the gradient agrees with autograd to the last ulp or two:
//   forward   BCE(p,y) = (y-1)*max(log1p(-p),-100) - y*max(log(p),-100)
//   backward  dBCE/dp  = g*(p-y) / max((1-p)*p, 1e-12)         (binary_cross_entropy_backward)
//             dMSE/dp  = (2*(p-t))*g                             (mse_loss_backward, reduction none)
//             dp/dz    = (g*(1-p))*p                             (sigmoid_backward)
// The stand-alone loss kernels are launched cooperatively: phase boundaries (stage-2 denominator, final
// loss reduction) are grid barriers, per-CTA partials are combined in CTA order -> deterministic, no
// atomics, no pre-zeroed workspace.  The batched round uses fill_loss_stage2_kernel below (round 2): mask
// fill fused in, ordinary launch, last-arriving CTA does the final reduction.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace fmlp {

constexpr int kLossThreads = 256;
constexpr int kLossUnroll = 4;
constexpr int kLossMaxGrid = 2048;  // workspace: 4 arrays of kLossMaxGrid 32-bit words

__device__ __forceinline__ float bce_fwd(float p, float y) {
    return __fsub_rn(__fmul_rn(__fsub_rn(y, 1.f), fmaxf(log1pf(-p), -100.f)),
                     __fmul_rn(y, fmaxf(logf(p), -100.f)));
}
__device__ __forceinline__ float bce_bwd(float g, float p, float y) {
    return __fdiv_rn(__fmul_rn(g, __fsub_rn(p, y)), fmaxf(__fmul_rn(__fsub_rn(1.f, p), p), 1e-12f));
}
__device__ __forceinline__ float mse_bwd(float g, float p, float t) {
    return __fmul_rn(__fmul_rn(2.f, __fsub_rn(p, t)), g);
}
__device__ __forceinline__ float sigmoid_bwd(float g, float p) {
    return __fmul_rn(__fmul_rn(g, __fsub_rn(1.f, p)), p);
}

// Sum `v` over the CTA; valid in thread 0.
__device__ __forceinline__ float block_sum(float v, float* s_red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.f;
        r = warp_sum(r);
    }
    return r;
}
__device__ __forceinline__ int block_sum_i(int v, int* s_red) {
    v = warp_sum_i(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = 0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0;
        r = warp_sum_i(r);
    }
    return r;
}

struct Loss1Args {
    const float *z1, *z2, *z3, *z4, *y;
    float *loss, *dz1, *dz2;
    float* ws;
    int64_t n;  // B*C
    int C;
    uint32_t active, missing;
    float den_sup, den_dis;  // (float)(bs*A), (float)(bs*M)
};

__global__ void __launch_bounds__(kLossThreads) loss_stage1_kernel(const __grid_constant__ Loss1Args a) {
    __shared__ float s_red[32];
    float sup = 0.f, dis = 0.f;
    // upstream gradients of the two sums (autograd: grad/den, then /2)
    const float g_sup = __fmul_rn(__fdiv_rn(1.f, a.den_sup), 0.5f);
    const float g_dis = __fmul_rn(__fdiv_rn(1.f, a.den_dis), 0.5f);
    const int64_t stride = (int64_t)gridDim.x * kLossThreads * kLossUnroll;
    for (int64_t base = (int64_t)blockIdx.x * kLossThreads * kLossUnroll + threadIdx.x; base < a.n; base += stride) {
        float v1[kLossUnroll], v2[kLossUnroll], v3[kLossUnroll], v4[kLossUnroll], vy[kLossUnroll];
#pragma unroll
        for (int u = 0; u < kLossUnroll; ++u) {
            const int64_t e = base + (int64_t)u * kLossThreads;
            const bool ok = e < a.n;
            v1[u] = ok ? a.z1[e] : 0.f; v2[u] = ok ? a.z2[e] : 0.f;
            v3[u] = ok ? a.z3[e] : 0.f; v4[u] = ok ? a.z4[e] : 0.f;
            vy[u] = ok ? a.y[e] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kLossUnroll; ++u) {
            const int64_t e = base + (int64_t)u * kLossThreads;
            if (e >= a.n) continue;
            const int c = (int)(e % a.C);
            const float p1 = sigmoid_ref(v1[u]), p2 = sigmoid_ref(v2[u]);
            float d1 = 0.f, d2 = 0.f;
            if ((a.active >> c) & 1u) {
                // (loss_sup1 + loss_sup2) / 2.   (:952-954)
                sup += __fmul_rn(__fadd_rn(bce_fwd(p1, vy[u]), bce_fwd(p2, vy[u])), 0.5f);
                d1 = sigmoid_bwd(bce_bwd(g_sup, p1, vy[u]), p1);
                d2 = sigmoid_bwd(bce_bwd(g_sup, p2, vy[u]), p2);
            } else if ((a.missing >> c) & 1u) {
                const float p3 = sigmoid_ref(v3[u]), p4 = sigmoid_ref(v4[u]);
                const float e1 = __fsub_rn(p1, p3), e2 = __fsub_rn(p2, p4);
                // (loss_dis1 + loss_dis2) / 2.   (:949-951)
                dis += __fmul_rn(__fadd_rn(__fmul_rn(e1, e1), __fmul_rn(e2, e2)), 0.5f);
                d1 = sigmoid_bwd(mse_bwd(g_dis, p1, p3), p1);
                d2 = sigmoid_bwd(mse_bwd(g_dis, p2, p4), p2);
            }
            a.dz1[e] = d1;
            a.dz2[e] = d2;
        }
    }
    const float bs_sup = block_sum(sup, s_red);
    const float bs_dis = block_sum(dis, s_red);
    if (threadIdx.x == 0) { a.ws[blockIdx.x] = bs_sup; a.ws[kLossMaxGrid + blockIdx.x] = bs_dis; }
    cg::this_grid().sync();
    if (blockIdx.x == 0) {
        float ts = 0.f, td = 0.f;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += kLossThreads) { ts += a.ws[b]; td += a.ws[kLossMaxGrid + b]; }
        ts = block_sum(ts, s_red);
        td = block_sum(td, s_red);
        if (threadIdx.x == 0) {
            // loss = loss_sup + 0.0*loss_unsup + loss_dis   (:956-963); the 0.0* term is +0
            const float l_sup = __fdiv_rn(ts, a.den_sup);
            const float l_dis = __fdiv_rn(td, a.den_dis);
            a.loss[0] = __fadd_rn(__fadd_rn(l_sup, 0.f), l_dis);
        }
    }
}

struct Loss2Args {
    const float *z, *zg, *y, *distill;
    float *loss, *dz;   // loss[S]
    float* ws;
    const int32_t* seg_class_distill;  // [S][C] number of distill==1 entries per (segment, class), or null
    int64_t el_per_cta;  // elements handled by one CTA (contiguous range)
    int C;
    int variant;
    SegTable seg;        // rows prefix per segment (client); masks unused
};

// Segmented stage-2 loss: every segment (client) has its own denominator sum(sup) and its own
// scalar loss, all in ONE cooperative launch.  CTA b owns the contiguous element range
// [b*E, (b+1)*E); the part of it inside segment s is reduced on its own and published in
// workspace slot b + s (injective because both b and s grow along the range), so the partials of
// a segment are combined in CTA order -> deterministic, no atomics, nothing to pre-zero.
__global__ void __launch_bounds__(kLossThreads) loss_stage2_kernel(const __grid_constant__ Loss2Args a) {
    __shared__ float s_red[32];
    __shared__ int s_redi[32];
    __shared__ float s_den[2];
    cg::grid_group grid = cg::this_grid();
    int* wsi = reinterpret_cast<int*>(a.ws);
    const int S = a.seg.S;
    const int64_t n_el = a.seg.rows[S] * a.C;
    const int64_t e_begin = (int64_t)blockIdx.x * a.el_per_cta;
    const int64_t e_end = min(n_el, e_begin + a.el_per_cta);
    const int s_first = e_begin < n_el ? find_segment(a.seg.rows, S, e_begin / a.C) : S;

    // ---- phase 1: per (CTA, segment) count of distilled entries; sup/distill are 0/1 masks, so
    //      the denominators are exact integers.  Skipped (with its grid barrier) when the caller
    //      already knows the counts (fmlp_tag_select reports them per client and class).
    const bool counted = a.seg_class_distill != nullptr;
    for (int s = counted ? S : s_first; s < S; ++s) {
        const int64_t lo = max(e_begin, a.seg.rows[s] * a.C), hi = min(e_end, a.seg.rows[s + 1] * a.C);
        if (lo >= e_end) break;
        int n_dis = 0;
        for (int64_t e = lo + threadIdx.x; e < hi; e += kLossThreads) n_dis += (a.distill[e] != 0.f) ? 1 : 0;
        const int b_dis = block_sum_i(n_dis, s_redi);
        if (threadIdx.x == 0) wsi[2 * kLossMaxGrid + blockIdx.x + s] = b_dis;
    }
    if (!counted) grid.sync();

    // ---- phase 2: numerators + gradient, segment by segment -----------------------------------
    for (int s = s_first; s < S; ++s) {
        const int64_t seg_lo = a.seg.rows[s] * a.C, seg_hi = a.seg.rows[s + 1] * a.C;
        const int64_t lo = max(e_begin, seg_lo), hi = min(e_end, seg_hi);
        if (lo >= e_end) break;
        if (hi <= lo) continue;  // empty segment
        // denominator of segment s: add the counts of every CTA that touches it, in CTA order
        const int b0 = (int)(seg_lo / a.el_per_cta), b1 = (int)((seg_hi - 1) / a.el_per_cta);
        int td = 0;
        if (counted) {
            if (threadIdx.x < a.C) td = a.seg_class_distill[(int64_t)s * a.C + threadIdx.x];
        } else {
            for (int b = b0 + threadIdx.x; b <= b1; b += kLossThreads) td += wsi[2 * kLossMaxGrid + b + s];
        }
        td = block_sum_i(td, s_redi);
        if (threadIdx.x == 0) {
            const float sum_dis = (float)td, sum_sup = (float)((seg_hi - seg_lo) - td);
            // :1188  sup_cls.sum()          :1187  sup_cls.sum() + distill_cls.sum()
            s_den[0] = a.variant == FMLP_LOSS2_SUP ? sum_sup : __fadd_rn(sum_sup, sum_dis);
        }
        __syncthreads();
        const float den = s_den[0];
        const float g = __fdiv_rn(1.f, den);  // d loss / d numerator
        float num_sup = 0.f, num_dis = 0.f;
        for (int64_t base = lo + threadIdx.x; base < hi; base += (int64_t)kLossThreads * kLossUnroll) {
            float vz[kLossUnroll], vg[kLossUnroll], vy[kLossUnroll], vd[kLossUnroll];
#pragma unroll
            for (int u = 0; u < kLossUnroll; ++u) {
                const int64_t e = base + (int64_t)u * kLossThreads;
                const bool ok = e < hi;
                vz[u] = ok ? a.z[e] : 0.f;
                vy[u] = ok ? a.y[e] : 0.f;
                vd[u] = ok ? a.distill[e] : 0.f;
                vg[u] = (ok && a.variant == FMLP_LOSS2_SUP_DIS) ? a.zg[e] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < kLossUnroll; ++u) {
                const int64_t e = base + (int64_t)u * kLossThreads;
                if (e >= hi) continue;
                const float p = sigmoid_ref(vz[u]);
                const float dist = vd[u];
                const float sup = (dist != 0.f) ? 0.f : 1.f;  // (~distill_cls.bool()).float()  (:1173)
                // loss_sup * sup_cls, backward: (g * sup) into BCE backward
                num_sup += __fmul_rn(bce_fwd(p, vy[u]), sup);
                float gp = bce_bwd(__fmul_rn(g, sup), p, vy[u]);
                if (a.variant == FMLP_LOSS2_SUP_DIS) {
                    const float pg = sigmoid_ref(vg[u]);
                    const float d = __fsub_rn(p, pg);
                    num_dis += __fmul_rn(__fmul_rn(d, d), dist);
                    gp = __fadd_rn(gp, mse_bwd(__fmul_rn(g, dist), p, pg));
                }
                a.dz[e] = sigmoid_bwd(gp, p);
            }
        }
        const float bn_sup = block_sum(num_sup, s_red);
        const float bn_dis = block_sum(num_dis, s_red);
        if (threadIdx.x == 0) { a.ws[blockIdx.x + s] = bn_sup; a.ws[kLossMaxGrid + blockIdx.x + s] = bn_dis; }
        __syncthreads();  // s_den reuse
    }
    grid.sync();

    // ---- final: one CTA per segment (round-robin) adds the partials in CTA order ------------
    for (int s = blockIdx.x; s < S; s += gridDim.x) {
        const int64_t seg_lo = a.seg.rows[s] * a.C, seg_hi = a.seg.rows[s + 1] * a.C;
        float ts = 0.f, tdis = 0.f;
        int td = 0;
        if (seg_hi > seg_lo) {
            const int b0 = (int)(seg_lo / a.el_per_cta), b1 = (int)((seg_hi - 1) / a.el_per_cta);
            for (int b = b0 + threadIdx.x; b <= b1; b += kLossThreads) {
                ts += a.ws[b + s]; tdis += a.ws[kLossMaxGrid + b + s];
                if (!counted) td += wsi[2 * kLossMaxGrid + b + s];
            }
            if (counted && threadIdx.x < a.C) td = a.seg_class_distill[(int64_t)s * a.C + threadIdx.x];
        }
        ts = block_sum(ts, s_red);
        tdis = block_sum(tdis, s_red);
        td = block_sum_i(td, s_redi);
        if (threadIdx.x == 0) {
            const float sum_dis = (float)td, sum_sup = (float)((seg_hi - seg_lo) - td);
            const float den = a.variant == FMLP_LOSS2_SUP ? sum_sup : __fadd_rn(sum_sup, sum_dis);
            const float num = a.variant == FMLP_LOSS2_SUP ? ts : __fadd_rn(ts, tdis);
            a.loss[s] = __fdiv_rn(num, den);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Round 2: label / mask fill (DatasetSplit_pseudo.__getitem__, utils/local_training.py:1456-1477, :1173)
// FUSED into the stage-2 loss (:1171-1188), as an ordinary (non-cooperative) launch.
//   * y / distill_cls / sup_cls are computed on the fly from the true labels and the tag state the select
//     kernel just updated, and written out as by-products (the training loop and the tests still read them):
//     one pass over [N, C] instead of two, one launch instead of two;
//   * the denominators come from the select kernel's per-(client, class) candidate counts, so no counting
//     phase and no grid barrier in front of the gradient;
//   * the final per-client reduction is done by whichever CTA finishes LAST (a counter in the workspace,
//     reset by that CTA), not behind a cooperative grid.sync(): the launch needs no co-residency, so it
//     overlaps the kernels of the other streams, and it is chained to the select kernel by a programmatic
//     dependent launch.  Partials are still combined in CTA order -> deterministic.
struct FillLossArgs {
    const float *z, *zg, *labels;
    const uint8_t* tag;
    float *y, *distill, *sup;          // by-products, any may be null
    float *loss, *dz;                  // loss[S]
    float* ws;                         // [0,G) sup partials, [G,2G) dis partials, word 3G = arrival counter (zero at first use)
    const int32_t* seg_class_distill;  // [S][C] rows with tag == 0 per (client, class) = distilled entries of missing classes
    int64_t ld_tag;
    int64_t el_per_cta;
    int C;
    int variant;
    SegTable seg;                      // mask_a = active classes, mask_b = missing classes
};

__global__ void __launch_bounds__(kLossThreads) fill_loss_stage2_kernel(const __grid_constant__ FillLossArgs a) {
    __shared__ float s_red[32];
    __shared__ float s_den[2];
    __shared__ int s_last;
    asm volatile("griddepcontrol.wait;" ::: "memory");     // tag state / counts of the select kernel
    const int S = a.seg.S;
    const int64_t n_el = a.seg.rows[S] * a.C;
    const int64_t e_begin = (int64_t)blockIdx.x * a.el_per_cta;
    const int64_t e_end = min(n_el, e_begin + a.el_per_cta);
    const int s_first = e_begin < n_el ? find_segment(a.seg.rows, S, e_begin / a.C) : S;

    for (int s = s_first; s < S; ++s) {
        const int64_t seg_lo = a.seg.rows[s] * a.C, seg_hi = a.seg.rows[s + 1] * a.C;
        const int64_t lo = max(e_begin, seg_lo), hi = min(e_end, seg_hi);
        if (lo >= e_end) break;
        if (hi <= lo) continue;  // empty segment
        if (threadIdx.x < 32) {      // C <= 32: one warp adds the client's per-class counts
            int td = 0;
            if (threadIdx.x < a.C && ((a.seg.mask_b[s] >> threadIdx.x) & 1u)) td = a.seg_class_distill[(int64_t)s * a.C + threadIdx.x];
            td = warp_sum_i(td);
            if (threadIdx.x == 0) {
                const float sum_dis = (float)td, sum_sup = (float)((seg_hi - seg_lo) - td);
                // :1188  sup_cls.sum()          :1187  sup_cls.sum() + distill_cls.sum()
                s_den[0] = a.variant == FMLP_LOSS2_SUP ? sum_sup : __fadd_rn(sum_sup, sum_dis);
            }
        }
        __syncthreads();
        const float den = s_den[0];
        const float g = __fdiv_rn(1.f, den);  // d loss / d numerator
        const uint32_t act = a.seg.mask_a[s], mis = a.seg.mask_b[s];
        float num_sup = 0.f, num_dis = 0.f;
        // (row, class) of this thread's elements are advanced incrementally: element e + 256 is (row + 256 / C,
        // c + 256 % C) with one wrap — the first version divided a 64-bit index by C for every element (r02 ncu:
        // 307 thread instructions per element)
        const int step_r = kLossThreads / a.C, step_c = kLossThreads - step_r * a.C;
        int64_t row = (lo + threadIdx.x) / a.C;
        int c = (int)((lo + threadIdx.x) - row * a.C);
        for (int64_t base = lo + threadIdx.x; base < hi; base += (int64_t)kLossThreads * kLossUnroll) {
            float vz[kLossUnroll], vg[kLossUnroll], vy[kLossUnroll], vd[kLossUnroll];
#pragma unroll
            for (int u = 0; u < kLossUnroll; ++u) {
                const int64_t e = base + (int64_t)u * kLossThreads;
                const bool ok = e < hi;
                vz[u] = ok ? a.z[e] : 0.f;
                vg[u] = (ok && a.variant == FMLP_LOSS2_SUP_DIS) ? a.zg[e] : 0.f;
                float y = 0.f, dis = 0.f;
                if (ok) {
                    if ((act >> c) & 1u) {
                        y = a.labels[e];                              // annotated class keeps its label (:1458-1460)
                    } else if ((mis >> c) & 1u) {
                        const uint8_t t = a.tag[(int64_t)c * a.ld_tag + row];
                        y = (t == 2) ? 1.f : 0.f;                     // in the noise list -> pseudo-positive (:1464-1466)
                        dis = (t == 0) ? 1.f : 0.f;                   // in neither list -> distilled (:1467-1468)
                    }
                    if (a.y) a.y[e] = y;
                    if (a.distill) a.distill[e] = dis;
                    if (a.sup) a.sup[e] = 1.f - dis;                  // sup_cls = ~distill_cls (:1173)
                }
                vy[u] = y; vd[u] = dis;
                row += step_r; c += step_c;
                if (c >= a.C) { c -= a.C; ++row; }
            }
#pragma unroll
            for (int u = 0; u < kLossUnroll; ++u) {
                const int64_t e = base + (int64_t)u * kLossThreads;
                if (e >= hi) continue;
                const float p = sigmoid_ref(vz[u]);
                const float dist = vd[u];
                const float sup = (dist != 0.f) ? 0.f : 1.f;
                num_sup += __fmul_rn(bce_fwd(p, vy[u]), sup);
                float gp = bce_bwd(__fmul_rn(g, sup), p, vy[u]);
                if (a.variant == FMLP_LOSS2_SUP_DIS) {
                    const float pg = sigmoid_ref(vg[u]);
                    const float d = __fsub_rn(p, pg);
                    num_dis += __fmul_rn(__fmul_rn(d, d), dist);
                    gp = __fadd_rn(gp, mse_bwd(__fmul_rn(g, dist), p, pg));
                }
                a.dz[e] = sigmoid_bwd(gp, p);
            }
        }
        const float bn_sup = block_sum(num_sup, s_red);
        const float bn_dis = block_sum(num_dis, s_red);
        if (threadIdx.x == 0) { a.ws[blockIdx.x + s] = bn_sup; a.ws[kLossMaxGrid + blockIdx.x + s] = bn_dis; }
        __syncthreads();  // s_den reuse
    }
    // ---- the CTA that arrives last adds every client's partials in CTA order ----------------------
    unsigned int* counter = reinterpret_cast<unsigned int*>(a.ws + 3 * kLossMaxGrid);
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(counter, 1u) + 1u == gridDim.x) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // one warp per client: lane-strided sums over the client's CTA slots, then a shuffle tree (a fixed order for
    // a given launch geometry -> deterministic); no block-wide barriers on the tail of the kernel
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = warp; s < S; s += kLossThreads / 32) {
        const int64_t seg_lo = a.seg.rows[s] * a.C, seg_hi = a.seg.rows[s + 1] * a.C;
        float ts = 0.f, tdis = 0.f;
        int td = 0;
        if (seg_hi > seg_lo) {
            const int b0 = (int)(seg_lo / a.el_per_cta), b1 = (int)((seg_hi - 1) / a.el_per_cta);
            for (int b = b0 + lane; b <= b1; b += 32) {
                ts += __ldcg(a.ws + b + s); tdis += __ldcg(a.ws + kLossMaxGrid + b + s);
            }
            if (lane < a.C && ((a.seg.mask_b[s] >> lane) & 1u)) td = a.seg_class_distill[(int64_t)s * a.C + lane];
        }
        ts = warp_sum(ts);
        tdis = warp_sum(tdis);
        td = warp_sum_i(td);
        if (lane == 0) {
            const float sum_dis = (float)td, sum_sup = (float)((seg_hi - seg_lo) - td);
            const float den = a.variant == FMLP_LOSS2_SUP ? sum_sup : __fadd_rn(sum_sup, sum_dis);
            const float num = a.variant == FMLP_LOSS2_SUP ? ts : __fadd_rn(ts, tdis);
            a.loss[s] = __fdiv_rn(num, den);
        }
    }
    if (threadIdx.x == 0) *counter = 0u;     // ready for the next launch (stream order makes it visible)
}

__global__ void scale_kernel(float* x, int64_t n, const float* scale) {
    const float s = *scale;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = __fmul_rn(x[i], s);
}

template <typename Kern>
static int coop_grid(Kern kern, int64_t n, int* grid_out) {
    static int blocks_per_sm = 0;  // per kernel instantiation
    if (blocks_per_sm == 0) {
        int b = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kLossThreads, 0);
        if (e != cudaSuccess) return (int)e;
        blocks_per_sm = b > 0 ? b : 1;
    }
    const int sms = sm_count();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    int64_t want = (n + (int64_t)kLossThreads * kLossUnroll - 1) / ((int64_t)kLossThreads * kLossUnroll);
    int64_t cap = (int64_t)sms * blocks_per_sm;
    if (cap > kLossMaxGrid) cap = kLossMaxGrid;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    *grid_out = (int)want;
    return FMLP_OK;
}

}  // namespace fmlp

using namespace fmlp;

extern "C" size_t fmlp_loss_ws_bytes(int64_t B, int C) {
    (void)B; (void)C;
    return (size_t)4 * kLossMaxGrid * sizeof(float);
}

extern "C" int fmlp_loss_stage1_f32(const float* z1, const float* z2, const float* z3, const float* z4,
                                    const float* y, int64_t B, int C, uint32_t active, uint32_t missing,
                                    int bs, float* loss, float* dz1, float* dz2, void* ws,
                                    size_t ws_bytes, fmlp_stream_t stream) {
    if (!z1 || !z2 || !z3 || !z4 || !y || !loss || !dz1 || !dz2 || !ws || B < 0 || C < 1 ||
        C > FMLP_MAX_CLASSES || bs < 1)
        return FMLP_ERR_BAD_ARG;
    if (ws_bytes < fmlp_loss_ws_bytes(B, C)) return FMLP_ERR_WORKSPACE;
    const uint32_t cmask = C < 32 ? ((1u << C) - 1u) : 0xffffffffu;
    Loss1Args a;
    a.z1 = z1; a.z2 = z2; a.z3 = z3; a.z4 = z4; a.y = y; a.loss = loss; a.dz1 = dz1; a.dz2 = dz2;
    a.ws = (float*)ws; a.n = B * C; a.C = C; a.active = active & cmask; a.missing = missing & cmask & ~active;
    // self.args.batch_size * self.args.annotation_num / * len(negetive_class_list_client) (:956-959)
    a.den_sup = (float)((int64_t)bs * __builtin_popcount(a.active));
    a.den_dis = (float)((int64_t)bs * __builtin_popcount(a.missing));
    int grid = 1;
    int rc = coop_grid(loss_stage1_kernel, a.n, &grid);
    if (rc != FMLP_OK) return rc;
    void* args[] = {(void*)&a};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)loss_stage1_kernel, dim3(grid), dim3(kLossThreads),
                                                args, 0, (cudaStream_t)stream);
    return e == cudaSuccess ? launch_status() : (int)e;
}

static int launch_loss2(const float* z, const float* zg, const float* y, const float* distill, int C, int S,
                        const int64_t* seg_rows, int variant, const int32_t* seg_class_distill, float* loss, float* dz,
                        void* ws, size_t ws_bytes, fmlp_stream_t stream) {
    if (!z || !y || !distill || !loss || !dz || !ws || C < 1 || C > FMLP_MAX_CLASSES) return FMLP_ERR_BAD_ARG;
    if (variant != FMLP_LOSS2_SUP && variant != FMLP_LOSS2_SUP_DIS) return FMLP_ERR_BAD_ARG;
    if (variant == FMLP_LOSS2_SUP_DIS && !zg) return FMLP_ERR_BAD_ARG;
    if (ws_bytes < fmlp_loss_ws_bytes(0, C)) return FMLP_ERR_WORKSPACE;
    Loss2Args a;
    int rc = fill_seg_table(a.seg, S, seg_rows, nullptr, nullptr);
    if (rc != FMLP_OK) return rc;
    a.z = z; a.zg = zg; a.y = y; a.distill = distill; a.loss = loss; a.dz = dz; a.ws = (float*)ws;
    a.C = C; a.variant = variant; a.seg_class_distill = seg_class_distill;
    const int64_t n_el = seg_rows[S] * C;
    int grid = 1;
    rc = coop_grid(loss_stage2_kernel, n_el, &grid);
    if (rc != FMLP_OK) return rc;
    // two grid barriers + per-segment slot sums: fewer, fatter CTAs are cheaper here (one per SM)
    const int sms = sm_count();
    if (sms > 0 && grid > sms) grid = sms;
    if (grid > kLossMaxGrid - FMLP_MAX_SEGMENTS) grid = kLossMaxGrid - FMLP_MAX_SEGMENTS;  // slots b + s
    const int64_t tile = (int64_t)kLossThreads * kLossUnroll;
    int64_t epc = (n_el + grid - 1) / grid;
    epc = (epc + tile - 1) / tile * tile;
    if (epc < tile) epc = tile;
    a.el_per_cta = epc;
    void* args[] = {(void*)&a};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)loss_stage2_kernel, dim3(grid), dim3(kLossThreads),
                                                args, 0, (cudaStream_t)stream);
    return e == cudaSuccess ? launch_status() : (int)e;
}

extern "C" int fmlp_loss_stage2_f32(const float* z, const float* zg, const float* y, const float* distill,
                                    int64_t B, int C, int variant, float* loss, float* dz, void* ws,
                                    size_t ws_bytes, fmlp_stream_t stream) {
    if (B < 0) return FMLP_ERR_BAD_ARG;
    const int64_t rows[2] = {0, B};
    return launch_loss2(z, zg, y, distill, C, 1, rows, variant, nullptr, loss, dz, ws, ws_bytes, stream);
}

extern "C" int fmlp_loss_stage2_seg_f32(const float* z, const float* zg, const float* y, const float* distill,
                                        int C, int S, const int64_t* seg_rows, int variant,
                                        const int32_t* seg_class_distill, float* loss, float* dz, void* ws,
                                        size_t ws_bytes, fmlp_stream_t stream) {
    return launch_loss2(z, zg, y, distill, C, S, seg_rows, variant, seg_class_distill, loss, dz, ws, ws_bytes, stream);
}

extern "C" int fmlp_fill_loss_stage2_f32(const float* labels, const uint8_t* tag, int64_t ld_tag, const float* z,
                                         const float* zg, int C, int S, const int64_t* seg_rows,
                                         const uint32_t* seg_active, const uint32_t* seg_missing, int variant,
                                         const int32_t* seg_class_distill, float* y, float* distill, float* sup,
                                         float* loss, float* dz, void* ws, size_t ws_bytes, fmlp_stream_t stream) {
    if (!labels || !tag || !z || !seg_active || !seg_missing || !seg_class_distill || !loss || !dz || !ws || C < 1 ||
        C > FMLP_MAX_CLASSES)
        return FMLP_ERR_BAD_ARG;
    if (variant != FMLP_LOSS2_SUP && variant != FMLP_LOSS2_SUP_DIS) return FMLP_ERR_BAD_ARG;
    if (variant == FMLP_LOSS2_SUP_DIS && !zg) return FMLP_ERR_BAD_ARG;
    if (ws_bytes < fmlp_loss_ws_bytes(0, C)) return FMLP_ERR_WORKSPACE;
    FillLossArgs a;
    int rc = fill_seg_table(a.seg, S, seg_rows, seg_active, seg_missing);
    if (rc != FMLP_OK) return rc;
    if (ld_tag < seg_rows[S]) return FMLP_ERR_BAD_ARG;
    a.z = z; a.zg = zg; a.labels = labels; a.tag = tag; a.y = y; a.distill = distill; a.sup = sup; a.loss = loss; a.dz = dz;
    a.ws = (float*)ws; a.seg_class_distill = seg_class_distill; a.ld_tag = ld_tag; a.C = C; a.variant = variant;
    const int64_t n_el = seg_rows[S] * C;
    const int sms = sm_count();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    const int64_t tile = (int64_t)kLossThreads * kLossUnroll;
    int64_t grid = (n_el + tile - 1) / tile;
    // one 1024-element tile per CTA while all CTAs are co-resident (8 x 256 threads per SM): every extra tile per CTA
    // is another dependent round of L2 latency in a latency-bound kernel (r02: 2 CTAs per SM meant 3 rounds at C = 14)
    if (grid > 8 * (int64_t)sms) grid = 8 * (int64_t)sms;
    if (grid > kLossMaxGrid - FMLP_MAX_SEGMENTS) grid = kLossMaxGrid - FMLP_MAX_SEGMENTS;  // slots b + s
    if (grid < 1) grid = 1;
    int64_t epc = (n_el + grid - 1) / grid;
    epc = (epc + tile - 1) / tile * tile;
    if (epc < tile) epc = tile;
    a.el_per_cta = epc;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kLossThreads); cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fill_loss_stage2_kernel, a);
    return e == cudaSuccess ? launch_status() : (int)e;
}

extern "C" int fmlp_scale_f32(float* x, int64_t n, const float* scale_dev, fmlp_stream_t stream) {
    if (!x || !scale_dev || n < 0) return FMLP_ERR_BAD_ARG;
    if (n == 0) return FMLP_OK;
    int64_t blocks = (n + 255) / 256;
    const int sms = sm_count();
    if (sms > 0 && blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    scale_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, scale_dev);
    return launch_status();
}
