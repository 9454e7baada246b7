// tag_sim.cu — K3: pseudo-label tagging similarity, one pass over the [N, D] features.
//
// Replaces utils/local_training.py:1052-1058 + CosineSimilarityFast.forward (:1417-1435):
//     sim[c][n] = cos(f_n, P[2c]) - cos(f_n, P[2c+1]),  cos(f,p) = (f.p) * (1/(|f|*|p|))
// The reference re-reads the feature matrix four times per missing class (2 GEMV + 2 norms);
// here each row is read ONCE and |f|^2 plus all 2*M dot products are accumulated together.
//
// Mapping: the prototype vectors of the classes to score live in shared memory for the whole
// (persistent) CTA; each warp owns a tile of R consecutive rows and walks D in 128-column
// steps (one coalesced 512-B request per row per step, software-pipelined one step ahead).
// A prototype float4 read from smem feeds 4*R FMAs, which keeps the smem pipe below the FMA
// pipe; per-row partials are combined with warp shuffles.  Work per byte is (2M+1)/4 FMA, so
// for C <= 8 the kernel is HBM-bound; FMLP_SIM_FOLDED halves the FMAs for larger C.
#include "common.cuh"

namespace fmlp {

struct SimArgs {
    const float* feat;
    const float* proto;
    float* sim;
    int64_t ld_feat;
    int64_t ld_sim;
    int64_t n_total;
    int D;
    int Dpad;  // D rounded up to 128 (smem row stride, zero padded)
    int C;
    int8_t cls[FMLP_MAX_CLASSES];  // classes scored by this launch (ascending), NPAIR entries
    SegTable seg;                  // mask_a = missing-class mask per segment
};

template <int NPAIR, bool FOLD>
struct SimCfg {
    static constexpr int NV = FOLD ? NPAIR : 2 * NPAIR;
    static constexpr int R = 4;
    // > 16 accumulators per row need more than the 128 registers a 512-thread CTA allows.
    static constexpr int THREADS = (NV > 16) ? 256 : 512;
};

template <int NPAIR, bool FOLD>
__global__ void __launch_bounds__(SimCfg<NPAIR, FOLD>::THREADS, 1)
tag_sim_kernel(const __grid_constant__ SimArgs a) {
    using Cfg = SimCfg<NPAIR, FOLD>;
    constexpr int NV = Cfg::NV;
    constexpr int R = Cfg::R;
    extern __shared__ __align__(16) float smem[];
    float* sP = smem;                          // [NV][Dpad]
    float* sNorm = smem + (size_t)NV * a.Dpad; // [2*NPAIR] prototype norms (pair order)

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int D = a.D, Dpad = a.Dpad;

    // ---- prologue: prototype norms, then stage the vectors ----------------------------
    // |P_j| = sqrt(sum p^2) (torch.norm), one warp per prototype row.
    for (int j = warp; j < 2 * NPAIR; j += nwarps) {
        const int prow = 2 * (int)a.cls[j >> 1] + (j & 1);
        const float* src = a.proto + (int64_t)prow * D;
        float ss = 0.f;
        for (int d = lane; d < D; d += 32) { float v = src[d]; ss = fmaf(v, v, ss); }
        ss = warp_sum(ss);
        if (lane == 0) sNorm[j] = sqrtf(ss);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < NV * Dpad; idx += blockDim.x) {
        const int j = idx / Dpad, d = idx - j * Dpad;
        float v = 0.f;
        if (d < D) {
            if (FOLD) {
                const int c = a.cls[j];
                const float p0 = a.proto[(int64_t)(2 * c) * D + d];
                const float p1 = a.proto[(int64_t)(2 * c + 1) * D + d];
                v = __fsub_rn(__fdiv_rn(p0, sNorm[2 * j]), __fdiv_rn(p1, sNorm[2 * j + 1]));
            } else {
                const int prow = 2 * (int)a.cls[j >> 1] + (j & 1);
                v = a.proto[(int64_t)prow * D + d];
            }
        }
        sP[idx] = v;
    }
    __syncthreads();

    // ---- main loop: one tile of R rows per warp -----------------------------------------
    const int64_t n_tiles = (a.n_total + R - 1) / R;
    const int64_t tile_stride = (int64_t)gridDim.x * nwarps;
    for (int64_t t = (int64_t)blockIdx.x * nwarps + warp; t < n_tiles; t += tile_stride) {
        const int64_t row0 = t * R;
        const float* fr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int64_t row = row0 + r;
            if (row >= a.n_total) row = a.n_total - 1;  // clamp: result discarded below
            fr[r] = a.feat + row * a.ld_feat;
        }
        float acc[R][NV];
        float nrm[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            nrm[r] = 0.f;
#pragma unroll
            for (int j = 0; j < NV; ++j) acc[r][j] = 0.f;
        }
        float4 f[R], fn[R];
        int col = lane * 4;
#pragma unroll
        for (int r = 0; r < R; ++r)
            f[r] = (col < D) ? ld_stream_f4(fr[r] + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (; col < Dpad; col += 128) {
            const int ncol = col + 128;
#pragma unroll
            for (int r = 0; r < R; ++r)
                fn[r] = (ncol < D) ? ld_stream_f4(fr[r] + ncol) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const float4 p = *reinterpret_cast<const float4*>(sP + (size_t)j * Dpad + col);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    acc[r][j] = fmaf(f[r].x, p.x, acc[r][j]);
                    acc[r][j] = fmaf(f[r].y, p.y, acc[r][j]);
                    acc[r][j] = fmaf(f[r].z, p.z, acc[r][j]);
                    acc[r][j] = fmaf(f[r].w, p.w, acc[r][j]);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                nrm[r] = fmaf(f[r].x, f[r].x, nrm[r]);
                nrm[r] = fmaf(f[r].y, f[r].y, nrm[r]);
                nrm[r] = fmaf(f[r].z, f[r].z, nrm[r]);
                nrm[r] = fmaf(f[r].w, f[r].w, nrm[r]);
                f[r] = fn[r];
            }
        }
        // ---- combine the 32 lane partials ---------------------------------------------
#pragma unroll
        for (int r = 0; r < R; ++r) {
            nrm[r] = warp_sum(nrm[r]);
#pragma unroll
            for (int j = 0; j < NV; ++j) acc[r][j] = warp_sum(acc[r][j]);
        }
        // ---- epilogue: reference op order (norm product, reciprocal, multiply, subtract)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t row = row0 + r;
            if (row < a.n_total) {
                const int s = find_segment(a.seg.rows, a.seg.S, row);
                const uint32_t missing = a.seg.mask_a[s];
                const float nf = sqrtf(nrm[r]);
#pragma unroll
                for (int q = 0; q < NPAIR; ++q) {
                    const int c = a.cls[q];
                    float v;
                    if (FOLD) {
                        v = __fmul_rn(acc[r][q], __frcp_rn(nf));
                    } else {
                        const float c0 = __fmul_rn(acc[r][2 * q], __frcp_rn(__fmul_rn(nf, sNorm[2 * q])));
                        const float c1 = __fmul_rn(acc[r][2 * q + 1], __frcp_rn(__fmul_rn(nf, sNorm[2 * q + 1])));
                        v = __fsub_rn(c0, c1);
                    }
                    if (((missing >> c) & 1u) && lane == ((r * NPAIR + q) & 31))
                        a.sim[(int64_t)c * a.ld_sim + row] = v;
                }
            }
        }
    }
}

template <int NPAIR, bool FOLD>
static int launch_sim(const SimArgs& a, cudaStream_t st) {
    using Cfg = SimCfg<NPAIR, FOLD>;
    const size_t smem = ((size_t)Cfg::NV * a.Dpad + 2 * NPAIR) * sizeof(float);
    if (smem > 227u * 1024u) return FMLP_ERR_UNSUPPORTED;
    auto kern = tag_sim_kernel<NPAIR, FOLD>;
    static size_t configured = 0;  // per template instance
    if (smem > 48u * 1024u && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    const int sms = sm_count();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    const int warps = Cfg::THREADS / 32;
    const int64_t n_tiles = (a.n_total + Cfg::R - 1) / Cfg::R;
    int64_t blocks = (n_tiles + warps - 1) / warps;
    if (blocks > sms) blocks = sms;  // persistent: one CTA per SM
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, Cfg::THREADS, smem, st>>>(a);
    return launch_status();
}

template <bool FOLD>
static int dispatch_sim(int npair, const SimArgs& a, cudaStream_t st) {
    switch (npair) {
#define FMLP_SIM_CASE(N) case N: return launch_sim<N, FOLD>(a, st);
        FMLP_SIM_CASE(1) FMLP_SIM_CASE(2) FMLP_SIM_CASE(3) FMLP_SIM_CASE(4)
        FMLP_SIM_CASE(5) FMLP_SIM_CASE(6) FMLP_SIM_CASE(7) FMLP_SIM_CASE(8)
        FMLP_SIM_CASE(9) FMLP_SIM_CASE(10) FMLP_SIM_CASE(11) FMLP_SIM_CASE(12)
        FMLP_SIM_CASE(13) FMLP_SIM_CASE(14) FMLP_SIM_CASE(15) FMLP_SIM_CASE(16)
#undef FMLP_SIM_CASE
        default: return FMLP_ERR_UNSUPPORTED;
    }
}

}  // namespace fmlp

using namespace fmlp;

extern "C" int fmlp_tag_sim_f32(const float* feat, int64_t ld_feat, int D, const float* proto, int C,
                                int S, const int64_t* seg_rows, const uint32_t* seg_missing,
                                float* sim, int64_t ld_sim, int mode, fmlp_stream_t stream) {
    if (!feat || !proto || !sim || !seg_missing || C < 1 || C > FMLP_MAX_CLASSES || D < 4)
        return FMLP_ERR_BAD_ARG;
    if ((D & 3) || (ld_feat & 3) || ld_feat < D || !aligned16(feat)) return FMLP_ERR_UNSUPPORTED;
    if (mode != FMLP_SIM_PAIR && mode != FMLP_SIM_FOLDED) return FMLP_ERR_BAD_ARG;
    SimArgs a;
    int rc = fill_seg_table(a.seg, S, seg_rows, seg_missing, nullptr);
    if (rc != FMLP_OK) return rc;
    a.n_total = seg_rows[S];
    if (ld_sim < a.n_total) return FMLP_ERR_BAD_ARG;
    if (a.n_total == 0) return FMLP_OK;
    uint32_t uni = 0;
    for (int s = 0; s < S; ++s) uni |= seg_missing[s];
    if (C < 32) uni &= (1u << C) - 1u;
    int npair = 0;
    for (int c = 0; c < FMLP_MAX_CLASSES; ++c) a.cls[c] = 0;
    for (int c = 0; c < C; ++c)
        if ((uni >> c) & 1u) a.cls[npair++] = (int8_t)c;
    if (npair == 0) return FMLP_OK;
    a.feat = feat; a.proto = proto; a.sim = sim; a.ld_feat = ld_feat; a.ld_sim = ld_sim;
    a.D = D; a.Dpad = (D + 127) & ~127; a.C = C;
    cudaStream_t st = (cudaStream_t)stream;
    // More than 16 classes in one launch: split the class set (features are re-read per group).
    if (npair > 16) {
        SimArgs b = a;
        for (int base = 0; base < npair; base += 16) {
            const int n = (npair - base) < 16 ? (npair - base) : 16;
            for (int q = 0; q < 16; ++q) b.cls[q] = q < n ? a.cls[base + q] : 0;
            rc = (mode == FMLP_SIM_FOLDED) ? dispatch_sim<true>(n, b, st) : dispatch_sim<false>(n, b, st);
            if (rc != FMLP_OK) return rc;
        }
        return FMLP_OK;
    }
    return (mode == FMLP_SIM_FOLDED) ? dispatch_sim<true>(npair, a, st) : dispatch_sim<false>(npair, a, st);
}
