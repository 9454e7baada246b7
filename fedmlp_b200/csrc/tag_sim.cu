// tag_sim.cu — K3: pseudo-label tagging similarity, one pass over the [N, D] features.
//
// Replaces utils/local_training.py:1052-1058 + CosineSimilarityFast.forward (:1417-1435):
//     sim[c][n] = cos(f_n, P[2c]) - cos(f_n, P[2c+1]),  cos(f,p) = (f.p) * (1/(|f|*|p|))
// The reference re-reads the feature matrix four times per missing class (2 GEMV + 2 norms);
// here each row is read ONCE and |f|^2 plus all 2*M dot products are accumulated together.
//
// Mapping: the prototype vectors of the classes to score live in shared memory for the whole
// (persistent) CTA; each warp owns a tile of R consecutive rows and walks D in 128-column
// chunks (one coalesced 512-B request per row per chunk, ping-pong buffered one chunk ahead).
// The first version of this kernel was ISSUE-bound (ncu r01: 981 warp-instructions per row,
// only 38 % of them FMAs), so this one is built to minimise instructions:
//   * packed fp32 FMAs (fma.rn.f32x2 -> SASS FFMA2): features and prototypes are loaded as
//     64-bit pairs, every accumulator is an (even, odd) pair -> half the FMA instructions;
//   * a prototype pair read from smem feeds R rows, the chunk loop is unrolled x2 so the
//     ping-pong needs no register moves, D % 128 == 0 is a template flag (no column predicates);
//   * the R*(NV+1) per-lane partials are combined with a transposed (recursive-halving) warp
//     reduction: ~V shuffles instead of 5V, then ONE lane per (row, class) does the epilogue
//     from a per-warp smem scratch instead of all 32 lanes doing all of them redundantly.
//   * features travel global -> shared with cp.async (LDGSTS) into a private ring per warp, DEPTH
//     batches ahead of the math: B200 HBM wants ~100 KB in flight per SM (tools/exp_readpattern.cu)
//     and the register file cannot hold that next to R*(NV+1) accumulator pairs.  Every lane reads
//     back exactly the 16 bytes it copied, so the ring needs no cross-lane synchronisation.
// Work per byte is (2M+1)/4 FMA; FMLP_SIM_FOLDED halves that for large C.
#include "common.cuh"

namespace fmlp {

using u64 = unsigned long long;

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
    // rows per warp tile / threads per CTA: accumulators cost 2*R*(NV+1) registers, and the kernel
    // needs >= 12 warps per SM to hide FFMA2 / LDS latency (r01 ncu at C=14: 4 rows x 14 vectors
    // = 192 registers -> 8 warps -> 38 % issue utilisation), so R shrinks as NV grows.
#ifndef FMLP_SIM_R_MID
#define FMLP_SIM_R_MID 3
#endif
#ifndef FMLP_SIM_THREADS_MID
#define FMLP_SIM_THREADS_MID 384
#endif
    static constexpr int R = (NV <= 10) ? 4 : (NV <= 16 ? FMLP_SIM_R_MID : 2);
#ifndef FMLP_SIM_THREADS_SMALL
#define FMLP_SIM_THREADS_SMALL 384
#endif
    static constexpr int THREADS = (NV <= 10) ? FMLP_SIM_THREADS_SMALL : (NV <= 16 ? FMLP_SIM_THREADS_MID : 256);
    static constexpr int V = R * (NV + 1);                  // values reduced per tile
    static constexpr int SCRATCH = (V + 3) & ~3;
#ifndef FMLP_SIM_DEPTH
#define FMLP_SIM_DEPTH 4
#endif
    static constexpr int DEPTH = FMLP_SIM_DEPTH;            // cp.async batches in flight per warp
    static constexpr int RING = DEPTH * R * 128;            // floats per warp
};

__device__ __forceinline__ void fma2(u64& acc, u64 a, u64 b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ float pair_sum(u64 v) {
    float lo, hi;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}
// 128-bit streaming load returned as two packed fp32 pairs
__device__ __forceinline__ void ldg_pairs(const float* p, u64& a, u64& b) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}
__device__ __forceinline__ void lds_pairs(uint32_t saddr, u64& a, u64& b) {
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"(saddr));
}
// 16-byte async copy global -> shared; src_bytes = 0 zero-fills the destination.  The .ca form
// goes through L1, which merges the 32 lanes' 16-byte pieces into 128-byte line fills; with .cg
// every lane pulled its own 32-byte sector from L2 and the L2->SM traffic doubled (r01 ncu:
// l1tex__m_xbar2l1tex_read_bytes 456 MB for 225 MB of features, L2 hit rate 51 %).
#ifndef FMLP_SIM_CPASYNC
#define FMLP_SIM_CPASYNC "cp.async.ca.shared.global"
#endif
__device__ __forceinline__ void cp_async16(uint32_t saddr, const float* g, int src_bytes) {
    asm volatile(FMLP_SIM_CPASYNC " [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One stage of the transposed warp reduction: CUR values per lane -> ceil(CUR/2), lanes whose
// MASK bit is set keep the upper half.
template <int CUR, int MASK>
__device__ __forceinline__ void reduce_stage(float* v, int lane) {
    constexpr int HALF = (CUR + 1) / 2;
    const bool up = (lane & MASK) != 0;
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        const float a = v[i];
        const float b = (i + HALF < CUR) ? v[i + HALF] : 0.f;
        const float send = up ? a : b;
        const float keep = up ? b : a;
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, MASK);
    }
}
template <int V>
struct Halving {
    static constexpr int h0 = (V + 1) / 2, h1 = (h0 + 1) / 2, h2 = (h1 + 1) / 2, h3 = (h2 + 1) / 2,
                         h4 = (h3 + 1) / 2;  // values per lane after the stages with mask 16, 8, 4, 2, 1
    // original value index held in slot i of `lane` after all five stages, or -1 for padding
    __device__ static int index_of(int i, int lane) {
        int s = i;
        if (s >= h4) return -1;
        s += h4 * (lane & 1);         if (s >= h3) return -1;
        s += h3 * ((lane >> 1) & 1);  if (s >= h2) return -1;
        s += h2 * ((lane >> 2) & 1);  if (s >= h1) return -1;
        s += h1 * ((lane >> 3) & 1);  if (s >= h0) return -1;
        s += h0 * ((lane >> 4) & 1);  if (s >= V) return -1;
        return s;
    }
};

template <int NPAIR, bool FOLD, bool ALIGNED>
__global__ void __launch_bounds__(SimCfg<NPAIR, FOLD>::THREADS, 1)
tag_sim_kernel(const __grid_constant__ SimArgs a) {
    using Cfg = SimCfg<NPAIR, FOLD>;
    constexpr int NV = Cfg::NV;
    constexpr int R = Cfg::R;
    constexpr int V = Cfg::V;
    using H = Halving<V>;
    extern __shared__ __align__(16) float smem[];
    const int D = a.D, Dpad = a.Dpad;
    float* sP = smem;                                // [Dpad/128][NV][128]: LDS offsets are immediates
    float* sNorm = smem + (size_t)NV * Dpad;         // [2*NPAIR] prototype norms (pair order), padded to 4
    float* sScratch = sNorm + ((2 * NPAIR + 3) & ~3);  // [warps][SCRATCH]
    float* sRing = sScratch + (size_t)(blockDim.x >> 5) * Cfg::SCRATCH;  // [warps][DEPTH][R][128]

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;

    // ---- producer: one tile of R rows per warp, features through the cp.async ring.  The ring is
    //      primed BEFORE the prototype prologue so the first DEPTH-1 batches are in flight while the
    //      norms and the class vectors are staged (the ring is private to the warp and independent of them)
    const int nchunks = Dpad >> 7;
    constexpr int DEPTH = Cfg::DEPTH;
    const int64_t n_tiles = (a.n_total + R - 1) / R;
    const int64_t tile_stride = (int64_t)gridDim.x * nwarps;
    const int64_t t_first = (int64_t)blockIdx.x * nwarps + warp;
    const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(sRing + (size_t)warp * Cfg::RING) + lane * 16;
    // producer cursor: batch = (tile, chunk), runs DEPTH-1 batches ahead of the math.  All address
    // arithmetic is running pointers / offsets (r01 ncu: a third of the instructions of the first
    // ring version were IMAD/ISETP/SEL/LEA in this path); only the single ragged tile clamps rows.
    int64_t t_issue = t_first;
    int c_issue = 0;
    uint32_t slot_issue = 0;                                        // byte offset of the ring slot
    const float* g_tile = a.feat + t_first * R * a.ld_feat + lane * 4;  // row0 of the producer's tile
    const int64_t g_stride = tile_stride * R * a.ld_feat;
    const int64_t t_ragged = (a.n_total % R) ? n_tiles - 1 : -1;
    auto issue = [&]() {
        if (t_issue < n_tiles) {
            const int col = c_issue * 128 + lane * 4;
            const int bytes = (ALIGNED || col < D) ? 16 : 0;
            const float* g = g_tile + (bytes ? c_issue * 128 : 0);
            if (t_issue != t_ragged) {
#pragma unroll
                for (int r = 0; r < R; ++r) cp_async16(ring_addr + slot_issue + r * 512, g + r * a.ld_feat, bytes);
            } else {
                const int rmax = (int)(a.n_total - 1 - t_issue * R);  // rows past the end re-read the last row
#pragma unroll
                for (int r = 0; r < R; ++r) cp_async16(ring_addr + slot_issue + r * 512, g + (r < rmax ? r : rmax) * a.ld_feat, bytes);
            }
            if (++c_issue == nchunks) { c_issue = 0; t_issue += tile_stride; g_tile += g_stride; }
        }
        cp_async_commit();  // empty groups keep the wait_group arithmetic uniform
        slot_issue = (slot_issue + R * 512 == DEPTH * R * 512) ? 0u : slot_issue + R * 512;
    };
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) issue();

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
        sP[((d >> 7) * NV + j) * 128 + (d & 127)] = v;
    }
    __syncthreads();

    // where the values this lane ends up with after the transposed reduction belong
    int out_idx[H::h4];
#pragma unroll
    for (int i = 0; i < H::h4; ++i) out_idx[i] = H::index_of(i, lane);
    float* scratch = sScratch + warp * Cfg::SCRATCH;
    const uint32_t sP_addr = (uint32_t)__cvta_generic_to_shared(sP) + lane * 16;

    uint32_t slot = 0;  // byte offset of the consumer's ring slot

    for (int64_t t = t_first; t < n_tiles; t += tile_stride) {
        const int64_t row0 = t * R;
        u64 acc[R][NV];
        u64 nrm[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            nrm[r] = 0ull;
#pragma unroll
            for (int j = 0; j < NV; ++j) acc[r][j] = 0ull;
        }
        for (int c = 0; c < nchunks; ++c) {
            issue();
            cp_async_wait<DEPTH - 1>();  // the batch issued DEPTH-1 calls ago has landed
            u64 f0[R], f1[R];
#pragma unroll
            for (int r = 0; r < R; ++r) lds_pairs(ring_addr + slot + r * 512, f0[r], f1[r]);
            slot = (slot + R * 512 == DEPTH * R * 512) ? 0u : slot + R * 512;
            const uint32_t base = sP_addr + (uint32_t)c * (NV * 512u);
            // two prototype vectors per step and the f0 products of all of them before any f1
            // product: the two FFMA2 that hit the same accumulator are 2R-1 independent FFMA2 apart
            // (back to back they stalled on the 4-cycle FMA latency: "wait" was the top stall reason)
#pragma unroll
            for (int j = 0; j < NV; j += 2) {
                u64 p0, p1, q0 = 0ull, q1 = 0ull;
                lds_pairs(base + j * 512, p0, p1);
                if (j + 1 < NV) lds_pairs(base + (j + 1) * 512, q0, q1);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    fma2(acc[r][j], f0[r], p0);
                    if (j + 1 < NV) fma2(acc[r][j + 1], f0[r], q0);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    fma2(acc[r][j], f1[r], p1);
                    if (j + 1 < NV) fma2(acc[r][j + 1], f1[r], q1);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) fma2(nrm[r], f0[r], f0[r]);
#pragma unroll
            for (int r = 0; r < R; ++r) fma2(nrm[r], f1[r], f1[r]);
        }

        // ---- combine the 32 lane partials (transposed reduction) ------------------------
        float v[V];
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int j = 0; j < NV; ++j) v[r * (NV + 1) + j] = pair_sum(acc[r][j]);
            v[r * (NV + 1) + NV] = pair_sum(nrm[r]);
        }
        reduce_stage<V, 16>(v, lane);
        reduce_stage<H::h0, 8>(v, lane);
        reduce_stage<H::h1, 4>(v, lane);
        reduce_stage<H::h2, 2>(v, lane);
        reduce_stage<H::h3, 1>(v, lane);
        __syncwarp();  // previous tile's epilogue reads are done
#pragma unroll
        for (int i = 0; i < H::h4; ++i)
            if (out_idx[i] >= 0) scratch[out_idx[i]] = v[i];
        __syncwarp();

        // ---- epilogue: one lane per (row, class); reference op order (norm product,
        //      reciprocal, multiply, subtract)
        for (int e = lane; e < R * NPAIR; e += 32) {
            const int r = e / NPAIR, q = e - r * NPAIR;
            const int64_t row = row0 + r;
            if (row < a.n_total) {
                const int c = a.cls[q];
                const int s = find_segment(a.seg.rows, a.seg.S, row);
                if ((a.seg.mask_a[s] >> c) & 1u) {
                    const float* rv = scratch + r * (NV + 1);
                    const float nf = sqrtf(rv[NV]);
                    float out;
                    if (FOLD) {
                        out = __fmul_rn(rv[q], __frcp_rn(nf));
                    } else {
                        const float c0 = __fmul_rn(rv[2 * q], __frcp_rn(__fmul_rn(nf, sNorm[2 * q])));
                        const float c1 = __fmul_rn(rv[2 * q + 1], __frcp_rn(__fmul_rn(nf, sNorm[2 * q + 1])));
                        out = __fsub_rn(c0, c1);
                    }
                    a.sim[(int64_t)c * a.ld_sim + row] = out;
                }
            }
        }
    }
}

template <int NPAIR, bool FOLD, bool ALIGNED>
static int launch_sim_inst(const SimArgs& a, cudaStream_t st) {
    using Cfg = SimCfg<NPAIR, FOLD>;
    const int warps = Cfg::THREADS / 32;
    const size_t smem = ((size_t)Cfg::NV * a.Dpad + ((2 * NPAIR + 3) & ~3) + (size_t)warps * (Cfg::SCRATCH + Cfg::RING)) * sizeof(float);
    if (smem > 227u * 1024u) return FMLP_ERR_UNSUPPORTED;
    auto kern = tag_sim_kernel<NPAIR, FOLD, ALIGNED>;
    static size_t configured = 0;  // per template instance
    if (smem > 48u * 1024u && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    const int sms = sm_count();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    const int64_t n_tiles = (a.n_total + Cfg::R - 1) / Cfg::R;
    int64_t blocks = (n_tiles + warps - 1) / warps;
    if (blocks > sms) blocks = sms;  // persistent: one CTA per SM
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, Cfg::THREADS, smem, st>>>(a);
    return launch_status();
}

template <int NPAIR, bool FOLD>
static int launch_sim(const SimArgs& a, cudaStream_t st) {
    return (a.D % 128 == 0) ? launch_sim_inst<NPAIR, FOLD, true>(a, st) : launch_sim_inst<NPAIR, FOLD, false>(a, st);
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
