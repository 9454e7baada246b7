// tag_sim.cu — K3: pseudo-label tagging similarity, one pass over the [N, D] features.
//
// Replaces utils/local_training.py:1052-1058 + CosineSimilarityFast.forward (:1417-1435):
//     sim[c][n] = cos(f_n, P[2c]) - cos(f_n, P[2c+1]),  cos(f,p) = (f.p) * (1/(|f|*|p|))
// The reference re-reads the feature matrix four times per missing class (2 GEMV + 2 norms);
// here each row is read ONCE and |f|^2 plus all dot products are accumulated together.
//
// Round-2 design (the round-1 kernel staged 16-byte cp.async pieces per lane, split every row over
// the 32 lanes of a warp and paid a cross-lane reduction per tile; ncu showed it issue/latency-bound
// at 128 registers and 12 warps per SM, 0.66 of the HBM roof at C=5 and 0.41 at C=14):
//   * THREAD PER ROW.  A tile is ROWS = 32*RT consecutive rows; lane l of every compute warp owns
//     rows l, l+32, ... of the tile and keeps their RT*(NV+1) accumulators (packed fp32 pairs,
//     fma.rn.f32x2 -> SASS FFMA2) in registers.  The class vectors are therefore the SAME address for
//     all lanes: one broadcast LDS.128 (a single shared-memory wavefront) feeds 2*RT FFMA2, where the
//     round-1 mapping needed a full 512-byte read per 2*R FFMA2.  No cross-lane reduction at all.
//   * TMA staging.  Features travel global -> shared as 2-D tiled bulk tensor copies
//     (cp.async.bulk.tensor.2d, SASS UTMALDG): one box = ROWS rows x 32 columns (128 bytes per row) in
//     the 128-byte swizzle, so lanes reading "their" row at the same column hit 8 distinct 16-byte
//     bank groups per quarter-warp: conflict-free without padding.  Each compute warp owns a private
//     ring of S boxes and its lane 0 re-arms the mbarrier and issues the next box the moment the warp
//     has consumed one: no registers are spent on loads in flight, no warp ever waits for another
//     warp's data, and S-1 boxes per warp (64..160 KB per SM) are always in flight.
//   * The W compute warps of a CTA split the COLUMNS of a tile (column group g of 32 goes to warp
//     g % W: together they pull 128*W contiguous bytes of every row), so the unit of load balance is a
//     tile per CTA — 860 tiles over 148 SMs at the bench shape — not a tile per warp.  Per tile the
//     warps' partial sums meet in shared memory (fixed warp order: deterministic) and the 32*W threads
//     share the epilogue in the reference's op order (norm product, reciprocal, multiply, subtract).
//   * The class vectors (pair mode: the prototypes; folded: q_c, with its IEEE divides) are built ONCE
//     by a small table kernel into the caller's workspace, in the [column quad][vector] order the inner
//     loop reads, and every CTA of the main kernel pulls them with one bulk copy.  The main kernel is
//     a programmatic dependent launch: it initialises its barriers and starts streaming features while
//     the table kernel is still running and only waits (griddepcontrol.wait) before the table copy.
//     (First round-2 version: every CTA rebuilt the table in its prologue — 32..130 dependent L2 round
//     trips per thread, 10..30 us in front of a 45..90 us kernel.)
//   * The per-tile epilogue (add the W warps' partial sums, norms, reciprocal, segment lookup, store) belongs to
//     TWO DEDICATED WARPS, one lane per row of the tile: the compute warps hand their partial sums over through
//     shared memory and a full/empty mbarrier pair and go straight on with the next tile.  (r02 ncu source view at 14
//     class vectors: 26 % of the kernel's samples sat in the epilogue, executed by all 8 compute warps between two
//     block-wide barriers while no FMA was issued.)
// Work per byte is (NV+1)/4 FMA with NV = 2M (pair mode) or M (FMLP_SIM_FOLDED: one dot against
// q_c = P0/|P0| - P1/|P1|).  fp32 FFMA2 only: TF32/BF16 tensor cores would flip signs at 1e-3.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace fmlp {

using u64 = unsigned long long;

struct SimArgs {
    const float* proto;
    const float* table;   // workspace: [NG*8 column quads][NV] float4 class vectors, then [2*NPAIR] prototype norms
    float* sim;
    int64_t ld_sim;
    int64_t n_total;
    int D;
    int NG;       // column groups of 32 (D rounded up)
    int S;        // ring stages per warp
    int C;
    int8_t cls[FMLP_MAX_CLASSES];  // classes scored by this launch (ascending), NPAIR entries
    SegTable seg;                  // mask_a = missing-class mask per segment
};

#ifndef FMLP_SIM_W
#define FMLP_SIM_W 8
#endif
#ifndef FMLP_SIM_RT_SMALL
#define FMLP_SIM_RT_SMALL 2
#endif
#ifndef FMLP_SIM_RT_LARGE
#define FMLP_SIM_RT_LARGE 2
#endif
#ifndef FMLP_SIM_TABLE_FIRST
#define FMLP_SIM_TABLE_FIRST 1
#endif
// warps that do nothing but the per-tile epilogue (see the kernel); 2 x 32 lanes = one lane per row of a 64-row tile
#ifndef FMLP_SIM_EPI_WARPS
#define FMLP_SIM_EPI_WARPS 2
#endif

template <int NPAIR, bool FOLD>
struct SimCfg {
    static constexpr int NV = FOLD ? NPAIR : 2 * NPAIR;
    // compute warps per CTA; wide pair-mode launches (NV > 16) halve it: their class-vector table and
    // partial-sum staging would not leave room for the rings otherwise
    static constexpr int W = (NV > 16 && FMLP_SIM_W > 4) ? 4 : FMLP_SIM_W;
    static constexpr int RT = (NV <= 8) ? FMLP_SIM_RT_SMALL : FMLP_SIM_RT_LARGE;  // rows per thread
    static constexpr int ROWS = 32 * RT;                                  // rows per tile
    static constexpr int EW = FMLP_SIM_EPI_WARPS;                         // epilogue warps
    static constexpr int THREADS = 32 * (W + EW);
    static constexpr int STAGE_BYTES = ROWS * 128;                        // one box: ROWS x 32 floats
    static constexpr int PV = NV + 1;                                     // partial values per row
};

#ifdef FMLP_SIM_TRACE
// tuning builds only (tools/exp_sim.cu): per-CTA time stamps of the last launch, [CTA][8] =
// {globaltimer at entry, table ready, first tile done, loop done, clock64 at the same four points}
__device__ unsigned long long g_sim_trace[160 * 8 + 8];
__device__ __forceinline__ unsigned long long sim_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define SIM_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x < 160) { g_sim_trace[blockIdx.x * 8 + (i)] = sim_gtime(); g_sim_trace[blockIdx.x * 8 + 4 + (i)] = clock64(); } } while (0)
#else
#define SIM_STAMP(i) do {} while (0)
#endif

__device__ __forceinline__ void fma2(u64& acc, u64 a, u64 b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ float pair_sum(u64 v) {
    float lo, hi;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}
__device__ __forceinline__ uint32_t sim_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sim_mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void sim_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sim_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool sim_mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void sim_mbar_wait(uint32_t bar, uint32_t parity) {
    while (!sim_mbar_try(bar, parity)) {}
}
// 2-D tiled TMA load: box = [ROWS rows][32 columns], 128-byte swizzle; OOB rows / columns read as 0
__device__ __forceinline__ void sim_tma_load(uint32_t dst, const CUtensorMap* tm, int col, int row, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tm), "r"(col), "r"(row), "r"(bar) : "memory");
}

// global -> shared bulk copy (SASS UBLKCP); completion is counted in bytes on `bar`
__device__ __forceinline__ void sim_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- class-vector table (one CTA per scored class) -------------------------------------------
// |P[2c]|, |P[2c+1]| (torch.norm: sqrt of the fp32 sum of squares), then the vectors the main kernel
// multiplies with, element (d, j) at table[((d >> 2) * NV + j) * 4 + (d & 3)], columns D..NG*32 zero.
__global__ void __launch_bounds__(256) sim_quad_table_kernel(const float* __restrict__ proto, int D, int NG, int fold,
                                                             int npair, const __grid_constant__ SimArgs a,
                                                             float* __restrict__ table) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the main kernel may start its prologue now
    // This kernel is itself launched as a programmatic dependent of whatever precedes it in the stream (its CTAs are
    // resident and waiting when the predecessor drains: the 3-6 us launch chain in front of the main kernel shrinks to
    // the table's own run time).  Everything it reads may be the predecessor's output, hence the wait first; the main
    // kernel touches global memory only behind its own wait, which covers this kernel and, transitively, that one.
    asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef FMLP_SIM_TRACE
    if (threadIdx.x == 0 && blockIdx.x == 0) g_sim_trace[160 * 8] = sim_gtime();
#endif
    __shared__ float red[2][8];
    __shared__ float nrm[2];
    const int q = blockIdx.x;
    const int c = a.cls[q];
    const float* p0 = proto + (int64_t)(2 * c) * D;
    const float* p1 = p0 + D;
    float s0 = 0.f, s1 = 0.f;
    for (int d = threadIdx.x; d < D; d += 256) { const float u = p0[d], v = p1[d]; s0 = fmaf(u, u, s0); s1 = fmaf(v, v, s1); }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        nrm[threadIdx.x] = sqrtf(s);
    }
    __syncthreads();
    const int nv = fold ? npair : 2 * npair;
    for (int d = threadIdx.x; d < NG * 32; d += 256) {
        const float u = d < D ? p0[d] : 0.f, v = d < D ? p1[d] : 0.f;
        float* dst = table + ((size_t)(d >> 2) * nv) * 4 + (d & 3);
        if (fold) {
            dst[q * 4] = d < D ? __fsub_rn(__fdiv_rn(u, nrm[0]), __fdiv_rn(v, nrm[1])) : 0.f;
        } else {
            dst[(2 * q) * 4] = u;
            dst[(2 * q + 1) * 4] = v;
        }
    }
    if (threadIdx.x < 2) table[(size_t)nv * NG * 32 + 2 * q + threadIdx.x] = nrm[threadIdx.x];
#ifdef FMLP_SIM_TRACE
    __syncthreads();
    if (threadIdx.x == 0) g_sim_trace[160 * 8 + 1 + (blockIdx.x == 0 ? 0 : 1)] = sim_gtime();
#endif
}

template <int NPAIR, bool FOLD>
__global__ void __launch_bounds__(SimCfg<NPAIR, FOLD>::THREADS, 1)
tag_sim_kernel(const __grid_constant__ SimArgs a, const __grid_constant__ CUtensorMap tmap) {
    using Cfg = SimCfg<NPAIR, FOLD>;
    constexpr int NV = Cfg::NV, RT = Cfg::RT, W = Cfg::W, ROWS = Cfg::ROWS, PV = Cfg::PV;
    constexpr int STAGE = Cfg::STAGE_BYTES;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int S = a.S, NG = a.NG;
    // [W][S] boxes (1024-byte aligned: the swizzle pattern is a function of the address bits)
    unsigned char* sRing = smem_raw;
    float* sP = reinterpret_cast<float*>(sRing + (size_t)W * S * STAGE);   // [NG*8 column quads][NV] float4
    float* sPart = sP + (size_t)NG * 32 * NV;                              // [W][PV][ROWS] partial sums
    float* sNorm = sPart + (size_t)W * PV * ROWS;                          // [2*NPAIR] prototype norms, padded to 4
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sNorm + ((2 * NPAIR + 3) & ~3));   // [W][S] ring barriers, table, full, empty

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    SIM_STAMP(0);
    // the selection kernel behind this one is a programmatic dependent launch too: let it be set up now
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int64_t n_tiles = (a.n_total + ROWS - 1) / ROWS;
    const bool is_epi = warp >= W;                                         // epilogue warps: W .. W+EW-1
    const int my_groups = (!is_epi && NG > warp) ? (NG - warp + W - 1) / W : 0;   // column groups warp, warp+W, ...
    const uint32_t ring_u32 = sim_smem_u32(sRing + (size_t)(is_epi ? 0 : warp) * S * STAGE);
    const uint32_t bar_u32 = sim_smem_u32(sBar + (is_epi ? 0 : warp) * S);

    // ---- producer state (lane 0 of each warp): next box = (tile, group) in consumption order
    int64_t p_tile = blockIdx.x;
    int p_gi = 0;
    int p_slot = 0;
    auto issue = [&]() {
        if (p_tile < n_tiles && my_groups > 0) {
            const uint32_t bar = bar_u32 + p_slot * 8;
            sim_mbar_expect_tx(bar, STAGE);                                // OOB parts of a box count as filled
            sim_tma_load(ring_u32 + p_slot * STAGE, &tmap, (warp + p_gi * W) * 32, (int)(p_tile * ROWS), bar);
            if (++p_gi == my_groups) { p_gi = 0; p_tile += gridDim.x; }
        }
        p_slot = (p_slot + 1 == S) ? 0 : p_slot + 1;
    };
    const uint32_t tbar_u32 = sim_smem_u32(sBar + W * S);
    const uint32_t full_u32 = tbar_u32 + 8, empty_u32 = tbar_u32 + 16;    // partial sums ready / consumed
    if (lane == 0 && !is_epi) {
        for (int s = 0; s < S; ++s) sim_mbar_init(bar_u32 + s * 8, 1);
        if (warp == 0) {
            sim_mbar_init(tbar_u32, 1);
            sim_mbar_init(full_u32, W);                 // lane 0 of every compute warp
            sim_mbar_init(empty_u32, Cfg::EW);          // lane 0 of every epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    auto load_table = [&]() {
        // the class vectors come from the table kernel this launch programmatically depends on
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const uint32_t bytes = (uint32_t)((NG * 32 * NV + ((2 * NPAIR + 3) & ~3)) * sizeof(float));
        sim_mbar_expect_tx(tbar_u32, bytes);
        // sP, sPart, sNorm are laid out so that table = [sP | norms] lands with two copies
        sim_bulk_g2s(sim_smem_u32(sP), a.table, (uint32_t)(NG * 32 * NV * sizeof(float)), tbar_u32);
        sim_bulk_g2s(sim_smem_u32(sNorm), a.table + (size_t)NG * 32 * NV, (uint32_t)(((2 * NPAIR + 3) & ~3) * sizeof(float)), tbar_u32);
    };
#if FMLP_SIM_TABLE_FIRST
    // The table copy goes to the TMA unit AHEAD of the feature boxes: queued behind W*S boxes of cold HBM reads it
    // landed 3.3 us after kernel entry (r02 trace), although the table kernel had long finished — by the time
    // this grid's CTAs start, griddepcontrol.wait returns at once, so nothing is lost by waiting first.
    if (threadIdx.x == 0) load_table();
    __syncthreads();
    if (lane == 0 && !is_epi)
        for (int s = 0; s < S; ++s) issue();
#else
    if (lane == 0 && !is_epi) {
        for (int s = 0; s < S; ++s) issue();     // the ring fills while the table kernel finishes
        if (warp == 0) load_table();
    }
#endif
    __syncthreads();                              // barrier inits are visible to every waiter
    sim_mbar_wait(tbar_u32, 0);
    SIM_STAMP(1);

    if (is_epi) {
        // ---- epilogue warps: lane <-> row of the tile; reference op order (norm product, reciprocal, multiply, subtract)
        const int et = threadIdx.x - 32 * W;
        uint32_t ph = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            sim_mbar_wait(full_u32, ph);            // the W compute warps have stored this tile's partial sums
            ph ^= 1u;
            const int64_t row0 = tile * ROWS;
            for (int r = et; r < ROWS; r += 32 * Cfg::EW) {
                const int64_t row = row0 + r;
                if (row >= a.n_total) continue;
                const uint32_t mask = a.seg.mask_a[find_segment(a.seg.rows, a.seg.S, row)];
                float ff = 0.f;
#pragma unroll
                for (int w = 0; w < W; ++w) ff += sPart[((size_t)w * PV + NV) * ROWS + r];     // fixed warp order: deterministic
                const float nf = sqrtf(ff);
                const float rnf = __frcp_rn(nf);
#pragma unroll
                for (int q = 0; q < NPAIR; ++q) {
                    const int c = a.cls[q];
                    if (!((mask >> c) & 1u)) continue;
                    float d0 = 0.f, d1 = 0.f;
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        const float* part = sPart + (size_t)w * PV * ROWS + r;
                        if (FOLD) {
                            d0 += part[q * ROWS];
                        } else {
                            d0 += part[(2 * q) * ROWS];
                            d1 += part[(2 * q + 1) * ROWS];
                        }
                    }
                    float out;
                    if (FOLD) {
                        out = __fmul_rn(d0, rnf);
                    } else {
                        const float c0 = __fmul_rn(d0, __frcp_rn(__fmul_rn(nf, sNorm[2 * q])));
                        const float c1 = __fmul_rn(d1, __frcp_rn(__fmul_rn(nf, sNorm[2 * q + 1])));
                        out = __fsub_rn(c0, c1);
                    }
                    a.sim[(int64_t)c * a.ld_sim + row] = out;
                }
            }
            __syncwarp();
            if (lane == 0) sim_mbar_arrive(empty_u32);      // the staging area may be overwritten
        }
        return;
    }

    // lane's row inside a box: 128 bytes per row, 16-byte chunk k of row r sits at chunk k ^ (r & 7)
    const uint32_t row_off = (uint32_t)lane * 128u + ((uint32_t)(lane & 7) << 4);
    int slot = 0;
    uint32_t parity = 0, eph = 0;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        u64 acc[RT][PV];
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int j = 0; j < PV; ++j) acc[r][j] = 0ull;

        for (int gi = 0; gi < my_groups; ++gi) {
            sim_mbar_wait(bar_u32 + slot * 8, parity);
            const unsigned char* box = sRing + ((size_t)warp * S + slot) * STAGE;
            const ulonglong2* pv = reinterpret_cast<const ulonglong2*>(sP) + (size_t)(warp + gi * W) * 8 * NV;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                ulonglong2 f[RT];
#pragma unroll
                for (int r = 0; r < RT; ++r)
                    f[r] = *reinterpret_cast<const ulonglong2*>(box + ((row_off ^ (uint32_t)(k << 4)) + r * 4096));
                // vectors in groups of JB: all .x products of a group before any .y product, so the two
                // FFMA2 that hit one accumulator are RT*JB - 1 independent FFMA2 apart (4-cycle FMA latency)
                // and only JB class-vector quads are live at a time
                constexpr int JB = 4;
#pragma unroll
                for (int j0 = 0; j0 < NV; j0 += JB) {
                    ulonglong2 p[JB];
#pragma unroll
                    for (int j = 0; j < JB; ++j)
                        if (j0 + j < NV) p[j] = pv[k * NV + j0 + j];
#pragma unroll
                    for (int j = 0; j < JB; ++j)
                        if (j0 + j < NV)
#pragma unroll
                            for (int r = 0; r < RT; ++r) fma2(acc[r][j0 + j], f[r].x, p[j].x);
#pragma unroll
                    for (int j = 0; j < JB; ++j)
                        if (j0 + j < NV)
#pragma unroll
                            for (int r = 0; r < RT; ++r) fma2(acc[r][j0 + j], f[r].y, p[j].y);
                }
#pragma unroll
                for (int r = 0; r < RT; ++r) fma2(acc[r][NV], f[r].x, f[r].x);
#pragma unroll
                for (int r = 0; r < RT; ++r) fma2(acc[r][NV], f[r].y, f[r].y);
            }
            __syncwarp();                       // every lane is done reading the box
            if (lane == 0) issue();             // refill it (same slot: the producer cursor trails by S)
            if (++slot == S) { slot = 0; parity ^= 1u; }
        }

        // ---- hand the partial sums to the epilogue warps and go on with the next tile -------------
        if (tile != (int64_t)blockIdx.x) { sim_mbar_wait(empty_u32, eph); eph ^= 1u; }   // previous tile's sums consumed
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int j = 0; j < PV; ++j) sPart[(warp * PV + j) * ROWS + r * 32 + lane] = pair_sum(acc[r][j]);
        __syncwarp();
        if (lane == 0) sim_mbar_arrive(full_u32);
#ifdef FMLP_SIM_TRACE
        if (tile == blockIdx.x) SIM_STAMP(2);
#endif
    }
    SIM_STAMP(3);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*SimTensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static SimTensorMapEncodeFn sim_tensor_map_encoder() {
    static SimTensorMapEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<SimTensorMapEncodeFn>(p);
    }
    return fn;
}

static bool sim_use_pdl() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FMLP_SIM_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

static size_t sim_table_floats(int nv, int npair, int NG) { return (size_t)NG * 32 * nv + ((2 * npair + 3) & ~3); }

// Shared memory of a launch that scores `npair` classes with two ring stages (mirrors SimCfg / launch_sim).
static size_t sim_smem_min(int npair, bool fold, int NG) {
    const int nv = fold ? npair : 2 * npair;
    const int W = (nv > 16 && FMLP_SIM_W > 4) ? 4 : FMLP_SIM_W;
    const int RT = (nv <= 8) ? FMLP_SIM_RT_SMALL : FMLP_SIM_RT_LARGE;
    const size_t rows = 32 * RT;
    const size_t fixed = ((size_t)NG * 32 * nv + (size_t)W * (nv + 1) * rows + ((2 * npair + 3) & ~3)) * sizeof(float);
    return (size_t)W * 2 * rows * 128 + fixed + ((size_t)W * 2 + 4) * sizeof(uint64_t);
}

struct SimFeat {
    const float* feat;
    int64_t ld_feat;
};

template <int NPAIR, bool FOLD>
static int launch_sim(SimArgs& a, const SimFeat& ft, cudaStream_t st) {
    using Cfg = SimCfg<NPAIR, FOLD>;
    const size_t fixed = ((size_t)a.NG * 32 * Cfg::NV + (size_t)Cfg::W * Cfg::PV * Cfg::ROWS + ((2 * NPAIR + 3) & ~3)) * sizeof(float);
    // Ring depth: as many stages as fit in `soft` — 200 KB, not the full 227 KB, so that a CTA of a concurrent
    // kernel with a little shared memory (the aggregation kernels of the other streams) can still share the SM;
    // a launch whose fixed part alone needs more falls back to two stages within the hard limit.
    const size_t hard = 227u * 1024u;
    static int max_stages = 0;
    if (max_stages == 0) {
        max_stages = 8;
        if (const char* e = getenv("FMLP_SIM_STAGES")) { int v = atoi(e); if (v >= 2 && v <= 16) max_stages = v; }
    }
    const size_t soft = (size_t)tuning_value(FMLP_TUNE_SIM_SMEM_BUDGET_KB, "FMLP_SIM_SMEM_KB", 64, 227, 200) * 1024u;
    size_t budget = soft;
    int S = max_stages;
    auto smem_of = [&](int s) { return (size_t)Cfg::W * s * Cfg::STAGE_BYTES + fixed + ((size_t)Cfg::W * s + 4) * sizeof(uint64_t); };   // ring barriers + table / full / empty
    while (S > 2 && smem_of(S) > budget) --S;
    size_t smem = smem_of(S);
    if (smem > hard) return FMLP_ERR_UNSUPPORTED;
    // tuning knob: request (not use) this many KB, so that no CTA of a concurrent kernel shares the SM
    const int request_kb = tuning_value(FMLP_TUNE_SIM_REQUEST_SMEM_KB, "FMLP_SIM_REQUEST_SMEM_KB", 0, 227, 0);
    if ((size_t)request_kb * 1024u > smem) smem = (size_t)request_kb * 1024u;
    a.S = S;
    auto kern = tag_sim_kernel<NPAIR, FOLD>;
    static size_t configured = 0;  // per template instance
    if (smem > 48u * 1024u && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    SimTensorMapEncodeFn enc = sim_tensor_map_encoder();
    if (!enc) return FMLP_ERR_UNSUPPORTED;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    cuuint64_t dims[2] = {(cuuint64_t)a.D, (cuuint64_t)a.n_total};
    cuuint64_t strides[1] = {(cuuint64_t)ft.ld_feat * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)Cfg::ROWS};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ft.feat), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return FMLP_ERR_UNSUPPORTED;
    const int sms = sm_count();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    {
        cudaLaunchConfig_t tcfg = {};
        cudaLaunchAttribute tattr[1];
        tattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        tattr[0].val.programmaticStreamSerializationAllowed = 1;
        tcfg.attrs = tattr; tcfg.numAttrs = sim_use_pdl() ? 1 : 0;
        tcfg.gridDim = dim3(NPAIR); tcfg.blockDim = dim3(256); tcfg.dynamicSmemBytes = 0; tcfg.stream = st;
        cudaError_t te = cudaLaunchKernelEx(&tcfg, sim_quad_table_kernel, a.proto, a.D, a.NG, FOLD ? 1 : 0, NPAIR, a,
                                            const_cast<float*>(a.table));
        int rc = te == cudaSuccess ? launch_status() : (int)te;
        if (rc != FMLP_OK) return rc;
    }
    const int64_t n_tiles = (a.n_total + Cfg::ROWS - 1) / Cfg::ROWS;
    const int64_t blocks = n_tiles < sms ? n_tiles : sms;     // persistent: one CTA per SM, tiles round-robin
    // programmatic dependent launch: the main kernel's barrier setup and first feature boxes overlap
    // the table kernel; it waits (griddepcontrol.wait) right before it copies the table
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = sim_use_pdl() ? 1 : 0;
    cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, tmap);
    return e == cudaSuccess ? launch_status() : (int)e;
}

template <bool FOLD>
static int dispatch_sim(int npair, SimArgs& a, const SimFeat& ft, cudaStream_t st) {
    switch (npair) {
#define FMLP_SIM_CASE(N) case N: return launch_sim<N, FOLD>(a, ft, st);
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

#ifdef FMLP_SIM_TRACE
extern "C" int fmlp_sim_trace_read(unsigned long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g_sim_trace, sizeof(unsigned long long) * (size_t)n);
}
#endif

extern "C" size_t fmlp_tag_sim_ws_bytes(int C, int D) {
    if (C < 1 || C > FMLP_MAX_CLASSES || D < 1) return 0;
    const int n = C < 16 ? C : 16;            // at most 16 classes per launch, pair mode = 2 vectors each
    return (sim_table_floats(2 * n, n, (D + 31) >> 5) * sizeof(float) + 15) & ~(size_t)15;
}

extern "C" int fmlp_tag_sim_f32(const float* feat, int64_t ld_feat, int D, const float* proto, int C,
                                int S, const int64_t* seg_rows, const uint32_t* seg_missing,
                                float* sim, int64_t ld_sim, int mode, void* ws, size_t ws_bytes,
                                fmlp_stream_t stream) {
    if (!feat || !proto || !sim || !seg_missing || C < 1 || C > FMLP_MAX_CLASSES || D < 4)
        return FMLP_ERR_BAD_ARG;
    if (!ws || !aligned16(ws)) return FMLP_ERR_BAD_ARG;
    if (ws_bytes < fmlp_tag_sim_ws_bytes(C, D)) return FMLP_ERR_WORKSPACE;
    if ((D & 3) || (ld_feat & 3) || ld_feat < D || !aligned16(feat)) return FMLP_ERR_UNSUPPORTED;
    if (mode != FMLP_SIM_PAIR && mode != FMLP_SIM_FOLDED) return FMLP_ERR_BAD_ARG;
    SimArgs a;
    int rc = fill_seg_table(a.seg, S, seg_rows, seg_missing, nullptr);
    if (rc != FMLP_OK) return rc;
    a.n_total = seg_rows[S];
    if (ld_sim < a.n_total) return FMLP_ERR_BAD_ARG;
    if (a.n_total == 0) return FMLP_OK;
    if (a.n_total > 0x7fffffffll) return FMLP_ERR_UNSUPPORTED;    // TMA row coordinates are int32
    uint32_t uni = 0;
    for (int s = 0; s < S; ++s) uni |= seg_missing[s];
    if (C < 32) uni &= (1u << C) - 1u;
    int npair = 0;
    for (int c = 0; c < FMLP_MAX_CLASSES; ++c) a.cls[c] = 0;
    for (int c = 0; c < C; ++c)
        if ((uni >> c) & 1u) a.cls[npair++] = (int8_t)c;
    if (npair == 0) return FMLP_OK;
    a.proto = proto; a.table = static_cast<const float*>(ws); a.sim = sim; a.ld_sim = ld_sim;
    a.D = D; a.NG = (D + 31) >> 5; a.C = C; a.S = 0;
    const SimFeat ft = {feat, ld_feat};
    cudaStream_t st = (cudaStream_t)stream;
    // Classes per launch: at most 16, and few enough for the class-vector table + partial staging + two ring
    // stages to fit in shared memory (wide D in pair mode).  A larger class set is split across launches;
    // the features are then re-read per group.
    const bool fold = mode == FMLP_SIM_FOLDED;
    int group = npair < 16 ? npair : 16;
    while (group > 1 && sim_smem_min(group, fold, a.NG) > 227u * 1024u) --group;
    if (sim_smem_min(group, fold, a.NG) > 227u * 1024u) return FMLP_ERR_UNSUPPORTED;
    if (npair > group) {
        SimArgs b = a;
        for (int base = 0; base < npair; base += group) {
            const int n = (npair - base) < group ? (npair - base) : group;
            for (int q = 0; q < FMLP_MAX_CLASSES; ++q) b.cls[q] = q < n ? a.cls[base + q] : 0;
            rc = fold ? dispatch_sim<true>(n, b, ft, st) : dispatch_sim<false>(n, b, ft, st);
            if (rc != FMLP_OK) return rc;
        }
        return FMLP_OK;
    }
    return (mode == FMLP_SIM_FOLDED) ? dispatch_sim<true>(npair, a, ft, st) : dispatch_sim<false>(npair, a, ft, st);
}
