// pool_tag.cu — SURVEY §8(f).1: tagging fused into feature extraction.
//
// The reference's tagging pass runs the backbone batch by batch, keeps only the pooled feature of
// every sample and concatenates them into f [N, D] (utils/local_training.py:1026-1049; the
// quadratic torch.cat), then scores f against the prototypes (:1052-1058).  The pooled feature
// itself is the tail of the (patched torchvision / efficientnet) model forward
//     feature = flatten(adaptive_avg_pool2d(relu(fmap), 1))      fmap [B, D, H, W]
// This kernel is that tail AND the scoring in one pass over the last feature map of a batch:
//     feat[b, d]  = (1/HW) * sum_hw act(fmap[b, d, hw])                  (written for the classifier)
//     sim[c][b]   = cos(feat_b, P[2c]) - cos(feat_b, P[2c+1])            (K3's epilogue, per batch)
// so f [N, D] is never materialised or re-read; sims land directly in column `b` of the class-major
// [C, N] matrix that fmlp_tag_select consumes.
//
// Mapping (B200): a thread-block CLUSTER owns one sample; its G CTAs split the D channels into G
// contiguous ranges, so a batch of 128 samples x 4 ranges is 512 CTAs and even a batch of 32 fills
// the machine.  Every CTA streams its [channels x HW] slab through a ring of shared-memory stages
// filled by the TMA engine (NCHW: one cp.async.bulk per stage; channels_last: one 3-D tiled
// cp.async.bulk.tensor box per stage; mbarrier complete_tx, no registers spent on the loads).  Four
// compute warps own one channel per thread: 49 shared-memory loads issued back to back (stride HW
// floats: conflict-free for odd HW), a tree sum, the divide, the feature store and nv branch-free FMAs
// against the class vectors of the CTA's channel range (staged once per CTA).  A fifth warp is the
// epilogue warp: it combines the four warps' partial dots, meets the other CTAs on the cluster
// barrier and, in rank 0, sums the G partials in rank order through distributed shared memory and
// writes the similarities — no atomics, no workspace, deterministic.  The compute warps only ever wait
// on the cluster barrier of the PREVIOUS sample, so they never stall on the epilogue.
// Four CTAs per SM with two 25 KB stages each beat fewer CTAs with deeper rings (r01 sweep: per-CTA
// consumption latency, not bytes in flight, was the limit).  HBM-bound: 4*B*D*HW bytes in, 4*B*D out;
// measured 5.9 TB/s (C=5, D=1024) / 5.6 TB/s (C=14, D=1280) at B=2048, 6.2 TB/s pooling only.
#include <cooperative_groups.h>
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace fmlp {

constexpr int kPoolThreads = 128;   // one thread per channel of a stage
constexpr int kPoolMaxCluster = 8;  // portable cluster size
constexpr int kPoolMaxVec = 32;     // class vectors per launch
constexpr int kPoolMaxStages = 8;

struct PoolArgs {
    const float* fmap;
    const float* table;   // [nv][D] class vectors, then [2*npair] prototype norms (pair mode)
    float* feat;          // [B, ld_feat] or nullptr
    float* sim;           // [C, ld_sim] (already offset to the batch's first column) or nullptr
    int64_t ld_feat, ld_sim;
    int B, D, HW;
    int Dr;               // channels per cluster rank (multiple of 4)
    int CH;               // channels per ring stage (<= kPoolThreads, multiple of 4)
    int nstage;
    int relu, nhwc, fold, npair, nv;
    int8_t cls[FMLP_MAX_CLASSES];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA engine, SASS UBLKCP); completion is counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// relu that propagates NaN like torch's clamp_min
__device__ __forceinline__ float relu_nan(float x) {
    float y;
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(y) : "f"(x));
    return y;
}

template <bool RELU>
__device__ __forceinline__ float act(float x) { return RELU ? relu_nan(x) : x; }

// sum of act(s[idx]) over one channel's HW values; NCHW: contiguous, NHWC: stride CH
template <bool RELU>
__device__ __forceinline__ float channel_sum(const float* s, int HW, int stride, int rot) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (rot < 0) {   // conflict-free as is
        int i = 0;
        for (; i + 4 <= HW; i += 4) {
            a0 += act<RELU>(s[(i + 0) * stride]); a1 += act<RELU>(s[(i + 1) * stride]);
            a2 += act<RELU>(s[(i + 2) * stride]); a3 += act<RELU>(s[(i + 3) * stride]);
        }
        for (; i < HW; ++i) a0 += act<RELU>(s[i * stride]);
    } else {         // even HW, contiguous channels: lanes start at different offsets of their run
        int idx = rot;
        for (int i = 0; i < HW; ++i) {
            a0 += act<RELU>(s[idx]);
            idx = (idx + 1 == HW) ? 0 : idx + 1;
        }
    }
    return (a0 + a1) + (a2 + a3);
}

// named barriers: 1 = the compute warps among themselves, 2 = hand-off of a sample's partials to the
// epilogue warp (compute warps only arrive)
__device__ __forceinline__ void bar_compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kPoolThreads) : "memory"); }
__device__ __forceinline__ void bar_handoff_arrive() { asm volatile("bar.arrive 2, %0;" ::"n"(kPoolThreads + 32) : "memory"); }
__device__ __forceinline__ void bar_handoff_sync() { asm volatile("bar.sync 2, %0;" ::"n"(kPoolThreads + 32) : "memory"); }
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// 3-D tiled TMA load (NHWC feature maps: box = [1 sample][HW rows][CH channels])
__device__ __forceinline__ void tma_load_3d(float* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

// Warps 0..3 stream and pool (one thread per channel of a stage); warp 4 is the epilogue warp: it
// turns the compute threads' per-sample partials into the CTA's partial dots, synchronises with the
// other CTAs of the cluster and (in rank 0) writes the similarities.  The compute warps never wait
// for it except through the cluster barrier of the PREVIOUS sample, which has long completed.
// Fully unrolled form for compile-time HW / stride (7x7 maps, 128-channel stages): all loads are issued
// back to back with immediate offsets and summed by a tree.  r01 ncu/knob sweep: with the generic loop
// (4 dependent chains, runtime trip count) one CTA needed ~3000 cycles per 25 KB stage and the kernel
// was bound by warp latency, not HBM.
template <bool RELU, int HW, int STRIDE>
__device__ __forceinline__ float channel_sum_fixed(const float* s) {
    float v[HW];
#pragma unroll
    for (int i = 0; i < HW; ++i) v[i] = act<RELU>(s[i * STRIDE]);
#pragma unroll
    for (int w = 1; w < HW; w <<= 1)
#pragma unroll
        for (int i = 0; i + w < HW; i += 2 * w) v[i] += v[i + w];
    return v[0];
}

template <int NVT>
__global__ void __launch_bounds__(kPoolThreads + 32) pool_tag_kernel(const __grid_constant__ PoolArgs a,
                                                                      const __grid_constant__ CUtensorMap tmap) {
    cg::cluster_group cluster = cg::this_cluster();
    const int G = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int64_t cluster_id = blockIdx.x / G;
    const int64_t n_clusters = gridDim.x / G;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int PV = a.nv + 1;      // partial values per sample: nv dots + |f|^2

    const int ch_begin = rank * a.Dr;
    const int ch_end = min(a.D, ch_begin + a.Dr);
    const int my_ch = max(0, ch_end - ch_begin);
    const int nsub = (my_ch + a.CH - 1) / a.CH;
    const int stage_floats = a.CH * a.HW;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);                       // [nstage][CH*HW]
    float* sV = ring + (size_t)a.nstage * stage_floats;                     // [nv][Dr]
    float* sAcc = sV + (size_t)a.nv * a.Dr;                                 // [kPoolMaxVec + 2][4]: per-warp partials
    float* sPart = sAcc + (kPoolMaxVec + 2) * 4;                            // [2][kPoolMaxVec + 2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sPart + 2 * (kPoolMaxVec + 2));   // [nstage]; all counts above are even

    if (t == 0) {
        for (int s = 0; s < a.nstage; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t my_samples = (cluster_id < a.B) ? (a.B - cluster_id + n_clusters - 1) / n_clusters : 0;
    const bool scoring = a.nv > 0;

    // producer (lane 0 of warp 0): fills ring stage `st` with the next item in (sample, sub-slab) order
    int64_t p_si = 0;
    int p_sub = 0;
    auto issue = [&](int st) {
        if (p_si >= my_samples || nsub == 0) return;
        const int64_t b = cluster_id + p_si * n_clusters;
        const int c0 = ch_begin + p_sub * a.CH;
        float* dst = ring + (size_t)st * stage_floats;
        if (!a.nhwc) {
            const uint32_t bytes = (uint32_t)min(a.CH, ch_end - c0) * (uint32_t)a.HW * 4u;
            mbar_expect_tx(bars + st, bytes);
            bulk_g2s(dst, a.fmap + ((int64_t)b * a.D + c0) * a.HW, bytes, bars + st);
        } else {
            mbar_expect_tx(bars + st, (uint32_t)stage_floats * 4u);      // the box is always full size
            tma_load_3d(dst, &tmap, c0, 0, (int)b, bars + st);
        }
        if (++p_sub == nsub) { p_sub = 0; ++p_si; }
    };
    if (t == 0)
        for (int s = 0; s < a.nstage; ++s) issue(s);    // the first loads fly while the class vectors are staged

    // class vectors of my channel range
    {
        const int dr4 = a.Dr >> 2;
        float4* dst4 = reinterpret_cast<float4*>(sV);
#pragma unroll 4
        for (int idx = t; idx < a.nv * dr4; idx += kPoolThreads + 32) {
            const int j = idx / dr4, d = (idx - j * dr4) << 2;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ch_begin + d < a.D) v = __ldg(reinterpret_cast<const float4*>(a.table + (int64_t)j * a.D + ch_begin + d));
            dst4[idx] = v;
        }
    }
    __syncthreads();

    if (warp == kPoolThreads / 32) {
        // ------------------------------------------------------------------ epilogue warp
        if (scoring) {
            int par = 0;
            for (int64_t si = 0; si < my_samples; ++si) {
                const int64_t b = cluster_id + si * n_clusters;
                bar_handoff_sync();                         // the compute warps have stored sample si's partials
                for (int j = lane; j < PV; j += 32) {
                    const float4 w = *reinterpret_cast<const float4*>(sAcc + j * 4);
                    sPart[par * (kPoolMaxVec + 2) + j] = (w.x + w.y) + (w.z + w.w);
                }
                __syncwarp();
                cluster_arrive_release();
                cluster_wait();
                if (rank == 0 && lane < a.npair) {
                    // combine the G partials in rank order (deterministic)
                    float d0 = 0.f, d1 = 0.f, ff = 0.f;
                    float r0[kPoolMaxCluster], r1[kPoolMaxCluster], rf[kPoolMaxCluster];
#pragma unroll
                    for (int g = 0; g < kPoolMaxCluster; ++g) {
                        r0[g] = r1[g] = rf[g] = 0.f;
                        if (g < G) {
                            const float* rp = cluster.map_shared_rank(sPart, g) + par * (kPoolMaxVec + 2);
                            if (a.fold) r0[g] = rp[lane];
                            else { r0[g] = rp[2 * lane]; r1[g] = rp[2 * lane + 1]; }
                            rf[g] = rp[a.nv];
                        }
                    }
#pragma unroll
                    for (int g = 0; g < kPoolMaxCluster; ++g) { d0 += r0[g]; d1 += r1[g]; ff += rf[g]; }
                    const float nf = sqrtf(ff);
                    float out;
                    if (a.fold) {
                        out = __fmul_rn(d0, __frcp_rn(nf));
                    } else {
                        const float* norms = a.table + (int64_t)a.nv * a.D;
                        const float c0 = __fmul_rn(d0, __frcp_rn(__fmul_rn(nf, norms[2 * lane])));
                        const float c1 = __fmul_rn(d1, __frcp_rn(__fmul_rn(nf, norms[2 * lane + 1])));
                        out = __fsub_rn(c0, c1);
                    }
                    a.sim[(int64_t)a.cls[lane] * a.ld_sim + b] = out;
                }
                par ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------------ compute warps
        const float den = (float)a.HW;
        const bool fast49 = (a.HW == 49 && a.CH == kPoolThreads);
        const int rot = (!a.nhwc && !(a.HW & 1)) ? (t % a.HW) : -1;
        float acc[NVT];
        float nrm = 0.f;
#pragma unroll
        for (int j = 0; j < NVT; ++j) acc[j] = 0.f;
        int st = 0;             // consumer cursor: ring stage and its mbarrier phase
        uint32_t phase = 0;
        const int nvm1 = a.nv - 1;

        for (int64_t si = 0; si < my_samples; ++si) {
            const int64_t b = cluster_id + si * n_clusters;
            for (int sub = 0; sub < nsub; ++sub) {
                mbar_wait(bars + st, phase);
                const int c0 = ch_begin + sub * a.CH;
                const int chn = min(a.CH, ch_end - c0);
                if (t < chn) {
                    const float* s = ring + (size_t)st * stage_floats + (a.nhwc ? t : t * a.HW);
                    const int stride = a.nhwc ? a.CH : 1;
                    float sum;
                    if (fast49) {
                        if (a.nhwc) sum = a.relu ? channel_sum_fixed<true, 49, kPoolThreads>(s) : channel_sum_fixed<false, 49, kPoolThreads>(s);
                        else        sum = a.relu ? channel_sum_fixed<true, 49, 1>(s) : channel_sum_fixed<false, 49, 1>(s);
                    } else {
                        sum = a.relu ? channel_sum<true>(s, a.HW, stride, rot) : channel_sum<false>(s, a.HW, stride, rot);
                    }
                    const float f = __fdiv_rn(sum, den);
                    if (a.feat) a.feat[b * a.ld_feat + c0 + t] = f;
                    if (scoring) {
                        // branch-free: rows past nv-1 re-read the last row into accumulators nobody uses
                        // (r01 ncu: a predicated loop compiled to nv dependent branch + LDS + FFMA chains)
                        const float* v = sV + (c0 - ch_begin) + t;
                        float vv[NVT];
                        int off = 0;
#pragma unroll
                        for (int j = 0; j < NVT; ++j) { vv[j] = v[off]; off += (j < nvm1) ? a.Dr : 0; }
#pragma unroll
                        for (int j = 0; j < NVT; ++j) acc[j] = fmaf(f, vv[j], acc[j]);
                        nrm = fmaf(f, f, nrm);
                    }
                }
                bar_compute_sync();                     // every compute thread is done with this stage
                if (t == 0) issue(st);
                if (++st == a.nstage) { st = 0; phase ^= 1u; }
            }
            if (scoring) {
                // Hand the warp's partials to the epilogue warp.  It read the previous sample's before it
                // arrived on the cluster barrier of that sample, which the wait below implies (that phase
                // completed long ago in steady state: the compute warps never really block here).
#pragma unroll
                for (int j = 0; j < NVT; ++j)
                    if (j < a.nv) acc[j] = warp_sum(acc[j]);
                nrm = warp_sum(nrm);
                if (si > 0) cluster_wait();             // phase si-1
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < NVT; ++j)
                        if (j < a.nv) sAcc[j * 4 + warp] = acc[j];
                    sAcc[a.nv * 4 + warp] = nrm;
                }
#pragma unroll
                for (int j = 0; j < NVT; ++j) acc[j] = 0.f;
                nrm = 0.f;
                bar_handoff_arrive();
                cluster_arrive_relaxed();               // phase si (the epilogue warp does the release)
            }
        }
        if (scoring && my_samples > 0) cluster_wait();
    }
    cluster.sync();   // nobody exits while rank 0 may still read its partials
}

// ---- class-vector table ---------------------------------------------------------------------
// One CTA per scored class: |P[2c]|, |P[2c+1]| (torch.norm: sqrt of the fp32 sum of squares),
// then the vectors the similarity kernels multiply with.
__global__ void __launch_bounds__(256) sim_table_kernel(const float* __restrict__ proto, int D, int fold, int npair,
                                                        const __grid_constant__ PoolArgs a, float* __restrict__ table) {
    __shared__ float red[2][8];
    __shared__ float nrm[2];
    const int q = blockIdx.x;
    const int c = a.cls[q];
    const float* p0 = proto + (int64_t)(2 * c) * D;
    const float* p1 = p0 + D;
    float s0 = 0.f, s1 = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) { s0 = fmaf(p0[d], p0[d], s0); s1 = fmaf(p1[d], p1[d], s1); }
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
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        if (fold) {
            table[(int64_t)q * D + d] = __fsub_rn(__fdiv_rn(p0[d], nrm[0]), __fdiv_rn(p1[d], nrm[1]));
        } else {
            table[(int64_t)(2 * q) * D + d] = p0[d];
            table[(int64_t)(2 * q + 1) * D + d] = p1[d];
        }
    }
    if (threadIdx.x < 2) table[(int64_t)nv * D + 2 * q + threadIdx.x] = nrm[threadIdx.x];
}

static int classes_of(uint32_t mask, int C, int8_t* cls) {
    int n = 0;
    for (int c = 0; c < FMLP_MAX_CLASSES; ++c) cls[c] = 0;
    for (int c = 0; c < C; ++c)
        if ((mask >> c) & 1u) cls[n++] = (int8_t)c;
    return n;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TensorMapEncodeFn>(p);
    }
    return fn;
}

static size_t pool_fixed_floats(int nv, int Dr) {
    return (size_t)nv * Dr + 6 * (kPoolMaxVec + 2);
}

template <int NVT>
static int launch_pool(const PoolArgs& a, const CUtensorMap& tmap, int G, cudaStream_t st) {
    auto kern = pool_tag_kernel<NVT>;
    const size_t fl = (size_t)a.nstage * a.CH * a.HW + pool_fixed_floats(a.nv, a.Dr);
    const size_t smem = fl * sizeof(float) + (size_t)a.nstage * sizeof(uint64_t);
    if (smem > 227u * 1024u) return FMLP_ERR_UNSUPPORTED;
    static size_t configured = 0;
    if (smem > 48u * 1024u && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    const int sms = sm_count();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(kPoolThreads + 32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    // persistent when the batch is larger than what is resident: clusters loop over samples
    static size_t occ_smem = 0;
    static int occ_G = 0, occ_clusters = 0;
    if (occ_smem != smem || occ_G != G) {
        cfg.gridDim = dim3((unsigned)(G * sms));
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        if (e != cudaSuccess) return (int)e;
        occ_smem = smem; occ_G = G; occ_clusters = n < 1 ? 1 : n;
    }
    int64_t clusters = occ_clusters;
    if (clusters > a.B) clusters = a.B;
    cfg.gridDim = dim3((unsigned)(clusters * G));
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, tmap);
    return e == cudaSuccess ? launch_status() : (int)e;
}

}  // namespace fmlp

using namespace fmlp;

extern "C" size_t fmlp_sim_table_bytes(int C, int D) {
    if (C < 1 || D < 1) return 0;
    return (((size_t)2 * C * D + 2 * C) * sizeof(float) + 15) & ~(size_t)15;
}

extern "C" int fmlp_sim_table_build_f32(const float* proto, int C, int D, uint32_t classes, int mode,
                                        float* table, fmlp_stream_t stream) {
    if (!proto || !table || C < 1 || C > FMLP_MAX_CLASSES || D < 1) return FMLP_ERR_BAD_ARG;
    if (mode != FMLP_SIM_PAIR && mode != FMLP_SIM_FOLDED) return FMLP_ERR_BAD_ARG;
    PoolArgs a = {};
    const int npair = classes_of(classes, C, a.cls);
    if (npair == 0) return FMLP_OK;
    sim_table_kernel<<<npair, 256, 0, (cudaStream_t)stream>>>(proto, D, mode == FMLP_SIM_FOLDED, npair, a, table);
    return launch_status();
}

extern "C" int fmlp_pool_tag_f32(const float* fmap, int layout, int B, int D, int HW, int relu,
                                 const float* table, int C, uint32_t classes, int mode, float* feat,
                                 int64_t ld_feat, float* sim, int64_t ld_sim, fmlp_stream_t stream) {
    if (!fmap || B < 0 || D < 4 || HW < 1 || (layout != FMLP_FMAP_NCHW && layout != FMLP_FMAP_NHWC)) return FMLP_ERR_BAD_ARG;
    if (mode != FMLP_SIM_PAIR && mode != FMLP_SIM_FOLDED) return FMLP_ERR_BAD_ARG;
    if ((D & 3) || !aligned16(fmap)) return FMLP_ERR_UNSUPPORTED;
    if (feat && ld_feat < D) return FMLP_ERR_BAD_ARG;
    PoolArgs a = {};
    int npair = 0;
    if (classes != 0) {
        if (!table || !sim || C < 1 || C > FMLP_MAX_CLASSES || ld_sim < B) return FMLP_ERR_BAD_ARG;
        npair = classes_of(classes, C, a.cls);
    }
    if (!feat && npair == 0) return FMLP_ERR_BAD_ARG;   // nothing to produce
    if (B == 0) return FMLP_OK;
    a.npair = npair;
    a.fold = (mode == FMLP_SIM_FOLDED);
    a.nv = a.fold ? npair : 2 * npair;
    if (a.nv > kPoolMaxVec) return FMLP_ERR_UNSUPPORTED;   // > 16 classes in pair mode: use FOLDED
    a.fmap = fmap; a.table = table; a.feat = feat; a.sim = sim; a.ld_feat = ld_feat; a.ld_sim = ld_sim;
    a.B = B; a.D = D; a.HW = HW; a.relu = relu != 0; a.nhwc = (layout == FMLP_FMAP_NHWC);
    // stage = CH channels x HW values, at most ~32 KB
    int CH = kPoolThreads;
    while (CH > 4 && (size_t)CH * HW * sizeof(float) > 32u * 1024u) CH >>= 1;
    if ((size_t)CH * HW * sizeof(float) > 48u * 1024u) return FMLP_ERR_UNSUPPORTED;
    a.CH = CH;
    // cluster: ranks split the channels; two stages' worth of channels per rank when D allows
    int G = (D + 2 * CH - 1) / (2 * CH);
    if (G < 1) G = 1;
    if (G > kPoolMaxCluster) G = kPoolMaxCluster;
    a.Dr = (((D + G - 1) / G) + 3) & ~3;
    const size_t stage_bytes = (size_t)CH * HW * sizeof(float);
    const size_t fixed = (pool_fixed_floats(a.nv, a.Dr) + 64) * sizeof(float);
    // per-CTA shared-memory budget: two CTAs per SM by default (tuning knob FMLP_POOL_SMEM_KB)
    static size_t budget = 0;
    if (budget == 0) {
        budget = 55u * 1024u;   // 4 CTAs per SM with two 25 KB stages: r01 sweep, more resident CTAs beat deeper rings
        if (const char* e = getenv("FMLP_POOL_SMEM_KB")) { int v = atoi(e); if (v >= 16 && v <= 226) budget = (size_t)v * 1024u; }
    }
    int nstage = (int)(budget > fixed ? (budget - fixed) / stage_bytes : 0);
    if (nstage > kPoolMaxStages) nstage = kPoolMaxStages;
    if (nstage < 2) nstage = 2;
    a.nstage = nstage;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (a.nhwc) {
        // [B][HW][D] as a rank-3 tensor (innermost first: D, HW, B); one box = CH channels x HW rows of one
        // sample, channels past D are zero-filled
        TensorMapEncodeFn enc = tensor_map_encoder();
        if (!enc) return FMLP_ERR_UNSUPPORTED;
        cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)HW, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)D * sizeof(float), (cuuint64_t)D * HW * sizeof(float)};
        cuuint32_t box[3] = {(cuuint32_t)CH, (cuuint32_t)HW, 1u};
        cuuint32_t estr[3] = {1u, 1u, 1u};
        CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(fmap), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return FMLP_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (a.nv <= 4) return launch_pool<4>(a, tmap, G, st);
    if (a.nv <= 8) return launch_pool<8>(a, tmap, G, st);
    if (a.nv <= 16) return launch_pool<16>(a, tmap, G, st);
    return launch_pool<32>(a, tmap, G, st);
}
