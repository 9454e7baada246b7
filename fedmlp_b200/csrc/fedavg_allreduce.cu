// fedavg_allreduce.cu — K1 fused with its collective: rank-local K-way weighted fold of the client
// parameter buffers + two-shot all-reduce over NVLink peer memory, in ONE kernel per rank.
//
// The reference has no collective at all (single process, main.py:130,196 pass Python lists); the
// multi-GPU design shards clients over ranks and the only exchange step of a round is this
// aggregation (utils/FedAvg.py:7-14 semantics: global weighted mean of all clients).  Calling the
// local fold and then ncclAllReduce serialises 40 us of HBM streaming with ~70-110 us of all-reduce;
// here the transfer overlaps the math slice by slice:
//
//   fold     for every chunk c of the parameter vector and every slice s of it (staggered start so
//            the G ranks hit G different links): fold this rank's K clients over the slice (HBM-bound,
//            8 x 128-bit loads in flight per thread, same arithmetic as fedavg_flat_kernel) and store
//            the partial directly into rank s's inbox row [rank] with peer st.global -> the
//            reduce-scatter traffic ((G-1)/G*4P bytes out) rides under the K*4P bytes of local reads.
//            The CTA that finishes a chunk last (device counter) release-stores this call's epoch into
//            every peer's flag word of that chunk.
//   reduce   one chunk behind the fold: acquire the G flags of the chunk, sum the G inbox rows of the
//            own slice in FIXED rank order (deterministic, and every element of the result is
//            computed by exactly one rank, so all ranks hold bit-identical parameters), store the
//            result slice into every rank's result buffer (all-gather by peer stores) and signal the
//            chunk the same way.  The kernel ends when all peers' result chunks have landed.
// No grid-wide barrier, no second kernel: the two transfer directions of neighbouring chunks overlap
// each other and the fold (r01: 118 us with two global barriers at 8 GPUs).
//
// Buffers (stage inbox [G][L], result [G*L], flags [2][8]) are symmetric-memory allocations
// made and exchanged by the host (torch.distributed._symmetric_memory); the kernel only sees raw
// peer pointers.  Weights arrive pre-normalised (n_k / sum n), so no divide pass is needed.
#include <cstdlib>

#include "common.cuh"

namespace fmlp {

constexpr int kArMaxRanks = 8;
constexpr int kArMaxChunks = FMLP_AR_MAX_CHUNKS;
#ifndef FMLP_AR_THREADS
#define FMLP_AR_THREADS 480   // round 2: half an SM's threads and register file, so that the kernel shares the SMs with the
                              // similarity / prototype kernels of the other streams (992 = the whole SM: 0.192 -> 0.185 ms
                              // per round at 2 GPUs, 0.212 -> 0.206 at 8, profiles/r02_multi_gpu_variants.txt)
#endif
constexpr int kArThreads = FMLP_AR_THREADS;   // compute threads per CTA (+ one signal warp); one CTA per SM: every CTA's
                                              // per-chunk release costs its SM ~1 us, so fewer, larger CTAs
constexpr int kArUnroll = 8;

// Flag words of one rank (uint32, symmetric memory, zero-initialised once):
//   [phase 0..1][chunk][source rank]   epoch flags written by the peers (phase 0: "my partials of this
//                                      chunk are in your inbox", phase 1: "my result slice of this chunk
//                                      is in your result buffer")
//   [phase 0..1][chunk]                local CTA counters (which CTA of this rank finishes a chunk last)
//   [1]                                local CTA counter for the end of the call
__host__ __device__ constexpr int ar_flag(int phase, int c, int g) { return (phase * kArMaxChunks + c) * kArMaxRanks + g; }
__host__ __device__ constexpr int ar_count(int phase, int c) { return 2 * kArMaxChunks * kArMaxRanks + phase * kArMaxChunks + c; }
constexpr int kArDoneCount = 2 * kArMaxChunks * kArMaxRanks + 2 * kArMaxChunks;
static_assert(kArDoneCount < FMLP_AR_FLAG_WORDS, "flag buffer too small");

struct ArArgs {
    const float* src[FMLP_MAX_CLIENTS];
    float w[FMLP_MAX_CLIENTS];
    float* stage[kArMaxRanks];      // stage[g] = rank g's inbox, [G][L]
    float* result[kArMaxRanks];     // result[g] = rank g's result buffer, >= G*L floats
    uint32_t* flags[kArMaxRanks];   // flags[g] = rank g's flag words
    int64_t P;                      // parameters (multiple of 4)
    int64_t L;                      // floats per rank over all chunks = NC * Lc
    int64_t Lc;                     // floats per (chunk, rank) slice (multiple of 4), G*L >= P
    int K, rank, G, NC;
    uint32_t* epoch_dev;            // this rank's call counter (device memory, graph-replay safe)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void red_release_cta_shared(int* p, int v) {
    asm volatile("red.release.cta.shared::cta.add.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta_shared(const int* p) {
    int v;
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}

constexpr int kArComputeWarps = kArThreads / 32;

// Compute warps: "this warp has issued all its peer stores of (phase, chunk)".  No fence, no wait:
// the signal warp makes them visible.  (An earlier version fenced in every compute thread at every
// chunk boundary; the drain cost ~20 us per chunk at 2 GPUs.)
__device__ __forceinline__ void warp_chunk_done(int* s_done) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) red_release_cta_shared(s_done, 1);
}
// Signal warp: once all compute warps of this CTA are done with (phase, chunk c), count the CTA with a
// GPU-scope release; the CTA of this rank that gets here last acquires the count and publishes the
// rank's epoch to every peer with a SYSTEM-scope release.  Only those G lanes of one CTA per (phase,
// chunk) execute a system-scope fence: MEMBAR.SYS in every CTA cost ~20 us per chunk at 2 GPUs (r01:
// 101 / 145 / 236 us for 1 / 4 / 8 chunks against 78 / 65 / 65 us with the fences compiled out).  The
// chain peer stores -> release.cta (s_done) -> acquire.cta -> release.gpu (counter) -> acquire.gpu ->
// release.sys (flag) -> acquire.sys is causality-ordered in the PTX memory model (every link is
// morally strong at its own scope), so the stores are visible before the flag.
__device__ __forceinline__ void signal_chunk(const ArArgs& a, uint32_t epoch, int phase, int c, const int* s_done) {
    const int lane = threadIdx.x & 31;
    int last = 0;
    if (lane == 0) {
        while (ld_acquire_cta_shared(s_done) < kArComputeWarps) { __nanosleep(20); }
        uint32_t old;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(a.flags[a.rank] + ar_count(phase, c)) : "memory");
        last = (old + 1u == epoch * gridDim.x);
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    __syncwarp();   // lane 0's acquire is ordered before the other lanes' release stores
    if (last && lane < a.G) st_release_sys(a.flags[lane] + ar_flag(phase, c, a.rank), epoch);
}
// Every rank has published (phase, chunk c) of this call (per warp: lanes poll one flag each).
__device__ __forceinline__ void wait_chunk(const ArArgs& a, uint32_t epoch, int phase, int c) {
    const int lane = threadIdx.x & 31;
    if (lane < a.G) {
        const uint32_t* f = a.flags[a.rank] + ar_flag(phase, c, lane);
        // epochs only grow; ">=" tolerates a peer that is already in its next call
        while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) { __nanosleep(32); }
    }
    __syncwarp();
}

// The parameter vector is cut into NC chunks of G slices (slice s of every chunk belongs to rank s).
// Compute warps:  fold(0), fold(1), reduce(0), fold(2), reduce(1), ...: the all-gather stores of chunk
// c-1 and the reduce-scatter stores of chunk c+1 share the links while the fold keeps HBM busy; a
// chunk's reduce only needs THAT chunk's partials from the peers, which were sent a whole chunk
// earlier.  The compute warps never fence and never meet at a barrier; a ninth warp per CTA does the
// system-scope fences and the signalling behind them.  There is no grid-wide barrier: the launch is
// cooperative only to guarantee that all CTAs are resident (reduce(c) waits for the other CTAs'
// fold(c)).
__global__ void __launch_bounds__(kArThreads + 32, (kArThreads + 32) > 512 ? 1 : (kArThreads + 32) > 320 ? 2 : 3) fedavg_allreduce_kernel(const __grid_constant__ ArArgs a) {
    __shared__ int s_done[2][kArMaxChunks];
    // The epoch lives in device memory so that a CUDA-graph replay (identical kernel arguments)
    // still advances it: every thread reads it on entry, the CTA that finishes last bumps it.
    const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(a.epoch_dev) + 1u;
    if (threadIdx.x < 2 * kArMaxChunks) (&s_done[0][0])[threadIdx.x] = 0;
    __syncthreads();

    if (threadIdx.x >= kArThreads) {
        // ------------------------------------------------------------------ signal warp
        for (int step = 0; step <= a.NC; ++step) {
            if (step < a.NC) signal_chunk(a, epoch, 0, step, &s_done[0][step]);
            if (step >= 1) signal_chunk(a, epoch, 1, step - 1, &s_done[1][step - 1]);
        }
        // the call is complete when every peer's result slices of every chunk have landed here
        if (blockIdx.x == 0)
            for (int c = 0; c < a.NC; ++c) wait_chunk(a, epoch, 1, c);
        if ((threadIdx.x & 31) == 0) {
            const uint32_t old = atomicAdd(a.flags[a.rank] + kArDoneCount, 1u);
            if (old + 1u == epoch * gridDim.x) {   // every CTA of this rank has read the epoch and is done
                __threadfence();
                *a.epoch_dev = epoch;
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- compute warps
    const int64_t tid = (int64_t)blockIdx.x * kArThreads + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * kArThreads;
    const int64_t Lv = a.Lc >> 2;   // float4 per (chunk, rank) slice
    const float* inbox = a.stage[a.rank];

    for (int step = 0; step <= a.NC; ++step) {
        if (step < a.NC) {
            // ---- fold chunk `step`: partial slices go straight to their owners (staggered start so
            //      the G ranks hit G different links)
            const int64_t chunk0 = (int64_t)step * a.G * a.Lc;
            for (int64_t idx = tid; idx < (int64_t)a.G * Lv; idx += nthreads) {
                const int t = (int)(idx / Lv);
                const int64_t v = idx - (int64_t)t * Lv;
                int s = a.rank + 1 + t;
                if (s >= a.G) s -= a.G;
                const int64_t e = chunk0 + (int64_t)s * a.Lc + (v << 2);
                if (e >= a.P) continue;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                int i = 0;
                for (; i + kArUnroll <= a.K; i += kArUnroll) {
                    float4 x[kArUnroll];
#pragma unroll
                    for (int u = 0; u < kArUnroll; ++u) x[u] = ld_stream_f4(a.src[i + u] + e);
#pragma unroll
                    for (int u = 0; u < kArUnroll; ++u) {
                        acc.x = fmaf(x[u].x, a.w[i + u], acc.x); acc.y = fmaf(x[u].y, a.w[i + u], acc.y);
                        acc.z = fmaf(x[u].z, a.w[i + u], acc.z); acc.w = fmaf(x[u].w, a.w[i + u], acc.w);
                    }
                }
                for (; i < a.K; ++i) {
                    const float4 x = ld_stream_f4(a.src[i] + e);
                    acc.x = fmaf(x.x, a.w[i], acc.x); acc.y = fmaf(x.y, a.w[i], acc.y);
                    acc.z = fmaf(x.z, a.w[i], acc.z); acc.w = fmaf(x.w, a.w[i], acc.w);
                }
                *reinterpret_cast<float4*>(a.stage[s] + (int64_t)a.rank * a.L + (int64_t)step * a.Lc + (v << 2)) = acc;   // peer store
            }
            warp_chunk_done(&s_done[0][step]);
        }
        if (step >= 1) {
            // ---- reduce my slice of chunk c in rank order (deterministic; every element of the result
            //      is computed by exactly one rank, so all ranks hold bit-identical parameters) and
            //      all-gather it by peer stores
            const int c = step - 1;
            const int64_t e0 = (int64_t)c * a.G * a.Lc + (int64_t)a.rank * a.Lc;
            const float* in_c = inbox + (int64_t)c * a.Lc;
            if (tid - (threadIdx.x & 31) < Lv) wait_chunk(a, epoch, 0, c);   // warp-uniform: only warps with work poll
            for (int64_t v = tid; v < Lv; v += nthreads) {
                const int64_t e = e0 + (v << 2);
                if (e >= a.P) continue;
                float4 acc = __ldcg(reinterpret_cast<const float4*>(in_c + (v << 2)));
                for (int g = 1; g < a.G; ++g) {
                    const float4 x = __ldcg(reinterpret_cast<const float4*>(in_c + (int64_t)g * a.L + (v << 2)));
                    acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
                }
                for (int g = 0; g < a.G; ++g) {
                    int dst = a.rank + g;           // start with the local copy, then walk the peers
                    if (dst >= a.G) dst -= a.G;
                    *reinterpret_cast<float4*>(a.result[dst] + e) = acc;
                }
            }
            warp_chunk_done(&s_done[1][c]);
        }
    }
}

}  // namespace fmlp

using namespace fmlp;

extern "C" int fmlp_fedavg_allreduce_f32(const float* const* srcs, const float* weights, int K, int64_t P,
                                         float* const* stage_ptrs, float* const* result_ptrs,
                                         uint32_t* const* flag_ptrs, int64_t slice_len, int n_chunks, int rank,
                                         int world, uint32_t* epoch_dev, fmlp_stream_t stream) {
    if (!srcs || !weights || !stage_ptrs || !result_ptrs || !flag_ptrs || !epoch_dev || K < 1 || K > FMLP_MAX_CLIENTS || P < 0 ||
        world < 1 || world > kArMaxRanks || rank < 0 || rank >= world || n_chunks < 1 || n_chunks > kArMaxChunks)
        return FMLP_ERR_BAD_ARG;
    if ((P & 3) || slice_len % (4 * (int64_t)n_chunks) || slice_len * world < P) return FMLP_ERR_UNSUPPORTED;
    ArArgs a;
    for (int i = 0; i < FMLP_MAX_CLIENTS; ++i) {
        a.src[i] = i < K ? srcs[i] : nullptr;
        a.w[i] = i < K ? weights[i] : 0.f;
        if (i < K && (!srcs[i] || !aligned16(srcs[i]))) return FMLP_ERR_UNSUPPORTED;
    }
    for (int g = 0; g < kArMaxRanks; ++g) {
        a.stage[g] = g < world ? stage_ptrs[g] : nullptr;
        a.result[g] = g < world ? result_ptrs[g] : nullptr;
        a.flags[g] = g < world ? flag_ptrs[g] : nullptr;
        if (g < world && (!stage_ptrs[g] || !result_ptrs[g] || !flag_ptrs[g] || !aligned16(stage_ptrs[g]) || !aligned16(result_ptrs[g])))
            return FMLP_ERR_BAD_ARG;
    }
    a.P = P; a.L = slice_len; a.Lc = slice_len / n_chunks; a.NC = n_chunks;
    a.K = K; a.rank = rank; a.G = world; a.epoch_dev = epoch_dev;
    const int sms = sm_count();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    static int per_sm = 0;
    if (per_sm == 0) {
        int b = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, fedavg_allreduce_kernel, kArThreads + 32, 0);
        if (e != cudaSuccess) return (int)e;
        per_sm = 1;   // one CTA per SM (see FMLP_AR_THREADS); FMLP_AR_CTAS_PER_SM raises it up to the occupancy limit
        (void)b;
        // tuning knob: a smaller footprint leaves SM resources to the concurrent tagging/prototype
        // kernels during the NVLink-bound phases
        if (const char* e = getenv("FMLP_AR_CTAS_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= b && v <= 4) per_sm = v; }
    }
    // The per-chunk CTA counters compare against epoch * gridDim.x, so the grid depends on nothing but
    // the device: one full wave (CTAs without work just count).
    const int64_t blocks = (int64_t)sms * per_sm;
    void* args[] = {(void*)&a};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)fedavg_allreduce_kernel, dim3((unsigned)blocks), dim3(kArThreads + 32),
                                                args, 0, (cudaStream_t)stream);
    return e == cudaSuccess ? launch_status() : (int)e;
}
