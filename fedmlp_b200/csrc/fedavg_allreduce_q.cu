// fedavg_allreduce_q.cu — round-2 server aggregation across GPUs: K-way weighted fold of this rank's
// client buffers + all-reduce over NVLink / NVSwitch in ONE non-cooperative kernel per rank.
//
// Semantics = utils/FedAvg.py:7-14 over the clients of all ranks (global weighted mean of the flat
// parameters), with the small tails of main.py:218-234 riding the same exchange:
//   fp32 vector   [ P parameters | T = 2C*D prototype sums ]   sum_k w_k * x_k   (w_k pre-normalised)
//   fp64 tail     [ M scalars ]  per-class weight sums / difficulty sums (FedAvg_proto / FedAvg_tao divisors)
//
// Why a new kernel (round-1's fedavg_allreduce.cu stays as the fallback for GPUs without multicast):
// the round-1 kernel was a cooperative launch of one 1024-thread CTA per SM — it could only start
// when the whole machine was free, so it serialised against the tagging chain on the other stream,
// and it moved 2*(G-1)/G*4P bytes per GPU by peer stores (115 us at 8 GPUs for a 28 MB vector).
//   * WORK QUEUE, no co-residency requirement.  The call is a list of items in a fixed order
//         F(0) F(1) R(0) F(2) R(1) ... F(NC-1) R(NC-2) R(NC-1) FINAL
//     F(c) = fold pieces of chunk c (HBM-bound), R(c) = reduce + broadcast pieces of this rank's slice
//     of chunk c (NVLink-bound).  Every CTA pulls the next item index with one atomic; an item only
//     ever waits for items EARLIER in the list (which are running or done, because their index was
//     handed out first) or for a peer's fold, which waits for nothing.  So the kernel is deadlock-free
//     for any number of resident CTAs and is launched like any other kernel: it starts on whatever SM
//     slots are free and shares the SMs with the tagging / prototype kernels of the other stream.
//   * NVLS.  With a multicast mapping of the symmetric buffers (NVSwitch), R(c) is
//     multimem.ld_reduce (the switch adds the G partial slices) + multimem.st (the switch broadcasts
//     the result slice): 4P bytes per GPU and direction instead of 2*(G-1)/G*4P each way, and no
//     inbox rows.  Without multicast the same items pull the G partial slices with peer loads in
//     rank order and push the result with peer stores.
//   * Every element of the result is computed by exactly one rank and broadcast, so all ranks hold
//     bit-identical values; the P2P path is also run-to-run deterministic (fixed rank order).
//   * Signalling as in round 1: compute warps never fence; a signal warp per CTA observes item
//     completion through shared memory, counts it with a GPU-scope acq_rel atomic, and the CTA that
//     completes a chunk publishes the epoch to every peer with st.release.sys.  The signal warp is
//     also the CTA's scheduler: it prefetches the next item while the compute warps work.
#include <cstdlib>

#include "common.cuh"

namespace fmlp {

constexpr int kQMaxRanks = 8;
constexpr int kQMaxChunks = FMLP_AR_MAX_CHUNKS;
#ifndef FMLP_ARQ_THREADS
#define FMLP_ARQ_THREADS 1024
#endif
#ifndef FMLP_ARQ_CTAS_PER_SM
#define FMLP_ARQ_CTAS_PER_SM 1
#endif
constexpr int kQThreads = FMLP_ARQ_THREADS - 32;      // compute threads per CTA (+ one signal / scheduler warp)
constexpr int kQCtasPerSm = FMLP_ARQ_CTAS_PER_SM;
constexpr int kQWarps = kQThreads / 32;
constexpr int kQUnroll = 8;

// Flag words of one rank (uint32, symmetric memory, zero-initialised once):
//   [phase][chunk][source rank]  epoch flags written by the peers (phase 0: "my partial of this chunk is
//                                complete", phase 1: "my result slice of this chunk is in your buffer")
//   [phase][chunk]               local completion counters
//   queue, done                  local work-queue head and finished-CTA counter
__host__ __device__ constexpr int q_flag(int phase, int c, int g) { return (phase * kQMaxChunks + c) * kQMaxRanks + g; }
__host__ __device__ constexpr int q_count(int phase, int c) { return 2 * kQMaxChunks * kQMaxRanks + phase * kQMaxChunks + c; }
constexpr int kQQueue = 2 * kQMaxChunks * kQMaxRanks + 2 * kQMaxChunks;
constexpr int kQDone = kQQueue + 1;
static_assert(kQDone < FMLP_AR_FLAG_WORDS, "flag buffer too small");

struct QArgs {
    const float* src0[FMLP_MAX_CLIENTS];   // parameters, P floats each
    const float* src1[FMLP_MAX_CLIENTS];   // tail vectors (prototypes), T floats each, or nullptr
    float w[FMLP_MAX_CLIENTS];
    float* partial[kQMaxRanks];            // partial[g] = rank g's partial buffer  [(P+T) floats | M doubles]
    float* result[kQMaxRanks];             // result[g]  = rank g's result buffer   (same layout)
    uint32_t* flags[kQMaxRanks];
    float* mc_partial;                     // multicast mappings (NVLS) or nullptr
    float* mc_result;
    const double* tail_src;                // this rank's M fp64 partial sums (device) or nullptr
    int64_t P, T;
    int64_t V;                             // float4 in the fp32 vector = (P + T) / 4
    int64_t Vc, Vs;                        // float4 per chunk / per (chunk, rank) slice
    int64_t Vf, Vr;                        // float4 per fold piece / reduce piece
    int64_t tail_off;                      // float offset of the fp64 tail in the buffers
    int M;
    int K, rank, G, NC, FJ, RJ;
    uint32_t* epoch_dev;
};

__device__ __forceinline__ void q_st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t q_ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void q_red_release_cta(int* p, int v) {
    asm volatile("red.release.cta.shared::cta.add.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int q_ld_acquire_cta(const int* p) {
    int v;
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void q_st_release_cta(int* p, int v) {
    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ float4 mm_ld_reduce_f4(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mm_st_f4(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ double mm_ld_reduce_f64(const double* mc) {
    double v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mm_st_f64(double* mc, double v) {
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc), "d"(v) : "memory");
}

#ifdef FMLP_ARQ_TRACE
// Debug build (tools/build_arq_variants.sh): per-CTA item timeline, 4 x u64 per record
//   [item index | kind<<32 | c<<40, t_published, t_done_observed, t_signalled]  (globaltimer ns)
constexpr int kTraceMax = 1 << 16;
__device__ unsigned long long g_trace[kTraceMax][4];
__device__ unsigned int g_trace_n;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#endif

// Item encoding: kind in the top bits.
enum : int { kItemFold = 0, kItemReduce = 1, kItemFinal = 2, kItemExit = 3 };
struct Item { int kind, c, j; };

// index in the fixed order  F(0) | F(1) R(0) | F(2) R(1) | ... | F(NC-1) R(NC-2) | R(NC-1) | FINAL
__device__ __forceinline__ Item decode_item(const QArgs& a, uint32_t idx) {
    Item it;
    const uint32_t FJ = a.FJ, RJ = a.RJ, NC = a.NC;
    if (idx < FJ) { it.kind = kItemFold; it.c = 0; it.j = (int)idx; return it; }
    idx -= FJ;
    const uint32_t per = FJ + RJ;
    const uint32_t s = idx / per;                    // group s (0-based) holds F(s+1) then R(s)
    if (s < NC - 1) {
        const uint32_t r = idx - s * per;
        if (r < FJ) { it.kind = kItemFold; it.c = (int)s + 1; it.j = (int)r; }
        else { it.kind = kItemReduce; it.c = (int)s; it.j = (int)(r - FJ); }
        return it;
    }
    idx -= (NC - 1) * per;
    if (idx < RJ) { it.kind = kItemReduce; it.c = (int)NC - 1; it.j = (int)idx; return it; }
    it.c = 0; it.j = 0;
    it.kind = (idx == RJ) ? kItemFinal : kItemExit;
    return it;
}

template <bool NVLS>
__global__ void __launch_bounds__(kQThreads + 32, kQCtasPerSm) fedavg_allreduce_q_kernel(const __grid_constant__ QArgs a) {
    __shared__ int s_item[2];       // published item index per slot
    __shared__ int s_ready[2];      // sequence number of the item in the slot (n + 1)
    __shared__ int s_done[2];       // compute warps that finished the slot's item
    // The epoch lives in device memory so that a CUDA-graph replay (identical kernel arguments) still
    // advances it: every CTA reads it on entry, the CTA that finishes last bumps it.
    const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(a.epoch_dev) + 1u;
    uint32_t* const my_flags = a.flags[a.rank];
    if (threadIdx.x < 2) { s_item[threadIdx.x] = 0; s_ready[threadIdx.x] = 0; s_done[threadIdx.x] = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t total_items = (uint32_t)(a.NC * (a.FJ + a.RJ) + 1);

    if (threadIdx.x >= kQThreads) {
        // ------------------------------------------------------------------ signal / scheduler warp
        // queue indices of this call are [ (epoch-1) * stride, ... ): the head is never reset, so a
        // CTA of the next call can never observe a stale head
        const uint32_t stride = total_items + gridDim.x;          // every CTA draws exactly one EXIT index
        const uint32_t base = (epoch - 1u) * stride;
        auto grab = [&]() -> uint32_t {
            uint32_t v = 0;
            if (lane == 0) v = atomicAdd(my_flags + kQQueue, 1u) - base;
            return __shfl_sync(0xffffffffu, v, 0);
        };
        auto publish = [&](int n, uint32_t idx) {
            if (lane == 0) {
                s_item[n & 1] = (int)idx;
                q_st_release_cta(&s_ready[n & 1], n + 1);
            }
        };
        auto wait_flags = [&](int phase, int c) {                  // all ranks have published (phase, chunk c)
            if (lane < a.G) {
                const uint32_t* f = my_flags + q_flag(phase, c, lane);
                while ((int32_t)(q_ld_acquire_sys(f) - epoch) < 0) { __nanosleep(32); }
            }
            __syncwarp();
        };
        auto flags_ready = [&](int phase, int c) -> bool {         // non-blocking form of wait_flags
            bool ok = true;
            if (lane < a.G) ok = (int32_t)(q_ld_acquire_sys(my_flags + q_flag(phase, c, lane)) - epoch) >= 0;
            ok = __all_sync(0xffffffffu, ok);
            __syncwarp();
            return ok;
        };
        int n = 0;
        uint32_t cur = grab();
        Item ci = decode_item(a, cur);
        if (ci.kind == kItemReduce) wait_flags(0, ci.c);
        if (ci.kind == kItemFinal) for (int c = 0; c < a.NC; ++c) wait_flags(1, c);
        publish(0, cur);
#ifdef FMLP_ARQ_TRACE
        unsigned long long t_pub_cur = gtime(), t_pub_nxt = 0;
#endif
        while (ci.kind != kItemExit) {
            const uint32_t nxt = grab();
            const Item ni = decode_item(a, nxt);
            // Prefetch: hand the next item to the compute warps before the current one has been signalled, so
            // they never idle between items — folds always, reduce items when the peers' partials of that chunk
            // are already published (the common case: R(c) sits a whole chunk behind F(c) in the list).
            bool published = false;
            if (ni.kind == kItemFold || ni.kind == kItemExit || (ni.kind == kItemReduce && flags_ready(0, ni.c))) {
                publish(n + 1, nxt);
                published = true;
#ifdef FMLP_ARQ_TRACE
                t_pub_nxt = gtime();
#endif
            }
            // ---- completion of the current item
            if (lane == 0) {
                while (q_ld_acquire_cta(&s_done[n & 1]) < kQWarps) { __nanosleep(20); }
                s_done[n & 1] = 0;
            }
            __syncwarp();
#ifdef FMLP_ARQ_TRACE
            const unsigned long long t_done = gtime();
#endif
            if (ci.kind != kItemFinal) {
                const int phase = ci.kind == kItemFold ? 0 : 1;
                const uint32_t need = (uint32_t)(phase == 0 ? a.FJ : a.RJ);
                int last = 0;
                if (lane == 0) {
                    uint32_t old;
                    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(my_flags + q_count(phase, ci.c)) : "memory");
                    last = (old + 1u == epoch * need);
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                __syncwarp();   // lane 0's acquire is ordered before the other lanes' release stores
                if (last && lane < a.G) q_st_release_sys(a.flags[lane] + q_flag(phase, ci.c, a.rank), epoch);
            }
            // ---- an item whose inputs are not there yet is only handed over once they are (never before the
            //      previous item has been signalled: no circular waits between ranks)
#ifdef FMLP_ARQ_TRACE
            if (lane == 0) {
                const unsigned int slot = atomicAdd(&g_trace_n, 1u);
                if (slot < kTraceMax) {
                    g_trace[slot][0] = (unsigned long long)cur | ((unsigned long long)ci.kind << 32) | ((unsigned long long)ci.c << 40) | ((unsigned long long)blockIdx.x << 48);
                    g_trace[slot][1] = t_pub_cur; g_trace[slot][2] = t_done; g_trace[slot][3] = gtime();
                }
            }
#endif
            if (!published) {
                if (ni.kind == kItemReduce) wait_flags(0, ni.c);
                else for (int c = 0; c < a.NC; ++c) wait_flags(1, c);
                publish(n + 1, nxt);
#ifdef FMLP_ARQ_TRACE
                t_pub_nxt = gtime();
#endif
            }
#ifdef FMLP_ARQ_TRACE
            t_pub_cur = t_pub_nxt;
#endif
            cur = nxt; ci = ni; ++n;
        }
        if (lane == 0) {
            const uint32_t old = atomicAdd(my_flags + kQDone, 1u);
            if (old + 1u == epoch * gridDim.x) {   // every CTA of this rank has read the epoch and is done
                __threadfence();
                *a.epoch_dev = epoch;
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- compute warps
    const int tid = threadIdx.x;
    float* const my_partial = a.partial[a.rank];
    for (int n = 0;; ++n) {
        if (lane == 0) { while (q_ld_acquire_cta(&s_ready[n & 1]) != n + 1) { __nanosleep(40); } }
        __syncwarp();
        const Item it = decode_item(a, (uint32_t)s_item[n & 1]);
        if (it.kind == kItemExit) break;
        if (it.kind == kItemFold) {
            // ---- fold piece j of chunk c into this rank's partial buffer (local stores)
            const int64_t v0 = (int64_t)it.c * a.Vc + (int64_t)it.j * a.Vf;
            int64_t v1 = v0 + a.Vf;
            const int64_t vend = min((int64_t)(it.c + 1) * a.Vc, a.V);
            if (v1 > vend) v1 = vend;
            for (int64_t v = v0 + tid; v < v1; v += kQThreads) {
                const int64_t e = v << 2;
                const bool tail = e >= a.P;
                const int64_t off = tail ? e - a.P : e;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                int i = 0;
                for (; i + kQUnroll <= a.K; i += kQUnroll) {
                    float4 x[kQUnroll];
#pragma unroll
                    for (int u = 0; u < kQUnroll; ++u) x[u] = ld_stream_f4((tail ? a.src1[i + u] : a.src0[i + u]) + off);
#pragma unroll
                    for (int u = 0; u < kQUnroll; ++u) {
                        acc.x = fmaf(x[u].x, a.w[i + u], acc.x); acc.y = fmaf(x[u].y, a.w[i + u], acc.y);
                        acc.z = fmaf(x[u].z, a.w[i + u], acc.z); acc.w = fmaf(x[u].w, a.w[i + u], acc.w);
                    }
                }
                for (; i < a.K; ++i) {
                    const float4 x = ld_stream_f4((tail ? a.src1[i] : a.src0[i]) + off);
                    acc.x = fmaf(x.x, a.w[i], acc.x); acc.y = fmaf(x.y, a.w[i], acc.y);
                    acc.z = fmaf(x.z, a.w[i], acc.z); acc.w = fmaf(x.w, a.w[i], acc.w);
                }
                *reinterpret_cast<float4*>(my_partial + e) = acc;
            }
            // the fp64 tail travels with the last chunk
            if (it.c == a.NC - 1 && it.j == 0 && tid < a.M)
                reinterpret_cast<double*>(my_partial + a.tail_off)[tid] = a.tail_src[tid];
        } else if (it.kind == kItemReduce) {
            // ---- piece j of my slice of chunk c: sum over ranks, broadcast to every rank
            const int64_t s0 = (int64_t)it.c * a.Vc + (int64_t)a.rank * a.Vs;
            const int64_t send = min(min(s0 + a.Vs, (int64_t)(it.c + 1) * a.Vc), a.V);
            const int64_t v0 = s0 + (int64_t)it.j * a.Vr;
            int64_t v1 = v0 + a.Vr;
            if (v1 > send) v1 = send;
            constexpr int U = 4;
            for (int64_t vb = v0 + tid; vb < v1; vb += (int64_t)kQThreads * U) {
                float4 acc[U];
                if (NVLS) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int64_t v = vb + (int64_t)u * kQThreads;
                        if (v < v1) acc[u] = mm_ld_reduce_f4(a.mc_partial + (v << 2));
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int64_t v = vb + (int64_t)u * kQThreads;
                        if (v < v1) mm_st_f4(a.mc_result + (v << 2), acc[u]);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int g = 0; g < a.G; ++g) {            // fixed rank order: deterministic
                        float4 x[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int64_t v = vb + (int64_t)u * kQThreads;
                            x[u] = v < v1 ? __ldcg(reinterpret_cast<const float4*>(a.partial[g] + (v << 2))) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) { acc[u].x += x[u].x; acc[u].y += x[u].y; acc[u].z += x[u].z; acc[u].w += x[u].w; }
                    }
                    for (int g = 0; g < a.G; ++g) {
                        int dst = a.rank + g;                  // start with the local copy, then walk the peers
                        if (dst >= a.G) dst -= a.G;
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int64_t v = vb + (int64_t)u * kQThreads;
                            if (v < v1) *reinterpret_cast<float4*>(a.result[dst] + (v << 2)) = acc[u];
                        }
                    }
                }
            }
            // fp64 tail: rank 0 sums and broadcasts it with the last chunk
            if (it.c == a.NC - 1 && it.j == 0 && a.rank == 0 && tid < a.M) {
                if (NVLS) {
                    const double s = mm_ld_reduce_f64(reinterpret_cast<const double*>(a.mc_partial + a.tail_off) + tid);
                    mm_st_f64(reinterpret_cast<double*>(a.mc_result + a.tail_off) + tid, s);
                } else {
                    double s = 0.0;
                    for (int g = 0; g < a.G; ++g) s += __ldcg(reinterpret_cast<const double*>(a.partial[g] + a.tail_off) + tid);
                    for (int g = 0; g < a.G; ++g) reinterpret_cast<double*>(a.result[g] + a.tail_off)[tid] = s;
                }
            }
        }
        // "this warp has issued all its stores of the item": no fence, no wait — the signal warp
        // makes them visible (release.cta here, acq_rel.gpu / release.sys there)
        __syncwarp();
        if (lane == 0) q_red_release_cta(&s_done[n & 1], 1);
    }
}

// ---- the small tails of the aggregation (main.py:218-234) ------------------------------------
// This rank's fp64 partial sums, M = 3C + J:
//   [0, C)     sum of n_k over the local clients that ANNOTATE class c          (FedAvg_proto divisor, FedAvg.py:72-93)
//   [C, 2C)    sum of n_k * t_k[c] over the local clients for which c is MISSING (FedAvg_tao numerator, :51-70)
//   [2C, 3C)   sum of n_k over those clients                                     (FedAvg_tao divisor)
//   [3C, 3C+J) sum of n_k * counter_k[j]  (int64 BatchNorm counters, exact in fp64 below 2^53; FedAvg.py:9-13)
struct TailArgs {
    double w[FMLP_MAX_CLIENTS];
    int64_t rows[FMLP_MAX_CLIENTS];
    uint32_t act[FMLP_MAX_CLIENTS];
    uint32_t neg[FMLP_MAX_CLIENTS];
    const int64_t* counters[FMLP_MAX_CLIENTS];
    int S, C, J;
};

__global__ void __launch_bounds__(256) agg_tail_pack_kernel(const int32_t* __restrict__ tcnt, const __grid_constant__ TailArgs a,
                                                            double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int C = a.C;
    if (i < C) {
        double s = 0.0;
        for (int k = 0; k < a.S; ++k) if ((a.act[k] >> i) & 1u) s += a.w[k];
        out[i] = s;
    } else if (i < 2 * C) {
        const int c = i - C;
        double s = 0.0;
        // t_k[c] = count / N_k as numpy float64 (local_training.py:1000,1249), times the client weight; the
        // counts are loaded eight clients at a time (independent loads), the sum stays in client order
        for (int k0 = 0; k0 < a.S; k0 += 8) {
            int cnt[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) cnt[u] = (tcnt && k0 + u < a.S) ? tcnt[(k0 + u) * C + c] : 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int k = k0 + u;
                if (k < a.S && ((a.neg[k] >> c) & 1u)) s += ((double)cnt[u] / (double)a.rows[k]) * a.w[k];
            }
        }
        out[i] = s;
    } else if (i < 3 * C) {
        const int c = i - 2 * C;
        double s = 0.0;
        for (int k = 0; k < a.S; ++k) if ((a.neg[k] >> c) & 1u) s += a.w[k];
        out[i] = s;
    } else if (i < 3 * C + a.J) {
        const int j = i - 3 * C;
        double s = 0.0;
        for (int k0 = 0; k0 < a.S; k0 += 8) {
            long long v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (k0 + u < a.S) ? a.counters[k0 + u][j] : 0ll;
#pragma unroll
            for (int u = 0; u < 8; ++u) if (k0 + u < a.S) s += (double)v[u] * a.w[k0 + u];
        }
        out[i] = s;
    }
}

// After the all-reduce: prototypes = (sum_k w_k P_k) * (N / class weight), 0/0 -> NaN like the reference;
// tao = num / den (1.0 where nobody misses the class); counters = (float)acc / (float)N.
__global__ void __launch_bounds__(256) agg_finalize_kernel(const float* __restrict__ proto_sum, const double* __restrict__ tail,
                                                           int C, int D, int J, double total, float* __restrict__ proto_out,
                                                           double* __restrict__ tao_out, float* __restrict__ counters_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)2 * C * D;
    if (proto_out && i < n) {
        const int c = (int)(i / (2 * (int64_t)D));
        const float scale = (float)(total / tail[c]);             // inf when nobody annotates the class
        proto_out[i] = proto_sum[i] * scale;                      // 0 * inf = NaN, as 0/0 in FedAvg.py:85-86
    }
    if (tao_out && i < C) {
        const double den = tail[2 * C + i];
        tao_out[i] = den > 0.0 ? tail[C + i] / den : 1.0;
    }
    if (counters_out && i < J) counters_out[i] = (float)(int64_t)tail[3 * C + i] / (float)total;
}

}  // namespace fmlp

using namespace fmlp;

extern "C" size_t fmlp_fedavg_allreduce_q_buffer_floats(int64_t P, int64_t T, int M) {
    if (P < 0 || T < 0 || M < 0 || (P & 3) || (T & 3)) return 0;
    return (size_t)(P + T) + 2 * (size_t)((M + 1) & ~1);      // fp64 tail after the fp32 vector (16-byte aligned)
}

extern "C" int fmlp_fedavg_allreduce_q_f32(const float* const* srcs, const float* const* tail_srcs, const float* weights,
                                           int K, int64_t P, int64_t T, const double* tail_f64, int M,
                                           float* const* partial_ptrs, float* const* result_ptrs,
                                           uint32_t* const* flag_ptrs, float* mc_partial, float* mc_result,
                                           int n_chunks, int rank, int world, uint32_t* epoch_dev, int max_ctas,
                                           int fold_iters_arg, int red_iters_arg, fmlp_stream_t stream) {
    if (!srcs || !weights || !partial_ptrs || !result_ptrs || !flag_ptrs || !epoch_dev || K < 1 || K > FMLP_MAX_CLIENTS ||
        P < 0 || T < 0 || M < 0 || M > kQThreads || world < 1 || world > kQMaxRanks || rank < 0 || rank >= world ||
        n_chunks < 1 || n_chunks > kQMaxChunks)
        return FMLP_ERR_BAD_ARG;
    if ((P & 3) || (T & 3) || P + T == 0) return FMLP_ERR_UNSUPPORTED;
    if ((T > 0 && !tail_srcs) || (M > 0 && !tail_f64)) return FMLP_ERR_BAD_ARG;
    if ((mc_partial == nullptr) != (mc_result == nullptr)) return FMLP_ERR_BAD_ARG;
    QArgs a;
    for (int i = 0; i < FMLP_MAX_CLIENTS; ++i) {
        a.src0[i] = i < K ? srcs[i] : nullptr;
        a.src1[i] = (i < K && T > 0) ? tail_srcs[i] : nullptr;
        a.w[i] = i < K ? weights[i] : 0.f;
        if (i < K && (!srcs[i] || !aligned16(srcs[i]))) return FMLP_ERR_UNSUPPORTED;
        if (i < K && T > 0 && (!tail_srcs[i] || !aligned16(tail_srcs[i]))) return FMLP_ERR_UNSUPPORTED;
    }
    for (int g = 0; g < kQMaxRanks; ++g) {
        a.partial[g] = g < world ? partial_ptrs[g] : nullptr;
        a.result[g] = g < world ? result_ptrs[g] : nullptr;
        a.flags[g] = g < world ? flag_ptrs[g] : nullptr;
        if (g < world && (!partial_ptrs[g] || !result_ptrs[g] || !flag_ptrs[g] || !aligned16(partial_ptrs[g]) || !aligned16(result_ptrs[g])))
            return FMLP_ERR_BAD_ARG;
    }
    const int sms = sm_count();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    int blocks = sms * kQCtasPerSm;
    if (max_ctas > 0 && max_ctas < blocks) blocks = max_ctas;
    a.mc_partial = mc_partial; a.mc_result = mc_result;
    a.tail_src = tail_f64; a.M = M;
    a.P = P; a.T = T; a.V = (P + T) >> 2;
    a.tail_off = P + T;
    a.K = K; a.rank = rank; a.G = world; a.NC = n_chunks; a.epoch_dev = epoch_dev;
    // chunk = G slices; slices and pieces are whole float4 counts
    const int64_t per_slice = (a.V + (int64_t)n_chunks * world - 1) / ((int64_t)n_chunks * world);
    a.Vs = per_slice < 1 ? 1 : per_slice;
    a.Vc = a.Vs * world;
    // Pieces are whole multiples of the CTA's thread count (a ragged last iteration would idle 31 of 32 warps).
    //   fold pieces: fold_iters float4 per thread (x K loads each): fine-grained, the scheduler warp prefetches
    //   reduce pieces: red_iters float4 per thread: the NVLink round trip is ~3 us, so the phase is latency-bound
    //   and wants EVERY element of the slice in flight at once -> as many pieces (CTAs) as the slice has work for
    static int fold_iters = 0, red_iters = 0;
    if (fold_iters == 0) {
        fold_iters = 2; red_iters = 2;
        if (const char* e = getenv("FMLP_ARQ_FOLD_ITERS")) { int v = atoi(e); if (v >= 1 && v <= 64) fold_iters = v; }
        if (const char* e = getenv("FMLP_ARQ_RED_ITERS")) { int v = atoi(e); if (v >= 1 && v <= 64) red_iters = v; }
    }
    const int fi = (fold_iters_arg >= 1 && fold_iters_arg <= 64) ? fold_iters_arg : fold_iters;
    const int ri = (red_iters_arg >= 1 && red_iters_arg <= 64) ? red_iters_arg : red_iters;
    a.Vf = (int64_t)kQThreads * fi;
    a.FJ = (int)((a.Vc + a.Vf - 1) / a.Vf);
    a.Vr = (int64_t)kQThreads * ri;
    a.RJ = (int)((a.Vs + a.Vr - 1) / a.Vr);
    if (a.FJ < 1) a.FJ = 1;
    if (a.RJ < 1) a.RJ = 1;
    if (mc_partial)
        fedavg_allreduce_q_kernel<true><<<blocks, kQThreads + 32, 0, (cudaStream_t)stream>>>(a);
    else
        fedavg_allreduce_q_kernel<false><<<blocks, kQThreads + 32, 0, (cudaStream_t)stream>>>(a);
    return launch_status();
}

extern "C" int fmlp_agg_tail_pack_f64(const int32_t* tcnt, int S, int C, const double* weights, const int64_t* rows,
                                      const uint32_t* act, const uint32_t* neg, const int64_t* const* counters, int J,
                                      double* out, fmlp_stream_t stream) {
    if (!weights || !rows || !act || !neg || !out || S < 1 || S > FMLP_MAX_CLIENTS || C < 1 || C > FMLP_MAX_CLASSES || J < 0)
        return FMLP_ERR_BAD_ARG;
    if (J > 0 && !counters) return FMLP_ERR_BAD_ARG;
    TailArgs a;
    for (int k = 0; k < FMLP_MAX_CLIENTS; ++k) {
        a.w[k] = k < S ? weights[k] : 0.0;
        a.rows[k] = k < S ? rows[k] : 1;
        a.act[k] = k < S ? act[k] : 0u;
        a.neg[k] = k < S ? neg[k] : 0u;
        a.counters[k] = (k < S && J > 0) ? counters[k] : nullptr;
        if (k < S && rows[k] < 1) a.rows[k] = 1;
        if (k < S && J > 0 && !counters[k]) return FMLP_ERR_BAD_ARG;
    }
    a.S = S; a.C = C; a.J = J;
    const int n = 3 * C + J;
    agg_tail_pack_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(tcnt, a, out);
    return launch_status();
}

extern "C" int fmlp_agg_finalize_f32(const float* proto_sum, const double* tail, int C, int D, int J, double total_weight,
                                     float* proto_out, double* tao_out, float* counters_out, fmlp_stream_t stream) {
    if (!tail || C < 1 || C > FMLP_MAX_CLASSES || D < 0 || J < 0 || !(total_weight > 0.0)) return FMLP_ERR_BAD_ARG;
    if (proto_out && (!proto_sum || D < 1)) return FMLP_ERR_BAD_ARG;
    int64_t n = proto_out ? (int64_t)2 * C * D : 0;
    if (n < C) n = C;
    if (n < J) n = J;
    agg_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(proto_sum, tail, C, D, J, total_weight,
                                                                                      proto_out, tao_out, counters_out);
    return launch_status();
}

#ifdef FMLP_ARQ_TRACE
// debug only: copies the trace to the host (synchronises) and resets it; returns the number of records
extern "C" int fmlp_arq_trace_dump(unsigned long long* out, int max_records) {
    unsigned int n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(n));
    if ((int)n > max_records) n = max_records;
    if (n > (unsigned)kTraceMax) n = kTraceMax;
    cudaMemcpyFromSymbol(out, g_trace, (size_t)n * 4 * sizeof(unsigned long long));
    const unsigned int zero = 0;
    cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(zero));
    return (int)n;
}
#endif
