// exp_peer_signal.cu — microbenchmark for round 2: what does it cost to tell a peer GPU "my stores have
// landed"?  (Standalone: one process, two GPUs with peer access, no torch.  NOT part of the library.)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o exp_peer_signal tools/exp_peer_signal.cu
//   ./exp_peer_signal [MB=28] [chunks=4] [iters=50]          (needs 2 GPUs: gpurun --gpus 2)
//
// Background (profiles/r01_multi_sweep.txt): the fused fold + all-reduce kernel pushes partial
// slices to its peers with plain st.global and signals per chunk.  With the fences compiled out the stage
// takes 65 us at 2 GPUs; a MEMBAR.SYS per CTA per chunk makes it 145 us, a MEMBAR.GPU per CTA + one
// MEMBAR.SYS per chunk 88 us.  A membar stalls on the SM's whole outstanding write stream.  This program
// measures candidate replacements in isolation.  Each GPU streams `MB` of local data to the other GPU in
// `chunks` chunks; a receiver CTA group on the other GPU waits for each chunk's signal and CHECKS the data
// (every word carries the iteration number), so a scheme that signals too early is caught, not just timed.
//
//   mode 0  st.global.v4, no ordering at all                      (lower bound, expected to FAIL the check sometimes)
//   mode 1  st.global.v4, __threadfence_system in every thread, CTA counter, last CTA st.release.sys
//   mode 2  st.global.v4, one signal thread per CTA: fence.acq_rel.gpu + atomic; last CTA st.release.sys   (shipped in r01)
//   mode 3  TMA bulk store smem -> peer (cp.async.bulk.global.shared::cta), wait_group 0, then one
//           red.release.sys.add on the peer's chunk counter per CTA (no fence with stores in flight on the LSU path)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kThreads = 256;
constexpr int kTileFloats = 4096;   // 16 KB staged per bulk store (mode 3)
constexpr int kMaxChunks = 32;

struct Args {
    const float* src;        // local [n]
    float* peer_dst;         // peer  [n]
    const float* my_dst;     // local [n], written by the peer
    uint32_t* peer_flags;    // peer  [kMaxChunks] epoch flags / counters
    uint32_t* my_flags;      // local [kMaxChunks]
    uint32_t* local_cnt;     // local [kMaxChunks] CTA counters (modes 1, 2)
    unsigned long long* bad; // local: mismatching words seen by the receivers
    int64_t n;               // floats, multiple of 4 * chunks
    int chunks;
    int senders;             // CTAs [0, senders) send, the rest receive
    uint32_t epoch;          // iteration number (1-based)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release_sys_add(uint32_t* p, uint32_t v) { asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(kThreads) push_kernel(Args a) {
    __shared__ __align__(128) float tile[MODE == 3 ? kTileFloats : 4];
    const int64_t chunk_len = a.n / a.chunks;
    if ((int)blockIdx.x < a.senders) {
        // ------------------------------------------------------------------ sender
        for (int c = 0; c < a.chunks; ++c) {
            const float* src = a.src + c * chunk_len;
            float* dst = a.peer_dst + c * chunk_len;
            if (MODE != 3) {
                for (int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4; i < chunk_len; i += (int64_t)a.senders * kThreads * 4) {
                    float4 v = *reinterpret_cast<const float4*>(src + i);
                    v.x += (float)a.epoch; v.y += (float)a.epoch; v.z += (float)a.epoch; v.w += (float)a.epoch;
                    *reinterpret_cast<float4*>(dst + i) = v;
                }
                if (MODE == 0) {
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        const uint32_t old = atomicAdd(a.local_cnt + c, 1u);
                        if (old + 1u == a.epoch * a.senders) *reinterpret_cast<volatile uint32_t*>(a.peer_flags + c) = a.epoch;
                    }
                } else if (MODE == 1) {
                    __threadfence_system();
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        const uint32_t old = atomicAdd(a.local_cnt + c, 1u);
                        if (old + 1u == a.epoch * a.senders) { __threadfence_system(); st_release_sys(a.peer_flags + c, a.epoch); }
                    }
                } else {   // MODE 2
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        uint32_t old;
                        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(a.local_cnt + c) : "memory");
                        if (old + 1u == a.epoch * a.senders) st_release_sys(a.peer_flags + c, a.epoch);
                    }
                }
            } else {
                // tiles of 16 KB: global -> registers (+epoch) -> smem -> bulk store to the peer
                const int64_t n_tiles = (chunk_len + kTileFloats - 1) / kTileFloats;
                for (int64_t t = blockIdx.x; t < n_tiles; t += a.senders) {
                    const int64_t base = t * kTileFloats;
                    const int len = (int)min((int64_t)kTileFloats, chunk_len - base);
                    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous tile left smem
                    __syncthreads();
                    for (int i = threadIdx.x * 4; i < len; i += kThreads * 4) {
                        float4 v = *reinterpret_cast<const float4*>(src + base + i);
                        v.x += (float)a.epoch; v.y += (float)a.epoch; v.z += (float)a.epoch; v.w += (float)a.epoch;
                        *reinterpret_cast<float4*>(tile + i) = v;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + base), "r"(smem_u32(tile)), "r"(len * 4) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                if (threadIdx.x == 0) {
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all my bulk stores are complete
                    asm volatile("fence.proxy.async;" ::: "memory");
                    red_release_sys_add(a.peer_flags + c, 1u);                    // one remote count per CTA per chunk
                }
                __syncthreads();
            }
        }
    } else {
        // ------------------------------------------------------------------ receiver: wait per chunk, then verify
        const int r = blockIdx.x - a.senders, nr = gridDim.x - a.senders;
        unsigned long long bad = 0;
        for (int c = 0; c < a.chunks; ++c) {
            const uint32_t want = (MODE == 3) ? a.epoch * (uint32_t)a.senders : a.epoch;
            if (threadIdx.x == 0)
                while ((int32_t)(ld_acquire_sys(a.my_flags + c) - want) < 0) { __nanosleep(32); }
            __syncthreads();
            const float* got = a.my_dst + c * chunk_len;
            for (int64_t i = ((int64_t)r * kThreads + threadIdx.x) * 4; i < chunk_len; i += (int64_t)nr * kThreads * 4) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(got + i));
                const float e = (float)a.epoch + (float)((c * chunk_len + i) & 1023);
                bad += (v.x != e) + (v.y != e + 1.f) + (v.z != e + 2.f) + (v.w != e + 3.f);
            }
        }
        if (bad) atomicAdd(a.bad, bad);
    }
}

__global__ void init_kernel(float* p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (float)(i & 1023);
}

template <int MODE>
static void run_mode(int iters, int64_t n, int chunks, float* src[2], float* dst[2], uint32_t* flags[2], uint32_t* cnt[2],
                     unsigned long long* bad[2], cudaStream_t st[2], int sms) {
    const int senders = sms, receivers = sms / 4;   // all co-resident: 1.25 CTAs of 256 threads per SM
    cudaEvent_t e0[2], e1[2];
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaMemsetAsync(flags[d], 0, kMaxChunks * 4, st[d])); CK(cudaMemsetAsync(cnt[d], 0, kMaxChunks * 4, st[d]));
        CK(cudaMemsetAsync(bad[d], 0, 8, st[d])); CK(cudaMemsetAsync(dst[d], 0, n * 4, st[d]));
        CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d]));
    }
    for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); }
    float total_ms = 0.f;
    for (int it = 1; it <= iters + 3; ++it) {
        for (int d = 0; d < 2; ++d) {
            CK(cudaSetDevice(d));
            Args a{src[d], dst[1 - d], dst[d], flags[1 - d], flags[d], cnt[d], bad[d], n, chunks, senders, (uint32_t)it};
            CK(cudaEventRecord(e0[d], st[d]));
            push_kernel<MODE><<<senders + receivers, kThreads, 0, st[d]>>>(a);
            CK(cudaEventRecord(e1[d], st[d]));
        }
        float worst = 0.f;
        for (int d = 0; d < 2; ++d) {
            CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d]));
            float ms; CK(cudaEventElapsedTime(&ms, e0[d], e1[d])); worst = ms > worst ? ms : worst;
        }
        if (it > 3) total_ms += worst;
    }
    unsigned long long h[2];
    for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaMemcpy(&h[d], bad[d], 8, cudaMemcpyDeviceToHost)); }
    const double us = total_ms / iters * 1e3;
    printf("mode %d: %8.1f us per exchange  %7.1f GB/s per direction  mismatching words: %llu + %llu\n", MODE, us,
           n * 4.0 / us / 1e3, h[0], h[1]);
}

int main(int argc, char** argv) {
    const double mb = argc > 1 ? atof(argv[1]) : 28.0;
    const int chunks = argc > 2 ? atoi(argv[2]) : 4;
    const int iters = argc > 3 ? atoi(argv[3]) : 50;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("needs 2 GPUs\n"); return 0; }
    if (chunks < 1 || chunks > kMaxChunks) { printf("chunks 1..%d\n", kMaxChunks); return 1; }
    int64_t n = (int64_t)(mb * 1e6 / 4);
    n = n / (1024 * chunks) * (1024 * chunks);   // chunk boundaries keep the (i & 1023) pattern and 16 KB tiles aligned
    float *src[2], *dst[2];
    uint32_t *flags[2], *cnt[2];
    unsigned long long* bad[2];
    cudaStream_t st[2];
    int sms = 0;
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        int can = 0; CK(cudaDeviceCanAccessPeer(&can, d, 1 - d));
        if (!can) { printf("no peer access %d -> %d\n", d, 1 - d); return 0; }
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d));
        CK(cudaMalloc(&src[d], n * 4)); CK(cudaMalloc(&dst[d], n * 4));
        CK(cudaMalloc(&flags[d], kMaxChunks * 4)); CK(cudaMalloc(&cnt[d], kMaxChunks * 4)); CK(cudaMalloc(&bad[d], 8));
        CK(cudaStreamCreate(&st[d]));
        init_kernel<<<1024, 256, 0, st[d]>>>(src[d], n);
        CK(cudaStreamSynchronize(st[d]));
    }
    printf("%.1f MB per direction, %d chunks, %d iterations, %d SMs\n", n * 4 / 1e6, chunks, iters, sms);
    run_mode<0>(iters, n, chunks, src, dst, flags, cnt, bad, st, sms);
    run_mode<1>(iters, n, chunks, src, dst, flags, cnt, bad, st, sms);
    run_mode<2>(iters, n, chunks, src, dst, flags, cnt, bad, st, sms);
    run_mode<3>(iters, n, chunks, src, dst, flags, cnt, bad, st, sms);
    return 0;
}
