// Read-bandwidth experiment: which access shape saturates B200 HBM for a [N, D] fp32 matrix?
//   mode 0: warp-per-rows: R rows x 512 B chunk per batch, walk D (the tag_sim shape), ping-pong
//   mode 1: warp-per-row:  1 row, R consecutive 512 B chunks per batch (contiguous 2 KB)
//   mode 2: CTA-per-rows:  thread <-> float4 column, U rows per batch (the proto shape)
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ float4 ldg(const float* p) {
    float4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v;
}
template <int R, bool PINGPONG>
__global__ void k_mode0(const float* __restrict__ f, long n, int D, float* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long ntiles = n / R;
    float acc = 0.f;
    for (long t = (long)blockIdx.x * nw + warp; t < ntiles; t += (long)gridDim.x * nw) {
        const float* base = f + t * R * (long)D + lane * 4;
        float4 a[R], b[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = ldg(base + (long)r * D);
        for (int c = 0; c < D / 128; c += 2) {
#pragma unroll
            for (int r = 0; r < R; ++r) b[r] = ldg(base + (long)r * D + (c + 1) * 128);
            if (!PINGPONG) { /* consume b immediately below as well */ }
#pragma unroll
            for (int r = 0; r < R; ++r) acc += a[r].x + a[r].y + a[r].z + a[r].w;
            if (c + 2 < D / 128) {
#pragma unroll
                for (int r = 0; r < R; ++r) a[r] = ldg(base + (long)r * D + (c + 2) * 128);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) acc += b[r].x + b[r].y + b[r].z + b[r].w;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}
__device__ __forceinline__ void prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}
// mode 3: mode0 + bulk L2 prefetch of the warp's tile `AHEAD` iterations ahead
template <int R, int AHEAD>
__global__ void k_mode3(const float* __restrict__ f, long n, int D, float* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long ntiles = n / R;
    const long stride = (long)gridDim.x * nw;
    float acc = 0.f;
    long t0 = (long)blockIdx.x * nw + warp;
    if (lane == 0) for (int k = 1; k < AHEAD; ++k) if (t0 + k * stride < ntiles) prefetch_l2(f + (t0 + k * stride) * R * (long)D, R * D * 4);
    for (long t = t0; t < ntiles; t += stride) {
        if (lane == 0 && t + AHEAD * stride < ntiles) prefetch_l2(f + (t + AHEAD * stride) * R * (long)D, R * D * 4);
        const float* base = f + t * R * (long)D + lane * 4;
        float4 a[R], b[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = ldg(base + (long)r * D);
        for (int c = 0; c < D / 128; c += 2) {
#pragma unroll
            for (int r = 0; r < R; ++r) b[r] = ldg(base + (long)r * D + (c + 1) * 128);
#pragma unroll
            for (int r = 0; r < R; ++r) acc += a[r].x + a[r].y + a[r].z + a[r].w;
            if (c + 2 < D / 128) {
#pragma unroll
                for (int r = 0; r < R; ++r) a[r] = ldg(base + (long)r * D + (c + 2) * 128);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) acc += b[r].x + b[r].y + b[r].z + b[r].w;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}
// mode 4: mode0 addressing, but the loads are cp.async (LDGSTS) into a per-warp smem ring of
// DEPTH batches (each lane later reads back only the 16 B it copied itself -> no cross-lane sync)
template <int R, int DEPTH>
__global__ void k_mode4(const float* __restrict__ f, long n, int D, float* out) {
    extern __shared__ __align__(16) float ring[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long ntiles = n / R;
    const long stride = (long)gridDim.x * nw;
    const int nch = D / 128;
    float* my = ring + (size_t)warp * DEPTH * R * 128 + lane * 4;
    const unsigned my_s = (unsigned)__cvta_generic_to_shared(my);
    float acc = 0.f;
    // flattened sequence of (tile, chunk) batches for this warp
    long t_issue = (long)blockIdx.x * nw + warp; int c_issue = 0; int slot_issue = 0;
    auto issue = [&]() {
        if (t_issue < ntiles) {
            const float* base = f + t_issue * R * (long)D + lane * 4 + c_issue * 128;
#pragma unroll
            for (int r = 0; r < R; ++r)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(my_s + (unsigned)((slot_issue * R + r) * 512)), "l"(base + (long)r * D) : "memory");
            if (++c_issue == nch) { c_issue = 0; t_issue += stride; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        slot_issue = (slot_issue + 1 == DEPTH) ? 0 : slot_issue + 1;
    };
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) issue();
    int slot = 0;
    for (long t = (long)blockIdx.x * nw + warp; t < ntiles; t += stride) {
        for (int c = 0; c < nch; ++c) {
            issue();
            asm volatile("cp.async.wait_group %0;" :: "n"(DEPTH - 1) : "memory");
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float4 v = *reinterpret_cast<const float4*>(my + (slot * R + r) * 128);
                acc += v.x + v.y + v.z + v.w;
            }
            slot = (slot + 1 == DEPTH) ? 0 : slot + 1;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}
template <int R>
__global__ void k_mode1(const float* __restrict__ f, long n, int D, float* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float acc = 0.f;
    for (long row = (long)blockIdx.x * nw + warp; row < n; row += (long)gridDim.x * nw) {
        const float* base = f + row * (long)D + lane * 4;
        for (int c = 0; c < D / 128; c += R) {
            float4 a[R];
#pragma unroll
            for (int r = 0; r < R; ++r) a[r] = ldg(base + (c + r) * 128);
#pragma unroll
            for (int r = 0; r < R; ++r) acc += a[r].x + a[r].y + a[r].z + a[r].w;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}
template <int U>
__global__ void k_mode2(const float* __restrict__ f, long n, int D, float* out) {
    const int col = threadIdx.x * 4;
    float acc = 0.f;
    const long rows_per = (n + gridDim.x - 1) / gridDim.x;
    const long r0 = blockIdx.x * rows_per, r1 = min(n, r0 + rows_per);
    for (long r = r0; r < r1; r += U) {
        float4 a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = (r + u < r1) ? ldg(f + (r + u) * (long)D + col) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += a[u].x + a[u].y + a[u].z + a[u].w;
    }
    if (acc == 123.456f) out[0] = acc;
}
// mode 5: CTA-per-rows like mode 2, but row blocks are INTERLEAVED over the CTAs (CTA b takes
// blocks b, b+G, ...), so the whole chip sweeps one linear front (the FedAvg access shape)
template <int U>
__global__ void k_mode5(const float* __restrict__ f, long n, int D, float* out) {
    const int col = threadIdx.x * 4;
    float acc = 0.f;
    for (long r = (long)blockIdx.x * U; r < n; r += (long)gridDim.x * U) {
        float4 a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = (r + u < n) ? ldg(f + (r + u) * (long)D + col) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += a[u].x + a[u].y + a[u].z + a[u].w;
    }
    if (acc == 123.456f) out[0] = acc;
}
// mode 6: flat grid-stride float4 read (the fedavg_flat shape with K = 1 array)
template <int U>
__global__ void k_mode6(const float* __restrict__ f, long nvec, float* out) {
    float acc = 0.f;
    const long nt = (long)gridDim.x * blockDim.x;
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += nt * U) {
        float4 a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = (v + u * nt < nvec) ? ldg(f + (v + u * nt) * 4) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += a[u].x + a[u].y + a[u].z + a[u].w;
    }
    if (acc == 123.456f) out[0] = acc;
}
template <typename F> float timeit(F launch, float* flush, size_t flush_n) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<float> ts;
    for (int i = 0; i < 25; ++i) {
        cudaMemsetAsync(flush, i, flush_n);
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (i >= 5) ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end()); return ts[ts.size() / 2];
}
#include <algorithm>
int main() {
    const long n = 55000; const int D = 1024;
    float *f, *out, *flush; const size_t fb = 512ull << 20;
    cudaMalloc(&f, n * D * 4); cudaMalloc(&out, 4); cudaMalloc(&flush, fb);
    cudaMemset(f, 0, n * D * 4);
    const double bytes = (double)n * D * 4;
    auto rep = [&](const char* name, float ms) { printf("%-44s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, bytes / ms / 1e6); };
    int sms = 148;
    rep("mode0 R=4 pingpong 12 warps x1 CTA", timeit([&] { k_mode0<4, true><<<sms, 384>>>(f, n, D, out); }, flush, fb));
    rep("mode0 R=4 pingpong 16 warps x1 CTA", timeit([&] { k_mode0<4, true><<<sms, 512>>>(f, n, D, out); }, flush, fb));
    rep("mode0 R=4 pingpong 8 warps x3 CTA", timeit([&] { k_mode0<4, true><<<sms * 3, 256>>>(f, n, D, out); }, flush, fb));
    rep("mode0 R=4 pingpong 8 warps x6 CTA", timeit([&] { k_mode0<4, true><<<sms * 6, 256>>>(f, n, D, out); }, flush, fb));
    rep("mode0 R=2 pingpong 8 warps x4 CTA", timeit([&] { k_mode0<2, true><<<sms * 4, 256>>>(f, n, D, out); }, flush, fb));
    rep("mode0 R=8 pingpong 8 warps x2 CTA", timeit([&] { k_mode0<8, true><<<sms * 2, 256>>>(f, n, D, out); }, flush, fb));
    rep("mode3 R=4 12 warps x1 + L2 prefetch 1 ahead", timeit([&] { k_mode3<4, 1><<<sms, 384>>>(f, n, D, out); }, flush, fb));
    rep("mode3 R=4 12 warps x1 + L2 prefetch 2 ahead", timeit([&] { k_mode3<4, 2><<<sms, 384>>>(f, n, D, out); }, flush, fb));
    rep("mode3 R=4 12 warps x1 + L2 prefetch 3 ahead", timeit([&] { k_mode3<4, 3><<<sms, 384>>>(f, n, D, out); }, flush, fb));
    rep("mode3 R=4 8 warps x1 + L2 prefetch 2 ahead", timeit([&] { k_mode3<4, 2><<<sms, 256>>>(f, n, D, out); }, flush, fb));
    rep("mode3 R=4 8 warps x1 + L2 prefetch 4 ahead", timeit([&] { k_mode3<4, 4><<<sms, 256>>>(f, n, D, out); }, flush, fb));
    rep("mode3 R=2 8 warps x1 + L2 prefetch 4 ahead", timeit([&] { k_mode3<2, 4><<<sms, 256>>>(f, n, D, out); }, flush, fb));
    cudaFuncSetAttribute(k_mode4<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_mode4<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_mode4<4, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    rep("mode4 cp.async ring R=4 depth4 12 warps", timeit([&] { k_mode4<4, 4><<<sms, 384, 12 * 4 * 2048>>>(f, n, D, out); }, flush, fb));
    rep("mode4 cp.async ring R=4 depth4 16 warps", timeit([&] { k_mode4<4, 4><<<sms, 512, 16 * 4 * 2048>>>(f, n, D, out); }, flush, fb));
    rep("mode4 cp.async ring R=4 depth6 12 warps", timeit([&] { k_mode4<4, 6><<<sms, 384, 12 * 6 * 2048>>>(f, n, D, out); }, flush, fb));
    rep("mode4 cp.async ring R=4 depth6 16 warps", timeit([&] { k_mode4<4, 6><<<sms, 512, 16 * 6 * 2048>>>(f, n, D, out); }, flush, fb));
    rep("mode4 cp.async ring R=4 depth8 12 warps", timeit([&] { k_mode4<4, 8><<<sms, 384, 12 * 8 * 2048>>>(f, n, D, out); }, flush, fb));
    rep("mode1 4 chunks/row 12 warps x1", timeit([&] { k_mode1<4><<<sms, 384>>>(f, n, D, out); }, flush, fb));
    rep("mode1 8 chunks/row 12 warps x1", timeit([&] { k_mode1<8><<<sms, 384>>>(f, n, D, out); }, flush, fb));
    rep("mode1 8 chunks/row 8 warps x4", timeit([&] { k_mode1<8><<<sms * 4, 256>>>(f, n, D, out); }, flush, fb));
    rep("mode5 interleaved U=4 256thr x8 CTA/SM", timeit([&] { k_mode5<4><<<sms * 8, 256>>>(f, n, D, out); }, flush, fb));
    rep("mode5 interleaved U=8 256thr x4 CTA/SM", timeit([&] { k_mode5<8><<<sms * 4, 256>>>(f, n, D, out); }, flush, fb));
    rep("mode5 interleaved U=8 256thr x8 CTA/SM", timeit([&] { k_mode5<8><<<sms * 8, 256>>>(f, n, D, out); }, flush, fb));
    rep("mode6 flat grid-stride U=8 256thr x4 CTA/SM", timeit([&] { k_mode6<8><<<sms * 4, 256>>>(f, n * D / 4, out); }, flush, fb));
    rep("mode6 flat grid-stride U=8 256thr x8 CTA/SM", timeit([&] { k_mode6<8><<<sms * 8, 256>>>(f, n * D / 4, out); }, flush, fb));
    rep("mode6 flat grid-stride U=4 256thr x8 CTA/SM", timeit([&] { k_mode6<4><<<sms * 8, 256>>>(f, n * D / 4, out); }, flush, fb));
    rep("mode2 U=8 256thr x2 CTA/SM", timeit([&] { k_mode2<8><<<sms * 2, 256>>>(f, n, D, out); }, flush, fb));
    rep("mode2 U=8 256thr x4 CTA/SM", timeit([&] { k_mode2<8><<<sms * 4, 256>>>(f, n, D, out); }, flush, fb));
    rep("mode2 U=4 256thr x8 CTA/SM", timeit([&] { k_mode2<4><<<sms * 8, 256>>>(f, n, D, out); }, flush, fb));
    rep("mode2 U=16 256thr x2 CTA/SM", timeit([&] { k_mode2<16><<<sms * 2, 256>>>(f, n, D, out); }, flush, fb));
    return 0;
}
