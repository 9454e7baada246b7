// common.cuh — shared device/host helpers for libfedmlp_b200 (sm_100a only).
#pragma once

#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fedmlp_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libfedmlp_b200 is written for sm_100a (B200); build with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace fmlp {

constexpr int kWarp = 32;

// Grid sizing is derived from the SM count of the current device (148 on B200), cached.
inline int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        cached[dev] = n;
    }
    return cached[dev];
}

// Process-wide count of kernel launches issued by this library (bench.py reports it).
extern unsigned long long g_launch_count;
extern int g_tuning[FMLP_TUNE_COUNT];
// value of a tuning knob: fmlp_set_tuning() first, then the environment variable, then the default
inline int tuning_value(int knob, const char* env, int lo, int hi, int dflt) {
    int v = __atomic_load_n(&g_tuning[knob], __ATOMIC_RELAXED);
    if (v < 0) { const char* e = getenv(env); v = e ? atoi(e) : dflt; }
    return (v < lo || v > hi) ? dflt : v;
}

// Launch epilogue: surface launch-configuration errors as a positive cudaError_t.
inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) __atomic_fetch_add(&g_launch_count, 1ull, __ATOMIC_RELAXED);
    return e == cudaSuccess ? FMLP_OK : (int)e;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- streaming 128-bit global accesses -------------------------------------------------
// Inputs on this path are read exactly once per launch: bypass L1 allocation so the 256 KB
// of L1/smem per SM is not churned.
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// ---- warp / block reductions --------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reference-faithful sigmoid: torch computes 1/(1+exp(-x)) in fp32 (no fast-math).
__device__ __forceinline__ float sigmoid_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

// Segment lookup: largest s with seg_rows[s] <= row (seg_rows is a prefix array of S+1 entries
// living in the kernel parameter block).
__device__ __forceinline__ int find_segment(const int64_t* seg_rows, int S, int64_t row) {
    int lo = 0, hi = S - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (seg_rows[mid] <= row) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Segment table carried by value in kernel parameters.
struct SegTable {
    int64_t rows[FMLP_MAX_SEGMENTS + 1];
    uint32_t mask_a[FMLP_MAX_SEGMENTS];  // meaning depends on the kernel (active / missing)
    uint32_t mask_b[FMLP_MAX_SEGMENTS];
    int S;
};

inline int fill_seg_table(SegTable& t, int S, const int64_t* seg_rows, const uint32_t* a,
                          const uint32_t* b) {
    if (S < 1 || S > FMLP_MAX_SEGMENTS || seg_rows == nullptr) return FMLP_ERR_BAD_ARG;
    t.S = S;
    for (int s = 0; s <= S; ++s) {
        t.rows[s] = seg_rows[s];
        if (s > 0 && seg_rows[s] < seg_rows[s - 1]) return FMLP_ERR_BAD_ARG;
    }
    if (seg_rows[0] != 0) return FMLP_ERR_BAD_ARG;
    for (int s = 0; s < S; ++s) {
        t.mask_a[s] = a ? a[s] : 0u;
        t.mask_b[s] = b ? b[s] : 0u;
    }
    for (int s = S; s < FMLP_MAX_SEGMENTS; ++s) { t.mask_a[s] = 0; t.mask_b[s] = 0; t.rows[s + 1] = seg_rows[S]; }
    return FMLP_OK;
}

}  // namespace fmlp
