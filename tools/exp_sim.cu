// exp_sim.cu — stand-alone timing + verification harness for fmlp_tag_sim_f32 (csrc/tag_sim.cu).
// Built by tools/build_exp_sim.sh in several tuning variants (-DFMLP_SIM_W / _RT_SMALL / _RT_LARGE);
// each binary checks both modes against a float64 reference kernel and prints one JSON line per case.
//   exp_sim <N> <D> <C> <segments> [iters]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include <cuda_runtime.h>

#include "fedmlp_b200.h"

namespace fmlp { unsigned long long g_launch_count = 0; int g_tuning[FMLP_TUNE_COUNT] = {-1, -1, -1, -1}; }
#ifdef FMLP_SIM_TRACE
extern "C" int fmlp_sim_trace_read(unsigned long long* host, int n);
#endif

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void fill_kernel(float* p, size_t n, uint32_t seed, int relu) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)(i * 2654435761u) ^ seed;
        x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
        float u = (float)(x & 0xffffff) / 16777216.0f;          // [0,1)
        uint32_t y = x * 747796405u + 2891336453u; y ^= y >> 13;
        float v = (float)(y & 0xffffff) / 16777216.0f;
        float g = sqrtf(-2.0f * logf(u + 1e-7f)) * cosf(6.2831853f * v);
        p[i] = relu ? fmaxf(g, 0.f) : g;
    }
}

// float64 reference: one warp per row
__global__ void ref_kernel(const float* feat, int64_t ld, int D, const float* proto, int C, int64_t N, uint32_t missing,
                           double* out /* [C][N] */) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= N) return;
    double ff = 0;
    for (int d = lane; d < D; d += 32) { double v = feat[row * ld + d]; ff += v * v; }
    for (int o = 16; o; o >>= 1) ff += __shfl_xor_sync(0xffffffffu, ff, o);
    for (int c = 0; c < C; ++c) {
        if (!((missing >> c) & 1u)) continue;
        double d0 = 0, d1 = 0, n0 = 0, n1 = 0;
        for (int d = lane; d < D; d += 32) {
            double f = feat[row * ld + d], p0 = proto[(2 * c) * (int64_t)D + d], p1 = proto[(2 * c + 1) * (int64_t)D + d];
            d0 += f * p0; d1 += f * p1; n0 += p0 * p0; n1 += p1 * p1;
        }
        for (int o = 16; o; o >>= 1) {
            d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            n0 += __shfl_xor_sync(0xffffffffu, n0, o); n1 += __shfl_xor_sync(0xffffffffu, n1, o);
        }
        if (lane == 0) out[(int64_t)c * N + row] = d0 / (sqrt(ff) * sqrt(n0)) - d1 / (sqrt(ff) * sqrt(n1));
    }
}

int main(int argc, char** argv) {
    const int64_t N = argc > 1 ? atoll(argv[1]) : 55000;
    const int D = argc > 2 ? atoi(argv[2]) : 1024;
    const int C = argc > 3 ? atoi(argv[3]) : 5;
    const int S = argc > 4 ? atoi(argv[4]) : 8;
    const int iters = argc > 5 ? atoi(argv[5]) : 30;
    float *feat, *proto, *sim;
    double* ref;
    CK(cudaMalloc(&feat, (size_t)N * D * 4));
    CK(cudaMalloc(&proto, (size_t)2 * C * D * 4));
    CK(cudaMalloc(&sim, (size_t)C * N * 4));
    CK(cudaMalloc(&ref, (size_t)C * N * 8));
    void* ws; const size_t ws_bytes = fmlp_tag_sim_ws_bytes(C, D);
    CK(cudaMalloc(&ws, ws_bytes));
    fill_kernel<<<1184, 256>>>(feat, (size_t)N * D, 1037u, 1);
    fill_kernel<<<64, 256>>>(proto, (size_t)2 * C * D, 77u, 1);
    std::vector<int64_t> rows(S + 1);
    std::vector<uint32_t> missing(S);
    for (int s = 0; s <= S; ++s) rows[s] = N * s / S;
    const uint32_t all = (C < 32 ? (1u << C) : 0u) - 1u;
    for (int s = 0; s < S; ++s) missing[s] = all & ~(1u << (s % C));
    // float64 reference for every (row, class); only the entries the kernel must write are compared
    CK(cudaMemset(ref, 0, (size_t)C * N * 8));
    ref_kernel<<<(unsigned)((N * 32 + 255) / 256), 256>>>(feat, D, D, proto, C, N, all, ref);
    CK(cudaDeviceSynchronize());
    std::vector<double> href((size_t)C * N);
    CK(cudaMemcpy(href.data(), ref, href.size() * 8, cudaMemcpyDeviceToHost));
    std::vector<float> hsim((size_t)C * N);

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int mode = 1; mode >= 0; --mode) {
        CK(cudaMemset(sim, 0xff, (size_t)C * N * 4));     // NaN pattern: unwritten entries stand out
        int rc = fmlp_tag_sim_f32(feat, D, D, proto, C, S, rows.data(), missing.data(), sim, N, mode, ws, ws_bytes, 0);
        if (rc != 0) { printf("{\"mode\": %d, \"rc\": %d}\n", mode, rc); continue; }
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hsim.data(), sim, hsim.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0; long long bad = 0, written = 0;
        for (int s = 0; s < S; ++s)
            for (int c = 0; c < C; ++c)
                for (int64_t r = rows[s]; r < rows[s + 1]; ++r) {
                    const float v = hsim[(size_t)c * N + r];
                    if ((missing[s] >> c) & 1u) {
                        ++written;
                        const double e = fabs((double)v - href[(size_t)c * N + r]);
                        if (!(e <= 1e-6)) ++bad;
                        if (e > maxerr || e != e) maxerr = e;
                    } else if (v == v) ++bad;       // must stay untouched (NaN pattern)
                }
        for (int i = 0; i < 5; ++i) fmlp_tag_sim_f32(feat, D, D, proto, C, S, rows.data(), missing.data(), sim, N, mode, ws, ws_bytes, 0);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) fmlp_tag_sim_f32(feat, D, D, proto, C, S, rows.data(), missing.data(), sim, N, mode, ws, ws_bytes, 0);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / iters;
        // the same call captured in a CUDA graph (does the programmatic dependent launch survive capture?)
        double us_graph = -1.0;
        {
            cudaStream_t cs; CK(cudaStreamCreate(&cs));
            cudaGraph_t g; cudaGraphExec_t ge;
            if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                fmlp_tag_sim_f32(feat, D, D, proto, C, S, rows.data(), missing.data(), sim, N, mode, ws, ws_bytes, cs);
                if (cudaStreamEndCapture(cs, &g) == cudaSuccess && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess) {
                    for (int i = 0; i < 5; ++i) cudaGraphLaunch(ge, cs);
                    CK(cudaStreamSynchronize(cs));
                    CK(cudaEventRecord(e0, cs));
                    for (int i = 0; i < iters; ++i) cudaGraphLaunch(ge, cs);
                    CK(cudaEventRecord(e1, cs));
                    CK(cudaStreamSynchronize(cs));
                    float msg; CK(cudaEventElapsedTime(&msg, e0, e1));
                    us_graph = msg * 1e3 / iters;
                }
            }
            cudaGetLastError();
        }
#ifdef FMLP_SIM_TRACE
        {   // per-CTA time stamps of ONE isolated launch (ns relative to the first CTA's entry)
            CK(cudaDeviceSynchronize());
            fmlp_tag_sim_f32(feat, D, D, proto, C, S, rows.data(), missing.data(), sim, N, mode, ws, ws_bytes, 0);
            CK(cudaDeviceSynchronize());
            std::vector<unsigned long long> tr(160 * 8 + 8);
            fmlp_sim_trace_read(tr.data(), (int)tr.size());
            unsigned long long t0 = tr[0];
            for (int b = 0; b < 148; ++b) t0 = tr[b * 8] < t0 ? tr[b * 8] : t0;     // first CTA entry
            double mn[4] = {1e18, 1e18, 1e18, 1e18}, mx[4] = {0, 0, 0, 0}, av[4] = {0, 0, 0, 0};
            const int ncta = 148;
            for (int b = 0; b < ncta; ++b)
                for (int i = 0; i < 4; ++i) {
                    const double v = (double)(long long)(tr[b * 8 + i] - t0);
                    mn[i] = v < mn[i] ? v : mn[i]; mx[i] = v > mx[i] ? v : mx[i]; av[i] += v / ncta;
                }
            printf("{\"trace\": \"%s\", \"table_cta0_done_ns\": %lld, \"table_other_done_ns\": %lld", mode ? "folded" : "pair",
                   (long long)(tr[160 * 8 + 1] - t0), (long long)(tr[160 * 8 + 2] - t0));
            const char* nm[4] = {"entry", "table_ready", "first_tile_done", "loop_done"};
            for (int i = 0; i < 4; ++i) printf(", \"%s_ns\": [%.0f, %.0f, %.0f]", nm[i], mn[i], av[i], mx[i]);
            printf("}\n");
        }
#endif
        const double bytes = 4.0 * N * D + 8.0 * C * D + 4.0 * (C - 1) * N;
        printf("{\"variant\": \"%s\", \"N\": %lld, \"D\": %d, \"C\": %d, \"S\": %d, \"mode\": \"%s\", \"us\": %.2f, \"us_graph\": %.2f, \"gbs\": %.1f, "
               "\"frac_6448\": %.3f, \"max_abs_err\": %.3g, \"bad\": %lld, \"checked\": %lld, \"stages_env\": \"%s\"}\n",
               EXP_VARIANT, (long long)N, D, C, S, mode ? "folded" : "pair", us, us_graph, bytes / us / 1e3, bytes / us / 1e3 / 6447.8,
               maxerr, bad, written, getenv("FMLP_SIM_STAGES") ? getenv("FMLP_SIM_STAGES") : "");
        fflush(stdout);
    }
    return 0;
}
