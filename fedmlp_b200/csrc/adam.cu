// adam.cu — fused Adam step over a flat parameter buffer (the optimizer half of the
// "update" entry point: utils/local_training.py:912-913,964-966 and :1149-1150,1189-1191 create
// torch.optim.Adam(lr, betas=(0.9, 0.999), weight_decay=5e-4) every round and step ~364 tensors per
// batch).  With the client model living in ONE flat fp32 buffer (flat.py) the whole step is one
// streaming launch: read p, g, m, v, write p, m, v = 28 bytes per parameter.
//
// Arithmetic = torch.optim.Adam (single-tensor path, amsgrad=False, maximize=False):
//     g   = g + wd * p
//     m   = m + (g - m) * (1 - b1)                      (exp_avg.lerp_)
//     v   = v * b2 + (1 - b2) * g * g                   (mul_ + addcmul_)
//     p   = p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Buffers that are not parameters (BatchNorm running statistics) sit between the parameters in the
// flat layout, so the kernel walks a chunk table of the parameter runs only.
#include "common.cuh"

namespace fmlp {

struct AdamArgs {
    float* p;
    const float* g;
    float* m;
    float* v;
    const int64_t* chunk_start;  // [n_chunks] element offset of the chunk in the flat buffers
    const int32_t* chunk_len;    // [n_chunks] <= FMLP_FEDAVG_CHUNK
    int64_t n_chunks;
    float lr_over_bc1, inv_sqrt_bc2, eps, wd, b1, b2;
    int zero_grad;               // also clear g (optimizer.zero_grad fused into the step)
    float* g_mut;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a) {
    g = fmaf(a.wd, p, g);
    m = fmaf(g - m, 1.f - a.b1, m);
    v = fmaf(v, a.b2, (1.f - a.b2) * g * g);
    const float denom = sqrtf(v) * a.inv_sqrt_bc2 + a.eps;
    p = p - a.lr_over_bc1 * (m / denom);
}

__global__ void __launch_bounds__(256, 4) adam_step_kernel(const __grid_constant__ AdamArgs a) {
    for (int64_t c = blockIdx.x; c < a.n_chunks; c += gridDim.x) {
        const int64_t s = a.chunk_start[c];
        const int len = a.chunk_len[c];
        const bool vec = ((s & 3) == 0);
        const int nvec = vec ? (len >> 2) : 0;
        for (int i = threadIdx.x; i < nvec; i += 256) {
            const int64_t e = s + 4 * i;
            float4 p = *reinterpret_cast<const float4*>(a.p + e);
            const float4 g = *reinterpret_cast<const float4*>(a.g + e);
            float4 m = *reinterpret_cast<const float4*>(a.m + e);
            float4 v = *reinterpret_cast<const float4*>(a.v + e);
            adam_one(p.x, g.x, m.x, v.x, a); adam_one(p.y, g.y, m.y, v.y, a);
            adam_one(p.z, g.z, m.z, v.z, a); adam_one(p.w, g.w, m.w, v.w, a);
            *reinterpret_cast<float4*>(a.p + e) = p;
            *reinterpret_cast<float4*>(a.m + e) = m;
            *reinterpret_cast<float4*>(a.v + e) = v;
            if (a.zero_grad) *reinterpret_cast<float4*>(a.g_mut + e) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int i = 4 * nvec + threadIdx.x; i < len; i += 256) {
            const int64_t e = s + i;
            float p = a.p[e], m = a.m[e], v = a.v[e];
            adam_one(p, a.g[e], m, v, a);
            a.p[e] = p; a.m[e] = m; a.v[e] = v;
            if (a.zero_grad) a.g_mut[e] = 0.f;
        }
    }
}

}  // namespace fmlp

using namespace fmlp;

extern "C" int fmlp_adam_step_f32(float* p, float* g, float* m, float* v, const int64_t* chunk_start_dev,
                                  const int32_t* chunk_len_dev, int64_t n_chunks, float lr, float beta1,
                                  float beta2, float eps, float weight_decay, int64_t step, int zero_grad,
                                  fmlp_stream_t stream) {
    if (!p || !g || !m || !v || !chunk_start_dev || !chunk_len_dev || n_chunks < 0 || step < 1) return FMLP_ERR_BAD_ARG;
    if (!aligned16(p) || !aligned16(g) || !aligned16(m) || !aligned16(v)) return FMLP_ERR_UNSUPPORTED;
    if (n_chunks == 0) return FMLP_OK;
    AdamArgs a;
    a.p = p; a.g = g; a.g_mut = g; a.m = m; a.v = v; a.chunk_start = chunk_start_dev; a.chunk_len = chunk_len_dev;
    a.n_chunks = n_chunks;
    // bias corrections in double on the host, like torch's Python scalars
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    a.lr_over_bc1 = (float)((double)lr / bc1);
    a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    a.eps = eps; a.wd = weight_decay; a.b1 = beta1; a.b2 = beta2; a.zero_grad = zero_grad;
    const int sms = sm_count();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    int64_t blocks = n_chunks < (int64_t)sms * 8 ? n_chunks : (int64_t)sms * 8;
    adam_step_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    return launch_status();
}
