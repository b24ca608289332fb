// Fused multi-tensor Adam (torch.optim.Adam semantics, no weight decay / amsgrad; reference optimiser:
// scripts/train.py:125-136, lr 1e-4, betas (0, 0.99)).  One launch updates every parameter of a model:
// the work list is a device table of (param, grad, exp_avg, exp_avg_sq, n) cut into fixed-size chunks.
#include "common.h"
#include "ptx.cuh"

namespace p2i {

constexpr int ADAM_CHUNK = 8192;

// step[0] += 1; step[1] = lr / (1 - beta1^t); step[2] = 1 / sqrt(1 - beta2^t)
__global__ void adam_tick_kernel(float* step, float lr, float beta1, float beta2) {
    const float t = step[0] + 1.f;
    step[0] = t;
    const double st = static_cast<double>(t);
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), st), bc2 = 1.0 - pow(static_cast<double>(beta2), st);
    step[1] = static_cast<float>(static_cast<double>(lr) / bc1);
    step[2] = static_cast<float>(1.0 / sqrt(bc2));
}

__global__ void __launch_bounds__(256) adam_kernel(const P2iAdamTensor* __restrict__ tensors, const int2* __restrict__ chunks,
                                                   const float* __restrict__ step_dev, float lr, float beta1, float beta2, float eps,
                                                   float grad_scale) {
    const float lr_over_bc1 = __ldg(step_dev + 1), inv_sqrt_bc2 = __ldg(step_dev + 2);
    const int2 ch = chunks[blockIdx.x];               // (tensor index, chunk index)
    const P2iAdamTensor t = tensors[ch.x];
    const long long start = static_cast<long long>(ch.y) * ADAM_CHUNK;
    const long long end = (start + ADAM_CHUNK < t.n) ? start + ADAM_CHUNK : t.n;
    const bool vec = ((reinterpret_cast<uintptr_t>(t.param) | reinterpret_cast<uintptr_t>(t.grad) |
                       reinterpret_cast<uintptr_t>(t.exp_avg) | reinterpret_cast<uintptr_t>(t.exp_avg_sq)) & 15) == 0;
    const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
    long long i0 = start;
    if (vec) {
        const long long n4 = (end - start) >> 2;
        float4* P = reinterpret_cast<float4*>(t.param + start);
        const float4* Gd = reinterpret_cast<const float4*>(t.grad + start);
        float4* M = reinterpret_cast<float4*>(t.exp_avg + start);
        float4* V = reinterpret_cast<float4*>(t.exp_avg_sq + start);
        for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
            float4 p = P[i], g = Gd[i], m = M[i], v = V[i];
            float* pp = &p.x; float* gp = &g.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gg = gp[k] * grad_scale;
                mp[k] = mp[k] + (gg - mp[k]) * omb1;
                vp[k] = vp[k] * beta2 + gg * gg * omb2;
                pp[k] -= lr_over_bc1 * (mp[k] / (sqrtf(vp[k]) * inv_sqrt_bc2 + eps));
            }
            P[i] = p; M[i] = m; V[i] = v;
        }
        i0 = start + (n4 << 2);
    }
    for (long long i = i0 + threadIdx.x; i < end; i += blockDim.x) {
        const float g = t.grad[i] * grad_scale;
        float m = t.exp_avg[i], v = t.exp_avg_sq[i];
        m = m + (g - m) * omb1;                       // lerp_
        v = v * beta2 + g * g * omb2;                 // mul_().addcmul_()
        const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
        t.param[i] -= lr_over_bc1 * (m / denom);
        t.exp_avg[i] = m;
        t.exp_avg_sq[i] = v;
    }
}

}  // namespace p2i

extern "C" int p2i_adam_tick(float* step_dev, float lr, float beta1, float beta2, void* stream) {
    P2I_CHECK_ARG(step_dev, "adam_tick: null pointer");
    p2i::adam_tick_kernel<<<1, 1, 0, p2i::as_stream(stream)>>>(step_dev, lr, beta1, beta2);
    P2I_CHECK_LAUNCH("adam_tick_kernel");
    return P2I_OK;
}

extern "C" int p2i_adam_apply(const P2iAdamTensor* tensors_dev, const int* chunks_dev, int n_chunks, const float* step_dev, float lr,
                              float beta1, float beta2, float eps, float grad_scale, void* stream) {
    P2I_CHECK_ARG(tensors_dev && chunks_dev && n_chunks > 0 && step_dev, "adam_apply: bad arguments");
    p2i::adam_kernel<<<n_chunks, 256, 0, p2i::as_stream(stream)>>>(tensors_dev, reinterpret_cast<const int2*>(chunks_dev), step_dev, lr,
                                                                   beta1, beta2, eps, grad_scale);
    P2I_CHECK_LAUNCH("adam_kernel");
    return P2I_OK;
}

extern "C" int p2i_adam_step(const P2iAdamTensor* tensors_dev, const int* chunks_dev, int n_chunks, float* step_dev, float lr,
                             float beta1, float beta2, float eps, float grad_scale, void* stream) {
    P2I_CHECK_ARG(tensors_dev && chunks_dev && n_chunks > 0 && step_dev, "adam_step: bad arguments");
    int rc = p2i_adam_tick(step_dev, lr, beta1, beta2, stream);
    if (rc) return rc;
    return p2i_adam_apply(tensors_dev, chunks_dev, n_chunks, step_dev, lr, beta1, beta2, eps, grad_scale, stream);
}

extern "C" int p2i_adam_chunk_elems(void) { return p2i::ADAM_CHUNK; }
