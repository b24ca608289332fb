// Fused multi-tensor Adam (torch.optim.Adam semantics, no weight decay / amsgrad; reference optimiser:
// scripts/train.py:125-136, lr 1e-4, betas (0, 0.99)).  One launch updates every parameter of a model:
// the work list is a device table of (param, grad, exp_avg, exp_avg_sq, n) cut into fixed-size chunks.
#include "common.h"
#include "ptx.cuh"

namespace p2i {

constexpr int ADAM_CHUNK = 65536;

__global__ void __launch_bounds__(256) adam_kernel(const P2iAdamTensor* __restrict__ tensors, const int2* __restrict__ chunks,
                                                   float lr_over_bc1, float inv_sqrt_bc2, float beta1, float beta2, float eps,
                                                   float grad_scale) {
    const int2 ch = chunks[blockIdx.x];               // (tensor index, chunk index)
    const P2iAdamTensor t = tensors[ch.x];
    const long long start = static_cast<long long>(ch.y) * ADAM_CHUNK;
    const long long end = (start + ADAM_CHUNK < t.n) ? start + ADAM_CHUNK : t.n;
    for (long long i = start + threadIdx.x; i < end; i += blockDim.x) {
        const float g = t.grad[i] * grad_scale;
        float m = t.exp_avg[i], v = t.exp_avg_sq[i];
        m = m + (g - m) * (1.f - beta1);              // lerp_
        v = v * beta2 + g * g * (1.f - beta2);        // mul_().addcmul_()
        const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
        t.param[i] -= lr_over_bc1 * (m / denom);
        t.exp_avg[i] = m;
        t.exp_avg_sq[i] = v;
    }
}

}  // namespace p2i

extern "C" int p2i_adam_step(const P2iAdamTensor* tensors_dev, const int* chunks_dev, int n_chunks, float lr, float beta1,
                             float beta2, float eps, int step, float grad_scale, void* stream) {
    P2I_CHECK_ARG(tensors_dev && chunks_dev && n_chunks > 0 && step >= 1, "adam_step: bad arguments");
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
    p2i::adam_kernel<<<n_chunks, 256, 0, p2i::as_stream(stream)>>>(tensors_dev, reinterpret_cast<const int2*>(chunks_dev),
                                                                   static_cast<float>(lr / bc1), static_cast<float>(1.0 / sqrt(bc2)),
                                                                   beta1, beta2, eps, grad_scale);
    P2I_CHECK_LAUNCH("adam_kernel");
    return P2I_OK;
}

extern "C" int p2i_adam_chunk_elems(void) { return p2i::ADAM_CHUNK; }
