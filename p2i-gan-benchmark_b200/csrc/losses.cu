// Loss reductions of the training step (p2igan_bench/modules/losses.py), HBM-bound warp-shuffle kernels.
//   ReconstructionLoss (losses.py:38-48): weighted L1 (:56-65) + KL between the temperature-softmax of the
//   temporal differences of prediction and target (:68-85).
//   AdversarialLoss (losses.py:192-226): hinge / nsgan (BCE on raw outputs) / lsgan (MSE).
#include "common.h"
#include "ptx.cuh"

namespace p2i {

__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
    if (warp == 0) v = warp_sum(v);
    return v;   // valid in warp 0
}
__device__ __forceinline__ float block_max(float v, float* sh) {
    v = warp_max(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : -INFINITY;
    if (warp == 0) v = warp_max(v);
    __syncthreads();
    if (threadIdx.x == 0) sh[0] = v;
    __syncthreads();
    return sh[0];   // valid in all threads
}

__device__ __forceinline__ float l1_weight(float y) { return 0.5f * expf(5.14f * fminf(y, 0.7f)) + 0.12f; }

// sums[0] += sum w(y) |p - y|
__global__ void __launch_bounds__(256) wl1_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                      long long n, float* __restrict__ sums) {
    __shared__ float sh[32];
    float acc = 0.f;
    const long long n4 = n >> 2;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(pred) + i), y = __ldg(reinterpret_cast<const float4*>(target) + i);
        acc += l1_weight(y.x) * fabsf(p.x - y.x) + l1_weight(y.y) * fabsf(p.y - y.y) + l1_weight(y.z) * fabsf(p.z - y.z) +
               l1_weight(y.w) * fabsf(p.w - y.w);
    }
    if (blockIdx.x == 0)
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) acc += l1_weight(target[i]) * fabsf(pred[i] - target[i]);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) atomicAdd(&sums[0], acc);
}

// One block per (b, t) row of temporal differences: log-sum-exp of dp/temp and dq/temp over the HW pixels, then
// KL_row = sum_i q_i (dq_i - dp_i) - lse_q + lse_p.   sums[1] += KL_row ; lse[row] = {lse_p, lse_q}.
__global__ void __launch_bounds__(1024) tkl_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ target, int T,
                                                       int HW, float inv_temp, float* __restrict__ sums, float* __restrict__ lse) {
    __shared__ float sh[32];
    const int row = blockIdx.x;
    const int b = row / (T - 1), t = row - b * (T - 1);
    const float* p0 = pred + (static_cast<size_t>(b) * T + t) * HW;
    const float* q0 = target + (static_cast<size_t>(b) * T + t) * HW;
    const float* p1 = p0 + HW;
    const float* q1 = q0 + HW;
    float mp = -INFINITY, mq = -INFINITY;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        mp = fmaxf(mp, (p1[i] - p0[i]) * inv_temp);
        mq = fmaxf(mq, (q1[i] - q0[i]) * inv_temp);
    }
    mp = block_max(mp, sh);
    mq = block_max(mq, sh);
    float sp = 0.f, sq = 0.f, cross = 0.f;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        const float dp = (p1[i] - p0[i]) * inv_temp, dq = (q1[i] - q0[i]) * inv_temp;
        const float eq = expf(dq - mq);
        sp += expf(dp - mp);
        sq += eq;
        cross += eq * (dq - dp);
    }
    sp = block_sum(sp, sh);
    sq = block_sum(sq, sh);
    cross = block_sum(cross, sh);
    if (threadIdx.x == 0) {
        const float lp = mp + logf(sp), lq = mq + logf(sq);
        lse[2 * row] = lp;
        lse[2 * row + 1] = lq;
        atomicAdd(&sums[1], cross / sq - lq + lp);
    }
}

// dpred = gscale * ( w(y) sign(p-y)/N  +  k1/(B*temp) * [ (p-q)_{row t-1} - (p-q)_{row t} ] )
__global__ void __launch_bounds__(256) rec_loss_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                           const float* __restrict__ lse, const float* __restrict__ gscale,
                                                           float k1, float inv_temp, float* __restrict__ dpred, int B, int T, int HW) {
    const long long n = static_cast<long long>(B) * T * HW;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pix = static_cast<int>(i % HW);
    const int t = static_cast<int>((i / HW) % T), b = static_cast<int>(i / (static_cast<long long>(HW) * T));
    const float gs = gscale ? __ldg(gscale) : 1.f;
    const float p = pred[i], y = target[i];
    const float d = p - y;
    float g = l1_weight(y) * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) / static_cast<float>(n);
    if (k1 != 0.f) {
        const float c = k1 * inv_temp / static_cast<float>(B);
        float acc = 0.f;
        if (t > 0) {          // row t-1 uses frames (t-1, t): d/dpred[t] = +(p - q)
            const int row = b * (T - 1) + t - 1;
            const float dp = (p - pred[i - HW]) * inv_temp, dq = (y - target[i - HW]) * inv_temp;
            acc += expf(dp - lse[2 * row]) - expf(dq - lse[2 * row + 1]);
        }
        if (t < T - 1) {      // row t uses frames (t, t+1): d/dpred[t] = -(p - q)
            const int row = b * (T - 1) + t;
            const float dp = (pred[i + HW] - p) * inv_temp, dq = (target[i + HW] - y) * inv_temp;
            acc -= expf(dp - lse[2 * row]) - expf(dq - lse[2 * row + 1]);
        }
        g += c * acc;
    }
    dpred[i] = gs * g;
    (void)pix;
}

// mode: 0 hinge D real  relu(1-x) | 1 hinge D fake  relu(1+x) | 2 hinge G  -x | 3 BCE(x, label) | 4 (x-label)^2
__device__ __forceinline__ float gan_term(float x, int mode, float label) {
    switch (mode) {
        case 0: return fmaxf(1.f - x, 0.f);
        case 1: return fmaxf(1.f + x, 0.f);
        case 2: return -x;
        case 3: return -(label * fmaxf(logf(x), -100.f) + (1.f - label) * fmaxf(logf(1.f - x), -100.f));
        default: return (x - label) * (x - label);
    }
}
__device__ __forceinline__ float gan_dterm(float x, int mode, float label) {
    switch (mode) {
        case 0: return (1.f - x > 0.f) ? -1.f : 0.f;
        case 1: return (1.f + x > 0.f) ? 1.f : 0.f;
        case 2: return -1.f;
        case 3: return -(label / fmaxf(x, 1e-12f) - (1.f - label) / fmaxf(1.f - x, 1e-12f));
        default: return 2.f * (x - label);
    }
}
__global__ void __launch_bounds__(256) gan_loss_fwd_kernel(const float* __restrict__ x, long long n, int mode, float label,
                                                           float* __restrict__ out) {
    __shared__ float sh[32];
    float acc = 0.f;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        acc += gan_term(x[i], mode, label);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) atomicAdd(out, acc / static_cast<float>(n));
}
__global__ void __launch_bounds__(256) gan_loss_bwd_kernel(const float* __restrict__ x, long long n, int mode, float label,
                                                           const float* __restrict__ gscale, float scale, float* __restrict__ dx) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gs = (gscale ? __ldg(gscale) : 1.f) * scale / static_cast<float>(n);
    dx[i] = gs * gan_dterm(x[i], mode, label);
}

}  // namespace p2i

using namespace p2i;

extern "C" int p2i_rec_loss_fwd(const float* pred, const float* target, int B, int T, int HW, float temperature,
                                float* sums, float* lse, void* stream) {
    P2I_CHECK_ARG(pred && target && sums && lse, "rec_loss_fwd: null pointer");
    P2I_CHECK_ARG(B > 0 && T > 1 && HW > 0 && temperature > 0.f, "rec_loss_fwd: bad shape");
    const long long n = static_cast<long long>(B) * T * HW;
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    wl1_fwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(pred, target, n, sums);
    P2I_CHECK_LAUNCH("wl1_fwd_kernel");
    tkl_fwd_kernel<<<B * (T - 1), 1024, 0, as_stream(stream)>>>(pred, target, T, HW, 1.f / temperature, sums, lse);
    P2I_CHECK_LAUNCH("tkl_fwd_kernel");
    return P2I_OK;
}

extern "C" int p2i_rec_loss_bwd(const float* pred, const float* target, const float* lse, const float* gscale, float k1_alpha,
                                float temperature, float* dpred, int B, int T, int HW, void* stream) {
    P2I_CHECK_ARG(pred && target && lse && dpred, "rec_loss_bwd: null pointer");
    const long long n = static_cast<long long>(B) * T * HW;
    rec_loss_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream)>>>(
        pred, target, lse, gscale, k1_alpha, 1.f / temperature, dpred, B, T, HW);
    P2I_CHECK_LAUNCH("rec_loss_bwd_kernel");
    return P2I_OK;
}

extern "C" int p2i_gan_loss_fwd(const float* logits, long long n, int mode, float label, float* out, void* stream) {
    P2I_CHECK_ARG(logits && out && n > 0 && mode >= 0 && mode <= 4, "gan_loss_fwd: bad arguments");
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    gan_loss_fwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(logits, n, mode, label, out);
    P2I_CHECK_LAUNCH("gan_loss_fwd_kernel");
    return P2I_OK;
}

extern "C" int p2i_gan_loss_bwd(const float* logits, long long n, int mode, float label, const float* gscale, float scale,
                                float* dlogits, void* stream) {
    P2I_CHECK_ARG(logits && dlogits && n > 0 && mode >= 0 && mode <= 4, "gan_loss_bwd: bad arguments");
    gan_loss_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream)>>>(logits, n, mode, label, gscale,
                                                                                                  scale, dlogits);
    P2I_CHECK_LAUNCH("gan_loss_bwd_kernel");
    return P2I_OK;
}
