// SSIM of RegressionMetrics (p2igan_bench/metrics/metric.py:36,55-56,69: torchmetrics.image
// StructuralSimilarityIndexMeasure(data_range), defaults gaussian_kernel=True, sigma=1.5, kernel_size=11, k1=0.01,
// k2=0.03, reduction='elementwise_mean').  PARITY UNPINNED: torchmetrics 1.0.3 is a third-party dependency absent from
// /root/reference and from this image, and no reference test holds an SSIM value; the algorithm below restates its
// published implementation (functional/image/ssim.py::_ssim_update): reflect-pad by 5, 11x11 gaussian filtering of
// p, t, p*p, t*t, p*t, the SSIM map, crop of the 5-pixel border, mean per image; state = sum of per-image means and
// image count.  The crop removes every output that saw padding, so only the (H-10) x (W-10) interior is evaluated.
// One pass: a block stages a 42 x 42 patch of both inputs (transformed to rain rate when asked), filters separably
// through shared memory (5 maps) and reduces its 32 x 32 SSIM values.  8 B/pixel read (x 1.7 halo, mostly L1/L2 hits).
#include "common.h"
#include "ptx.cuh"

namespace p2i {

struct SsimGauss { float g[11]; };

constexpr int SS_T = 32;            // output tile edge: a 42 x 42 staged patch per 32 x 32 outputs (1.7x halo; 16 x 16 tiles read 2.6x)
constexpr int SS_S = SS_T + 10;

__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ pred, const float* __restrict__ target, int N, int H, int W,
                                                   int apply_transform, float c1, float c2, SsimGauss gw, float inv_npix,
                                                   double* __restrict__ state) {
    __shared__ float sp[SS_S][SS_S + 1], st[SS_S][SS_S + 1];
    __shared__ float hz[5][SS_S][SS_T + 1];
    __shared__ float red[8];
    const int n = blockIdx.z;
    const int oy0 = blockIdx.y * SS_T, ox0 = blockIdx.x * SS_T;   // interior coordinates: output (oy, ox) <- rows oy..oy+10
    const float* P = pred + static_cast<size_t>(n) * H * W;
    const float* T = target + static_cast<size_t>(n) * H * W;
    for (int e = threadIdx.x; e < SS_S * SS_S; e += 256) {
        const int r = e / SS_S, c = e - r * SS_S;
        const int y = oy0 + r, x = ox0 + c;
        float a = 0.f, b = 0.f;
        if (y < H && x < W) {
            a = P[static_cast<size_t>(y) * W + x];
            b = T[static_cast<size_t>(y) * W + x];
            if (apply_transform) {                                  // metric.py:16-20: 0.036 * 10^(x/16) = 0.036 * 2^(x * log2(10)/16)
                a = exp2f(a * 0.20762050593046f) * 0.036f;
                b = exp2f(b * 0.20762050593046f) * 0.036f;
            }
        }
        sp[r][c] = a;
        st[r][c] = b;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < SS_S * SS_T; e += 256) {          // horizontal pass: 42 rows x 32 output columns
        const int r = e / SS_T, c = e - r * SS_T;
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float a = sp[r][c + k], b = st[r][c + k], g = gw.g[k];
            m0 = fmaf(g, a, m0); m1 = fmaf(g, b, m1); m2 = fmaf(g, a * a, m2); m3 = fmaf(g, b * b, m3); m4 = fmaf(g, a * b, m4);
        }
        hz[0][r][c] = m0; hz[1][r][c] = m1; hz[2][r][c] = m2; hz[3][r][c] = m3; hz[4][r][c] = m4;
    }
    __syncthreads();
    float v = 0.f;
    for (int e = threadIdx.x; e < SS_T * SS_T; e += 256) {          // vertical pass + SSIM map: 4 outputs per thread
        const int ty = e / SS_T, tx = e - ty * SS_T;
        if (oy0 + ty >= H - 10 || ox0 + tx >= W - 10) continue;
        float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float g = gw.g[k];
#pragma unroll
            for (int q = 0; q < 5; ++q) m[q] = fmaf(g, hz[q][ty + k][tx], m[q]);
        }
        const float mu_pp = m[0] * m[0], mu_tt = m[1] * m[1], mu_pt = m[0] * m[1];
        const float s_p = m[2] - mu_pp, s_t = m[3] - mu_tt, s_pt = m[4] - mu_pt;
        const float upper = 2.f * s_pt + c2, lower = s_p + s_t + c2;
        v += ((2.f * mu_pt + c1) * upper) / ((mu_pp + mu_tt + c1) * lower);
    }
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        atomicAdd(&state[0], static_cast<double>(s) * static_cast<double>(inv_npix));      // fp64: thousands of partial sums
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) atomicAdd(&state[1], static_cast<double>(N));
    }
}

}  // namespace p2i

extern "C" int p2i_ssim_update(const float* pred, const float* target, int N, int H, int W, int apply_transform, float data_range,
                               double* state, void* stream) {
    P2I_CHECK_ARG(pred && target && state, "ssim_update: null pointer");
    P2I_CHECK_ARG(N > 0 && H > 10 && W > 10, "ssim_update: images must be larger than the 11x11 window (got %dx%d)", H, W);
    p2i::SsimGauss gw;
    double g[11], sum = 0.0;
    for (int i = 0; i < 11; ++i) {
        const double d = (i - 5) / 1.5;
        g[i] = exp(-d * d / 2.0);
        sum += g[i];
    }
    for (int i = 0; i < 11; ++i) gw.g[i] = static_cast<float>(g[i] / sum);
    const float c1 = (0.01f * data_range) * (0.01f * data_range), c2 = (0.03f * data_range) * (0.03f * data_range);
    dim3 grid(p2i::cdiv(W - 10, p2i::SS_T), p2i::cdiv(H - 10, p2i::SS_T), N);
    P2I_CHECK_ARG(N <= 65535, "ssim_update: at most 65535 frames per call");
    p2i::ssim_kernel<<<grid, 256, 0, p2i::as_stream(stream)>>>(pred, target, N, H, W, apply_transform, c1, c2, gw,
                                                               1.f / (static_cast<float>(H - 10) * static_cast<float>(W - 10)), state);
    P2I_CHECK_LAUNCH("ssim_kernel");
    return P2I_OK;
}
