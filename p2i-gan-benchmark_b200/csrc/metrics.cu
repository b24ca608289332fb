// Evaluation metric reductions (p2igan_bench/metrics/metric.py): one pass over prediction and target computes
//   RegressionMetrics   (:28-74)   sum |d|, sum d^2 on rain rates R = 0.036 * 10^(x/16)   (:16-20)
//   CategoricalMetrics  (:77-134)  hits / misses / false alarms / correct negatives per threshold
//   FractionalSkillScore (:137-183) per (threshold, scale): sum (Pbar - Tbar)^2 and sum (Pbar^2 + Tbar^2) where
//                                   Xbar = avg_pool2d(mask, k, stride 1, pad k/2) with the zero padding counted
//                                   (even k => (H+1) x (W+1) outputs)
// instead of the reference's 1 + 4 + 32 full-tensor passes.  HBM-bound: 8 bytes read per pixel.
#include "common.h"
#include "ptx.cuh"

namespace p2i {

constexpr int MT = 32;          // output tile edge
constexpr int MH = 4;           // halo on the low side (k = 8 reaches o-4 .. o+3)
constexpr int MS = MT + 8;      // staged edge

struct MetricParams {
    int N, H, W;                // frames, spatial size
    int n_thr, n_scale;
    float thr[4];
    int scale[4];               // subset of {1,2,4,8}
    int apply_transform;        // regression sums on transformed values?
};

// R = 0.036 * 10^(x/16) (metric.py:16-20) as one exp2: 10^(x/16) = 2^(x * log2(10)/16)
__device__ __forceinline__ float rain(float x) { return exp2f(x * 0.20762050593046f) * 0.036f; }

// acc layout (double): [0] abs_sum [1] sq_sum | [2 + 4*t + {0..3}] hits, misses, false, correct |
//                      [18 + 2*(t*4+s) + {0,1}] fss numerator sum, denominator sum
__global__ void __launch_bounds__(256) metrics_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                      double* __restrict__ acc, const MetricParams p) {
    __shared__ unsigned long long sm[MS * MS];     // per pixel: byte lane 2*t = pred mask, 2*t+1 = target mask (threshold t)
    __shared__ unsigned long long hs[3][MS * MT];  // horizontal box sums of widths 2 / 4 / 8 (staged row x tile column)
    __shared__ float red[8][50];
    const int n = blockIdx.z;
    const int oy0 = blockIdx.y * MT, ox0 = blockIdx.x * MT;
    const float* P = pred + static_cast<size_t>(n) * p.H * p.W;
    const float* Tg = target + static_cast<size_t>(n) * p.H * p.W;
    float a_abs = 0.f, a_sq = 0.f;
    float cont[4][4];                              // [threshold][hits, misses, false alarms, correct negatives]: static indices only
    float fs[16][2];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int k = 0; k < 4; ++k) cont[t][k] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) fs[i][0] = fs[i][1] = 0.f;

    for (int e = threadIdx.x; e < MS * MS; e += blockDim.x) {
        const int sy = e / MS, sx = e - sy * MS;
        const int y = oy0 - MH + sy, x = ox0 - MH + sx;
        unsigned long long bits = 0ull;
        if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
            const float pv = P[static_cast<size_t>(y) * p.W + x], tv = Tg[static_cast<size_t>(y) * p.W + x];
            const float pr = rain(pv), tr = rain(tv);
            const bool own = (sy >= MH && sy < MH + MT && sx >= MH && sx < MH + MT);   // pixel belongs to this tile
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (t < p.n_thr) {
                    const bool a = pr >= p.thr[t], o = tr >= p.thr[t];
                    bits |= (a ? 1ull : 0ull) << (16 * t);
                    bits |= (o ? 1ull : 0ull) << (16 * t + 8);
                    if (own) {
                        cont[t][0] += (a && o) ? 1.f : 0.f;
                        cont[t][1] += (!a && o) ? 1.f : 0.f;
                        cont[t][2] += (a && !o) ? 1.f : 0.f;
                        cont[t][3] += (!a && !o) ? 1.f : 0.f;
                    }
                }
            }
            if (own) {
                const float d = p.apply_transform ? (pr - tr) : (pv - tv);
                a_abs += fabsf(d);
                a_sq += d * d;
            }
        }
        sm[e] = bits;
    }
    __syncthreads();
    // separable box sums of the packed masks (byte lanes never overflow: at most 64 pixels per box): horizontal sums of
    // widths 2 / 4 / 8 per staged row (8 shared-memory loads), then vertical sums (14 loads) -- 64 loads per output before
    for (int e = threadIdx.x; e < MS * MT; e += blockDim.x) {
        const int sy = e / MT, tx = e - sy * MT;
        const unsigned long long* r = sm + sy * MS + tx + MH;
        const unsigned long long h2 = r[-1] + r[0];
        const unsigned long long h4 = h2 + r[-2] + r[1];
        const unsigned long long h8 = h4 + r[-4] + r[-3] + r[2] + r[3];
        hs[0][e] = h2; hs[1][e] = h4; hs[2][e] = h8;
    }
    __syncthreads();
    // every thread owns 4 output positions of the tile
    for (int e = threadIdx.x; e < MT * MT; e += blockDim.x) {
        const int ty = e / MT, tx = e - ty * MT;
        const int oy = oy0 + ty, ox = ox0 + tx;
        if (oy > p.H || ox > p.W) continue;
        const int cy = ty + MH, cx = tx + MH;                 // staged coordinate of input pixel (oy, ox)
        const unsigned long long* c = &hs[0][cy * MT + tx];
        const unsigned long long c2 = c[-MT] + c[0];
        c = &hs[1][cy * MT + tx];
        const unsigned long long c4 = c[-2 * MT] + c[-MT] + c[0] + c[MT];
        c = &hs[2][cy * MT + tx];
        const unsigned long long c8 = c[-4 * MT] + c[-3 * MT] + c[-2 * MT] + c[-MT] + c[0] + c[MT] + c[2 * MT] + c[3 * MT];
        const unsigned long long c1 = sm[cy * MS + cx];
        const bool in1 = (oy < p.H && ox < p.W);              // k = 1 has an H x W output domain
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (s >= p.n_scale) continue;
            const int k = p.scale[s];
            if (k == 1 && !in1) continue;
            const unsigned long long cc = (k == 1) ? c1 : (k == 2 ? c2 : (k == 4 ? c4 : c8));
            const float inv = 1.f / static_cast<float>(k * k);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (t >= p.n_thr) continue;
                const float a = static_cast<float>((cc >> (16 * t)) & 0xffull) * inv;
                const float b = static_cast<float>((cc >> (16 * t + 8)) & 0xffull) * inv;
                fs[t * 4 + s][0] += (a - b) * (a - b);
                fs[t * 4 + s][1] += a * a + b * b;
            }
        }
    }
    // block reduction of the 2 + 16 + 32 partial sums -> double atomics
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float vals[50];
    vals[0] = a_abs; vals[1] = a_sq;
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int k = 0; k < 4; ++k) vals[2 + 4 * t + k] = cont[t][k];
#pragma unroll
    for (int i = 0; i < 16; ++i) { vals[18 + 2 * i] = fs[i][0]; vals[19 + 2 * i] = fs[i][1]; }
#pragma unroll
    for (int j = 0; j < 50; ++j) {
        const float r = warp_sum(vals[j]);
        if (lane == 0) red[warp][j] = r;
    }
    __syncthreads();
    if (threadIdx.x < 50) {
        float r = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) r += red[w8][threadIdx.x];
        if (r != 0.f) atomicAdd(&acc[threadIdx.x], static_cast<double>(r));
    }
}

// Streaming-state update with the reference's semantics: regression sums and contingency counts are added as float32,
// every (threshold, scale) gets  score_sum += 1 - mean_num / (mean_den + 1e-10)  and  counts += 1  PER UPDATE CALL.
// state (float32): [0] abs_sum [1] squared_sum [2] n_obs | [3 + 4*t + k] hits/misses/false/correct | [19 + t*4+s] score_sum
//                  | [35 + t*4+s] counts
__global__ void metrics_finalize_kernel(double* __restrict__ acc, float* __restrict__ state, MetricParams p) {
    const int i = threadIdx.x;
    const double n_pix = static_cast<double>(p.N) * p.H * p.W;
    if (i == 0) {
        state[0] += static_cast<float>(acc[0]);
        state[1] += static_cast<float>(acc[1]);
        state[2] += static_cast<float>(n_pix);
    }
    if (i < 16 && (i >> 2) < p.n_thr) state[3 + i] += static_cast<float>(acc[2 + i]);
    if (i < 16) {
        const int t = i >> 2, s = i & 3;
        if (t < p.n_thr && s < p.n_scale) {
            const int k = p.scale[s];
            const double cnt = (k & 1) ? n_pix : static_cast<double>(p.N) * (p.H + 1) * (p.W + 1);
            const float num = static_cast<float>(acc[18 + 2 * i] / cnt), den = static_cast<float>(acc[19 + 2 * i] / cnt);
            state[19 + i] += 1.0f - num / (den + 1e-10f);
            state[35 + i] += 1.0f;
        }
    }
    __syncthreads();
    if (i < 50) acc[i] = 0.0;       // scratch is left clean for the next update
}

// FSS for one (threshold, box size k <= 32): separable box sums of the two exceedance masks through shared memory.
constexpr int FT = 32;                 // output tile edge
constexpr int FSM = FT + 31;           // staged edge for k = 32
__global__ void __launch_bounds__(256) fss_generic_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                          double* __restrict__ acc, int N, int H, int W, float thr, int k) {
    __shared__ unsigned char sp[FSM * FSM], st[FSM * FSM];
    __shared__ unsigned short hp[FSM * FT], ht[FSM * FT];
    __shared__ float red[8][2];
    const int n = blockIdx.z;
    const int pad = k >> 1;
    const int Ho = H + 2 * pad - k + 1, Wo = W + 2 * pad - k + 1;
    const int oy0 = blockIdx.y * FT, ox0 = blockIdx.x * FT;
    const int S = FT + k - 1;
    const float* P = pred + static_cast<size_t>(n) * H * W;
    const float* Tg = target + static_cast<size_t>(n) * H * W;
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
        const int sy = e / S, sx = e - sy * S;
        const int y = oy0 - pad + sy, x = ox0 - pad + sx;
        unsigned char a = 0, b = 0;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            a = rain(P[static_cast<size_t>(y) * W + x]) >= thr;
            b = rain(Tg[static_cast<size_t>(y) * W + x]) >= thr;
        }
        sp[sy * FSM + sx] = a;
        st[sy * FSM + sx] = b;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < S * FT; e += blockDim.x) {      // horizontal box sums
        const int sy = e / FT, tx = e - sy * FT;
        int a = 0, b = 0;
        for (int d = 0; d < k; ++d) { a += sp[sy * FSM + tx + d]; b += st[sy * FSM + tx + d]; }
        hp[sy * FT + tx] = static_cast<unsigned short>(a);
        ht[sy * FT + tx] = static_cast<unsigned short>(b);
    }
    __syncthreads();
    float num = 0.f, den = 0.f;
    const float inv = 1.f / static_cast<float>(k * k);
    for (int e = threadIdx.x; e < FT * FT; e += blockDim.x) {
        const int ty = e / FT, tx = e - ty * FT;
        if (oy0 + ty >= Ho || ox0 + tx >= Wo) continue;
        int a = 0, b = 0;
        for (int d = 0; d < k; ++d) { a += hp[(ty + d) * FT + tx]; b += ht[(ty + d) * FT + tx]; }
        const float fa = static_cast<float>(a) * inv, fb = static_cast<float>(b) * inv;
        num += (fa - fb) * (fa - fb);
        den += fa * fa + fb * fb;
    }
    num = warp_sum(num);
    den = warp_sum(den);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[warp][0] = num; red[warp][1] = den; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float r = 0.f;
        for (int w8 = 0; w8 < 8; ++w8) r += red[w8][threadIdx.x];
        if (r != 0.f) atomicAdd(&acc[threadIdx.x], static_cast<double>(r));
    }
}

__global__ void fss_generic_finalize_kernel(double* __restrict__ acc, float* __restrict__ state, double cnt) {
    const float num = static_cast<float>(acc[0] / cnt), den = static_cast<float>(acc[1] / cnt);
    state[0] += 1.0f - num / (den + 1e-10f);
    state[1] += 1.0f;
    acc[0] = 0.0;
    acc[1] = 0.0;
}

}  // namespace p2i

using namespace p2i;

extern "C" int p2i_metrics_update(const float* pred, const float* target, int N, int H, int W, const float* thresholds, int n_thr,
                                  const int* scales, int n_scale, int apply_transform, double* scratch, float* state,
                                  void* stream) {
    P2I_CHECK_ARG(pred && target && scratch && state && thresholds && scales, "metrics_update: null pointer");
    P2I_CHECK_ARG(n_thr >= 1 && n_thr <= 4 && n_scale >= 1 && n_scale <= 4, "metrics_update: at most 4 thresholds and 4 scales");
    MetricParams p;
    p.N = N; p.H = H; p.W = W; p.n_thr = n_thr; p.n_scale = n_scale; p.apply_transform = apply_transform;
    for (int i = 0; i < 4; ++i) { p.thr[i] = i < n_thr ? thresholds[i] : 0.f; p.scale[i] = i < n_scale ? scales[i] : 1; }
    for (int i = 0; i < n_scale; ++i)
        P2I_CHECK_ARG(scales[i] == 1 || scales[i] == 2 || scales[i] == 4 || scales[i] == 8, "metrics_update: scale %d not in {1,2,4,8}",
                      scales[i]);
    dim3 grid(cdiv(W + 1, MT), cdiv(H + 1, MT), N);
    metrics_kernel<<<grid, 256, 0, as_stream(stream)>>>(pred, target, scratch, p);
    P2I_CHECK_LAUNCH("metrics_kernel");
    metrics_finalize_kernel<<<1, 64, 0, as_stream(stream)>>>(scratch, state, p);
    P2I_CHECK_LAUNCH("metrics_finalize_kernel");
    return P2I_OK;
}

extern "C" int p2i_fss_update(const float* pred, const float* target, int N, int H, int W, float threshold, int scale,
                              double* scratch, float* state, void* stream) {
    P2I_CHECK_ARG(pred && target && scratch && state, "fss_update: null pointer");
    P2I_CHECK_ARG(scale >= 1 && scale <= 32, "fss_update: box size %d outside 1..32", scale);
    P2I_CHECK_ARG(N >= 1 && N <= 65535, "fss_update: N = %d outside 1..65535 (split the batch)", N);
    const int pad = scale / 2;
    const int Ho = H + 2 * pad - scale + 1, Wo = W + 2 * pad - scale + 1;
    dim3 grid(cdiv(Wo, FT), cdiv(Ho, FT), N);
    fss_generic_kernel<<<grid, 256, 0, as_stream(stream)>>>(pred, target, scratch, N, H, W, threshold, scale);
    P2I_CHECK_LAUNCH("fss_generic_kernel");
    fss_generic_finalize_kernel<<<1, 1, 0, as_stream(stream)>>>(scratch, state, static_cast<double>(N) * Ho * Wo);
    P2I_CHECK_LAUNCH("fss_generic_finalize_kernel");
    return P2I_OK;
}
