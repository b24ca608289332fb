// Backward of the HBM-bound generator glue (see gen_glue.cu for the forward and the reference citations).
#include "common.h"
#include "ptx.cuh"

#include <cstdlib>

namespace p2i {

constexpr int UPMOD_CPT_DEFAULT = 16;     // channels per thread of upmod_bwd_hi_kernel (measured: profiles/r2_upmod_ab.txt)

// ------------------------------------------------------------------------------------------------
// ConvsOut + tanh backward (p2igan.py:109-111):  dz = dout * (1 - out^2);
//   dx[16g+i] = sum_{o in group g} dz[o] * w[o][i] ;  dw[o][i] += sum_pix dz[o] * x[16g+i]
// 4 threads per pixel (one per group) so that the 128-B pixel rows are read/written coalesced.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                       const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                       __nv_bfloat16* __restrict__ dx, float* __restrict__ dw, long long npix,
                                                       int HW) {
    __shared__ float sw[256];
    __shared__ float sdw[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { sw[i] = w[i]; sdw[i] = 0.f; }
    __syncthreads();
    const int g = threadIdx.x & 3;
    float acc[4][16];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[o][i] = 0.f;
    const int np32 = static_cast<int>(npix);                    // < 2^31 (checked on the host)
    for (int p = blockIdx.x * 64 + (threadIdx.x >> 2); p < np32; p += gridDim.x * 64) {
        const int b = p / HW, pix = p - b * HW;
        float dz[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const size_t oi = (static_cast<size_t>(b) * 16 + g * 4 + o) * HW + pix;
            const float t = out[oi];
            dz[o] = dout[oi] * (1.f - t * t);
        }
        const uint4* xp = reinterpret_cast<const uint4*>(x + p * 64 + g * 16);
        const uint4 u0 = __ldg(xp), u1 = __ldg(xp + 1);
        const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
        float in[16], d[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 f = unpack_bf16x2(uu[i]);
            in[2 * i] = f.x;
            in[2 * i + 1] = f.y;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float s = 0.f;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                s = fmaf(dz[o], sw[(g * 4 + o) * 16 + i], s);
                acc[o][i] = fmaf(dz[o], in[i], acc[o][i]);
            }
            d[i] = s;
        }
        uint4 o0, o1;
        o0.x = pack_bf16x2(d[0], d[1]);   o0.y = pack_bf16x2(d[2], d[3]);
        o0.z = pack_bf16x2(d[4], d[5]);   o0.w = pack_bf16x2(d[6], d[7]);
        o1.x = pack_bf16x2(d[8], d[9]);   o1.y = pack_bf16x2(d[10], d[11]);
        o1.z = pack_bf16x2(d[12], d[13]); o1.w = pack_bf16x2(d[14], d[15]);
        uint4* dp = reinterpret_cast<uint4*>(dx + p * 64 + g * 16);
        dp[0] = o0;
        dp[1] = o1;
    }
    // reduce over the 8 pixels of the warp (lane bits 2..4), then the 8 warps add into shared memory in turn
    // (shared-memory float atomics are CAS loops: 64-way contention made this the slowest part of the kernel)
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float a = acc[o][i];
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            a += __shfl_xor_sync(0xffffffffu, a, 8);
            a += __shfl_xor_sync(0xffffffffu, a, 16);
            acc[o][i] = a;
        }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int wv = 0; wv < 8; ++wv) {
        if (warp == wv && lane < 4) {
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int i = 0; i < 16; ++i) sdw[(g * 4 + o) * 16 + i] += acc[o][i];
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) atomicAdd(&dw[i], sdw[i]);
}

// ------------------------------------------------------------------------------------------------
// UPPos tail backward, pass 1 (high resolution): recompute pre = s*up(z)+bias,
//   dpre = (pre > 0) ? dout : 0 ;  g = s * dpre (bf16, input of the transposed upsample)
//   dbias[c] += sum dpre ;  dpos[Y,X] = s*(1 - s/2) * sum_{b,c} dpre * up(z)
// one thread per (pixel, 8 NQ channels); a pixel's threads are consecutive lanes of one warp.
// ------------------------------------------------------------------------------------------------
// NQ = 16-byte channel groups per thread (1, 2 or 4: 8 / 16 / 32 channels).  The bilinear geometry, sigmoid(pos) and the index
// arithmetic of a pixel are computed once per thread, so more channels per thread means fewer instructions (ncu on the 8-channel
// form: 47-66 % issue-slot utilisation, 212 instructions per 8 channels, ~130 of them per-pixel overhead) but also more registers
// and fewer resident warps; the host picks NQ (P2I_UPMOD_CPT, profiles/r2_upmod_ab.txt).
template <int NQ>
__global__ void __launch_bounds__(256, NQ == 4 ? 2 : 3) upmod_bwd_hi_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ pos,
                                                           const float* __restrict__ bias, const __nv_bfloat16* __restrict__ dout,
                                                           __nv_bfloat16* __restrict__ gout, float* __restrict__ dbias,
                                                           float* __restrict__ dpos, int B, int h, int w, int C) {
    extern __shared__ float s_db[];   // [C]
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_db[i] = 0.f;
    __syncthreads();
    const int cg = C >> 3;                                      // 16-byte channel groups per pixel
    const int cq = cg / NQ;                                     // threads per pixel (power of two <= 32, checked on the host)
    const int total = B * 4 * h * w * cq;                       // < 2^31 (checked on the host): 32-bit index arithmetic
    const int H2 = 2 * h, W2 = 2 * w;
    const float sy = (h > 1) ? static_cast<float>(h - 1) / static_cast<float>(H2 - 1) : 0.f;
    const float sx = (w > 1) ? static_cast<float>(w - 1) / static_cast<float>(W2 - 1) : 0.f;
    float dbv[NQ][8];
#pragma unroll
    for (int a = 0; a < NQ; ++a)
#pragma unroll
        for (int k = 0; k < 8; ++k) dbv[a][k] = 0.f;
    // the trip count is warp-uniform (the ds reduction below shuffles with a full mask); lanes past the end carry ds = 0
    for (int idx0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); idx0 < total; idx0 += gridDim.x * blockDim.x) {
        const int idx = idx0 + (threadIdx.x & 31);
        const bool valid = idx < total;
        float ds = 0.f, s = 0.f;
        int Y = 0, X = 0;
        if (valid) {
        const int q4 = idx & (cq - 1);
        const int pix = idx / cq;
        const int r_ = pix / W2, b = r_ / H2;
        X = pix - r_ * W2;
        Y = r_ - b * H2;
        const float fy = sy * Y, fx = sx * X;
        const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
        const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
        const float ly = fy - y0, lx = fx - x0;
        const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
        s = 2.f / (1.f + __expf(-pos[static_cast<size_t>(Y) * W2 + X]));
        const uint4* zb = reinterpret_cast<const uint4*>(z + static_cast<size_t>(b) * h * w * C) + q4 * NQ;
        const uint4* za = zb + (static_cast<size_t>(y0) * w + x0) * cg;
        const uint4* zbq = zb + (static_cast<size_t>(y0) * w + x1) * cg;
        const uint4* zc = zb + (static_cast<size_t>(y1) * w + x0) * cg;
        const uint4* zd = zb + (static_cast<size_t>(y1) * w + x1) * cg;
        const size_t o = static_cast<size_t>(pix) * cg + q4 * NQ;
#pragma unroll
        for (int a4 = 0; a4 < NQ; ++a4) {
            const uint4 a = __ldg(za + a4), bq = __ldg(zbq + a4), c = __ldg(zc + a4), d = __ldg(zd + a4);
            const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
            const uint32_t cv[4] = {c.x, c.y, c.z, c.w}, dv[4] = {d.x, d.y, d.z, d.w};
            const uint4 gq = __ldg(reinterpret_cast<const uint4*>(dout) + o + a4);
            const uint32_t gv[4] = {gq.x, gq.y, gq.z, gq.w};
            const float4 bia0 = __ldg(reinterpret_cast<const float4*>(bias) + (q4 * NQ + a4) * 2);
            const float4 bia1 = __ldg(reinterpret_cast<const float4*>(bias) + (q4 * NQ + a4) * 2 + 1);
            const float bb[8] = {bia0.x, bia0.y, bia0.z, bia0.w, bia1.x, bia1.y, bia1.z, bia1.w};
            uint32_t ov[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 fa = unpack_bf16x2(av[i]), fb = unpack_bf16x2(bv[i]), fc = unpack_bf16x2(cv[i]), fd = unpack_bf16x2(dv[i]);
                const float2 gg = unpack_bf16x2(gv[i]);
                const float u0 = w00 * fa.x + w01 * fb.x + w10 * fc.x + w11 * fd.x;
                const float u1 = w00 * fa.y + w01 * fb.y + w10 * fc.y + w11 * fd.y;
                const float d0 = (fmaf(s, u0, bb[2 * i]) > 0.f) ? gg.x : 0.f;
                const float d1 = (fmaf(s, u1, bb[2 * i + 1]) > 0.f) ? gg.y : 0.f;
                ds += d0 * u0 + d1 * u1;
                dbv[a4][2 * i] += d0;  // the thread keeps its channel slice for the whole loop (stride % cq == 0)
                dbv[a4][2 * i + 1] += d1;
                ov[i] = pack_bf16x2(s * d0, s * d1);
            }
            reinterpret_cast<uint4*>(gout)[o + a4] = make_uint4(ov[0], ov[1], ov[2], ov[3]);
        }
        }
        // reduce ds over the pixel's cq consecutive threads (cq is a power of two <= 32: groups never straddle a warp)
        for (int off = cq >> 1; off > 0; off >>= 1) ds += __shfl_xor_sync(0xffffffffu, ds, off);
        if (valid && (threadIdx.x & (cq - 1)) == 0 && ds != 0.f)
            atomicAdd(&dpos[static_cast<size_t>(Y) * W2 + X], ds * s * (1.f - 0.5f * s));
    }
    // bias gradient: registers -> lanes sharing a channel quad folded by shuffles -> shared -> global
    {
        const int q4 = static_cast<int>((blockIdx.x * blockDim.x + threadIdx.x) & (cq - 1));
        for (int off = 16; off >= cq; off >>= 1) {
#pragma unroll
            for (int a = 0; a < NQ; ++a)
#pragma unroll
                for (int k = 0; k < 8; ++k) dbv[a][k] += __shfl_xor_sync(0xffffffffu, dbv[a][k], off);
        }
        if ((threadIdx.x & 31) < cq) {
#pragma unroll
            for (int a = 0; a < NQ; ++a)
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (dbv[a][k] != 0.f) atomicAdd(&s_db[(q4 * NQ + a) * 8 + k], dbv[a][k]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x)
        if (s_db[i] != 0.f) atomicAdd(&dbias[i], s_db[i]);
}

// pass 2 (low resolution): dz = transpose(bilinear x2, align_corners) applied to g.  Gather form: every low-res pixel visits
// the high-res pixels whose interpolation footprint contains it.  One thread per (pixel, 32 channels); the horizontal weights of
// the (at most 10) candidate columns are computed once per thread, the vertical weight once per candidate row.
__global__ void __launch_bounds__(256, 2) upmod_bwd_lo_kernel(const __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ dz,
                                                           int B, int h, int w, int C) {
    const int cg = C >> 3, cq = C >> 5;
    const int total = B * h * w * cq;                          // < 2^31 (checked on the host)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int pix = idx / cq, q4 = idx - pix * cq;
    const int r_ = pix / w, x = pix - r_ * w, b = r_ / h, y = r_ - b * h;
    const int H2 = 2 * h, W2 = 2 * w;
    const float sy = (h > 1) ? static_cast<float>(h - 1) / static_cast<float>(H2 - 1) : 0.f;
    const float sx = (w > 1) ? static_cast<float>(w - 1) / static_cast<float>(W2 - 1) : 0.f;
    // candidate high-res rows / columns: those with floor(s*Y) in {y-1, y}  (a window of at most 7 for a x2 upsample)
    int Ya = (sy > 0.f) ? static_cast<int>(floorf((y - 1) / sy)) - 1 : 0, Yb = (sy > 0.f) ? static_cast<int>(ceilf((y + 1) / sy)) + 1 : H2 - 1;
    int Xa = (sx > 0.f) ? static_cast<int>(floorf((x - 1) / sx)) - 1 : 0, Xb = (sx > 0.f) ? static_cast<int>(ceilf((x + 1) / sx)) + 1 : W2 - 1;
    Ya = Ya < 0 ? 0 : Ya; Xa = Xa < 0 ? 0 : Xa;
    Yb = Yb > H2 - 1 ? H2 - 1 : Yb; Xb = Xb > W2 - 1 ? W2 - 1 : Xb;
    auto weight = [](int P, float sc, int n, int p) -> float {       // weight of low-res index p in high-res index P
        const float f = sc * P;
        const int p0 = static_cast<int>(f);
        const int p1 = p0 + (p0 < n - 1 ? 1 : 0);
        const float l = f - p0;
        return (p0 == p ? 1.f - l : 0.f) + (p1 == p ? l : 0.f);
    };
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[a][k] = 0.f;
    const uint4* gb = reinterpret_cast<const uint4*>(g + static_cast<size_t>(b) * H2 * W2 * C) + q4 * 4;
    if (Xb - Xa < 10) {                                        // always, for w >= 2: the window is 4 + 2 / (w - 1) wide plus margins
        float wx[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) wx[k] = (Xa + k <= Xb) ? weight(Xa + k, sx, w, x) : 0.f;
        for (int Y = Ya; Y <= Yb; ++Y) {
            const float wy = weight(Y, sy, h, y);
            if (wy == 0.f) continue;
            const uint4* row = gb + (static_cast<size_t>(Y) * W2 + Xa) * cg;
#pragma unroll
            for (int k = 0; k < 10; ++k) {
                if (wx[k] == 0.f) continue;
                const float wt = wy * wx[k];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const uint4 q = __ldg(row + static_cast<size_t>(k) * cg + a);
                    const uint32_t qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 f = unpack_bf16x2(qv[i]);
                        acc[a][2 * i] = fmaf(wt, f.x, acc[a][2 * i]);
                        acc[a][2 * i + 1] = fmaf(wt, f.y, acc[a][2 * i + 1]);
                    }
                }
            }
        }
    } else {                                                   // degenerate geometry (w == 1): every column is a candidate
        for (int Y = Ya; Y <= Yb; ++Y) {
            const float wy = weight(Y, sy, h, y);
            if (wy == 0.f) continue;
            for (int X = Xa; X <= Xb; ++X) {
                const float wxv = weight(X, sx, w, x);
                if (wxv == 0.f) continue;
                const float wt = wy * wxv;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const uint4 q = __ldg(gb + (static_cast<size_t>(Y) * W2 + X) * cg + a);
                    const uint32_t qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 f = unpack_bf16x2(qv[i]);
                        acc[a][2 * i] = fmaf(wt, f.x, acc[a][2 * i]);
                        acc[a][2 * i + 1] = fmaf(wt, f.y, acc[a][2 * i + 1]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
        reinterpret_cast<uint4*>(dz)[static_cast<size_t>(pix) * cg + q4 * 4 + a] =
            make_uint4(pack_bf16x2(acc[a][0], acc[a][1]), pack_bf16x2(acc[a][2], acc[a][3]), pack_bf16x2(acc[a][4], acc[a][5]),
                       pack_bf16x2(acc[a][6], acc[a][7]));
}

// ------------------------------------------------------------------------------------------------
// Pyramid backward: route d(x4), d(x8) to the arg-max pixel of each 4x4 / 8x8 block of the stem output.
// Ties follow nested 2x2 max-pools (first element in window order wins at every level, as ATen's
// max_pool2d backward does).  One warp per 8x8 block, lane = channel pair; writes all 64 pixels.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void argmax2x2(const float (&v)[4], const int (&id)[4], float& mv, int& mi) {
    mv = v[0]; mi = id[0];
#pragma unroll
    for (int k = 1; k < 4; ++k)
        if (v[k] > mv) { mv = v[k]; mi = id[k]; }
}

__global__ void __launch_bounds__(256) pyramid_bwd_kernel(const __nv_bfloat16* __restrict__ stem, const __nv_bfloat16* __restrict__ dx4,
                                                          const __nv_bfloat16* __restrict__ dx8, __nv_bfloat16* __restrict__ dstem,
                                                          int B, int H, int W) {
    const int lane = threadIdx.x & 31;
    const int wid = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int bw = W >> 3, bh = H >> 3;
    if (wid >= B * bh * bw) return;
    const int b = wid / (bh * bw), r = wid - b * bh * bw, by = r / bw, bx = r - by * bw;
    const size_t base = ((static_cast<size_t>(b) * H + by * 8) * W + bx * 8) * 32 + lane;
    const uint32_t* in = reinterpret_cast<const uint32_t*>(stem) + base;
    uint32_t* out = reinterpret_cast<uint32_t*>(dstem) + base;
    // level-1 winners (2x2) for both channels of the pair
    float v1[2][16];
    int i1[2][16];
#pragma unroll
    for (int qy = 0; qy < 4; ++qy)
#pragma unroll
        for (int qx = 0; qx < 4; ++qx) {
            float a[2][4];
            int id[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int py = qy * 2 + (k >> 1), px = qx * 2 + (k & 1);
                const float2 f = unpack_bf16x2(__ldg(in + (static_cast<size_t>(py) * W + px) * 32));
                a[0][k] = f.x; a[1][k] = f.y;
                id[k] = py * 8 + px;
            }
            argmax2x2(a[0], id, v1[0][qy * 4 + qx], i1[0][qy * 4 + qx]);
            argmax2x2(a[1], id, v1[1][qy * 4 + qx], i1[1][qy * 4 + qx]);
        }
    // gradients per channel: d4 (four 4x4 blocks) = sum of 4 duplicated channels, d8 = sum of 8
    const int H4 = H >> 2, W4 = W >> 2;
    float g4[2][4], g8[2];
#pragma unroll
    for (int sy = 0; sy < 2; ++sy)
#pragma unroll
        for (int sx = 0; sx < 2; ++sx) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(dx4 + ((static_cast<size_t>(b) * H4 + by * 2 + sy) * W4 + bx * 2 + sx) * 256) + lane);
            const float2 a = unpack_bf16x2(q.x), c = unpack_bf16x2(q.y), d = unpack_bf16x2(q.z), e = unpack_bf16x2(q.w);
            g4[0][sy * 2 + sx] = a.x + a.y + c.x + c.y;
            g4[1][sy * 2 + sx] = d.x + d.y + e.x + e.y;
        }
    {
        const uint4* q8 = reinterpret_cast<const uint4*>(dx8 + ((static_cast<size_t>(b) * bh + by) * bw + bx) * 512) + lane * 2;
        const uint4 q0 = __ldg(q8), q1 = __ldg(q8 + 1);
        const uint32_t u0[4] = {q0.x, q0.y, q0.z, q0.w}, u1[4] = {q1.x, q1.y, q1.z, q1.w};
        g8[0] = 0.f; g8[1] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f0 = unpack_bf16x2(u0[i]), f1 = unpack_bf16x2(u1[i]);
            g8[0] += f0.x + f0.y;
            g8[1] += f1.x + f1.y;
        }
    }
    int i2[2][4], i3[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        float v2[4];
#pragma unroll
        for (int sy = 0; sy < 2; ++sy)
#pragma unroll
            for (int sx = 0; sx < 2; ++sx) {
                float a[4];
                int id[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int q = (sy * 2 + (k >> 1)) * 4 + sx * 2 + (k & 1);
                    a[k] = v1[ch][q];
                    id[k] = i1[ch][q];
                }
                argmax2x2(a, id, v2[sy * 2 + sx], i2[ch][sy * 2 + sx]);
            }
        float v3;
        argmax2x2(v2, i2[ch], v3, i3[ch]);
    }
#pragma unroll
    for (int k = 0; k < 64; ++k) {
        float ga = (i3[0] == k) ? g8[0] : 0.f, gb = (i3[1] == k) ? g8[1] : 0.f;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            ga += (i2[0][s] == k) ? g4[0][s] : 0.f;
            gb += (i2[1][s] == k) ? g4[1][s] : 0.f;
        }
        out[(static_cast<size_t>(k >> 3) * W + (k & 7)) * 32] = pack_bf16x2(ga, gb);
    }
}

// ------------------------------------------------------------------------------------------------
// Stem backward (p2igan.py:79).  dy [B,H,W,64] bf16 ->
//   dx [B,16,H,W] f32 : transposed grouped conv + the repeat_interleave(4) path
//   dw [64,4,9]   f32 : sum_pix dy[pix][oc] * x[g*4+icl][pix + tap]   (register-tiled, atomics at the end)
// ------------------------------------------------------------------------------------------------
// thread = (4 consecutive pixels of a row, group g): every LDS.128 of weights (4 input channels of one (tap, oc)) feeds
// 16 FMAs (4 pixels x 4 ci) -- with one pixel per thread the kernel was shared-memory (MIO) throttled at 4 FMAs per LDS.
// The 6 gradient pixels a 3-tap row window needs are loaded once per (ky, oc half) as 16-byte vectors; the four group
// lanes of a pixel quad read the 4 x 32 bytes of each 128-byte pixel row.  Weights: [g][tap][oc][ci], group pitch 580
// floats (disjoint banks for the four groups of a warp).  Requires W % 4 == 0 (else the one-pixel variant below).
__global__ void __launch_bounds__(128) stem_bwd_dx4_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ w,
                                                           float* __restrict__ dx, int H, int W) {
    __shared__ __align__(16) float sw[4 * 580];
    for (int i = threadIdx.x; i < 4 * 576; i += blockDim.x) {
        const int ci = i & 3, oc = (i >> 2) & 15, r = i >> 6, tap = r % 9, g = r / 9;
        sw[g * 580 + (tap * 16 + oc) * 4 + ci] = w[((g * 16 + oc) * 4 + ci) * 9 + tap];
    }
    __syncthreads();
    const int b = blockIdx.z;
    const int g = threadIdx.x & 3;
    const int x0 = (blockIdx.x * 32 + (threadIdx.x >> 2)) * 4;
    if (x0 >= W) return;
    const size_t HW = static_cast<size_t>(H) * W;
    const float4* wg = reinterpret_cast<const float4*>(sw + g * 580);
    for (int yy = blockIdx.y * 4; yy < H && yy < blockIdx.y * 4 + 4; ++yy) {
        float acc[4][4];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[p][c] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int oy = yy - ky + 1;
            if (oy < 0 || oy >= H) continue;
            const __nv_bfloat16* rowp = dy + ((static_cast<size_t>(b) * H + oy) * W) * 64 + g * 16;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                // gradient pixels x0-1 .. x0+4 (window position q = 0..5), output channels 8*half .. 8*half+7 of the group
                float d[6][8];
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const int ox = x0 - 1 + q;
                    uint4 u = make_uint4(0u, 0u, 0u, 0u);
                    if (ox >= 0 && ox < W) u = __ldg(reinterpret_cast<const uint4*>(rowp + static_cast<size_t>(ox) * 64 + half * 8));
                    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 f = unpack_bf16x2(uu[k]);
                        d[q][2 * k] = f.x; d[q][2 * k + 1] = f.y;
                    }
                }
                // dx[x0+p] += sum_kx dy[x0+p - kx + 1] * w[kx]  ->  window position q = p + 2 - kx
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        const float4 w4 = wg[(ky * 3 + kx) * 16 + half * 8 + o];
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const float v = d[p + 2 - kx][o];
                            acc[p][0] = fmaf(v, w4.x, acc[p][0]);
                            acc[p][1] = fmaf(v, w4.y, acc[p][1]);
                            acc[p][2] = fmaf(v, w4.z, acc[p][2]);
                            acc[p][3] = fmaf(v, w4.w, acc[p][3]);
                        }
                    }
                if (ky == 1) {          // repeat_interleave: output channel oc of the group feeds input channel oc / 4 (centre tap)
#pragma unroll
                    for (int o = 0; o < 8; ++o)
#pragma unroll
                        for (int p = 0; p < 4; ++p) acc[p][(half * 8 + o) >> 2] += d[p + 1][o];
                }
            }
        }
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
            *reinterpret_cast<float4*>(dx + (static_cast<size_t>(b) * 16 + g * 4 + ci) * HW + static_cast<size_t>(yy) * W + x0) =
                make_float4(acc[0][ci], acc[1][ci], acc[2][ci], acc[3][ci]);
    }
}

// lane = (pixel, group): the four lanes of a pixel read its 128 contiguous gradient bytes, so a warp request covers
// 8 pixels x 128 B = 8 cache lines (one pixel per lane with a fixed group touched 32).  Weights: [g][tap][oc][ci] with a
// group pitch of 580 floats so that the four groups of a warp read disjoint banks (LDS.128 broadcast per group).
__global__ void __launch_bounds__(128) stem_bwd_dx_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ w,
                                                          float* __restrict__ dx, int H, int W) {
    __shared__ __align__(16) float sw[4 * 580];
    for (int i = threadIdx.x; i < 4 * 576; i += blockDim.x) {
        const int ci = i & 3, oc = (i >> 2) & 15, r = i >> 6, tap = r % 9, g = r / 9;
        sw[g * 580 + (tap * 16 + oc) * 4 + ci] = w[((g * 16 + oc) * 4 + ci) * 9 + tap];
    }
    __syncthreads();
    const int b = blockIdx.z;
    const int g = threadIdx.x & 3;
    const int xx = blockIdx.x * 32 + (threadIdx.x >> 2);
    if (xx >= W) return;
    const size_t HW = static_cast<size_t>(H) * W;
    for (int yy = blockIdx.y * 4; yy < H && yy < blockIdx.y * 4 + 4; ++yy) {      // 4 rows per block amortise the weight load
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int oy = yy - ky + 1;
        if (oy < 0 || oy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ox = xx - kx + 1;
            if (ox < 0 || ox >= W) continue;
            const uint4* dp = reinterpret_cast<const uint4*>(dy + ((static_cast<size_t>(b) * H + oy) * W + ox) * 64 + g * 16);
            const uint4 u0 = __ldg(dp), u1 = __ldg(dp + 1);
            const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
            const float4* wt = reinterpret_cast<const float4*>(sw + g * 580) + (ky * 3 + kx) * 16;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 f = unpack_bf16x2(uu[i]);
                const float4 wa = wt[2 * i], wb = wt[2 * i + 1];
                acc[0] = fmaf(f.x, wa.x, fmaf(f.y, wb.x, acc[0]));
                acc[1] = fmaf(f.x, wa.y, fmaf(f.y, wb.y, acc[1]));
                acc[2] = fmaf(f.x, wa.z, fmaf(f.y, wb.z, acc[2]));
                acc[3] = fmaf(f.x, wa.w, fmaf(f.y, wb.w, acc[3]));
                if (ky == 1 && kx == 1) {   // repeat_interleave: output channel oc feeds input channel oc/4
                    acc[(2 * i) >> 2] += f.x;
                    acc[(2 * i + 1) >> 2] += f.y;
                }
            }
        }
    }
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) dx[(static_cast<size_t>(b) * 16 + g * 4 + ci) * HW + static_cast<size_t>(yy) * W + xx] = acc[ci];
    }
}

// Weight gradient, tiled: a block stages one image-row segment (64 pixels: the 16 x 3 x 66 input patch and the
// 64 x 64 gradient tile) in shared memory; thread = (output channel, input channel of its group) with the 9 taps in
// registers across all of the block's tiles; one global atomic per (block, weight) at the end.
__global__ void __launch_bounds__(256) stem_bwd_dw2_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x,
                                                           float* __restrict__ dw, int B, int H, int W) {
    __shared__ __align__(16) float sx[16][3][68];
    __shared__ __align__(16) __nv_bfloat16 sdy[64][64];
    const int oc = threadIdx.x & 63, icl = threadIdx.x >> 6, g = oc >> 4;
    const int ntx = (W + 63) >> 6;
    const int ntiles = B * H * ntx;
    const size_t HW = static_cast<size_t>(H) * W;
    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tx = tile % ntx, r1 = tile / ntx, y = r1 % H, b = r1 / H;
        const int x0 = tx << 6;
        __syncthreads();
        {   // 48 patch rows (16 channels x 3 image rows) of 66 floats: one warp per row, no per-element div/mod
            const int lane = threadIdx.x & 31;
            for (int row = threadIdx.x >> 5; row < 48; row += 8) {
                const int ch = row / 3, r = row - ch * 3;
                const int yi = y + r - 1;
                const bool rowok = yi >= 0 && yi < H;
                const float* src = x + (static_cast<size_t>(b) * 16 + ch) * HW + static_cast<size_t>(rowok ? yi : 0) * W + x0 - 1;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int c = lane + 32 * k;
                    if (c < 66) {
                        const int xi = x0 + c - 1;
                        sx[ch][r][c] = (rowok && xi >= 0 && xi < W) ? __ldg(src + c) : 0.f;
                    }
                }
            }
        }
        for (int e = threadIdx.x; e < 64 * 8; e += 256) {
            const int px = e >> 3, q = e & 7;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (x0 + px < W) v = __ldg(reinterpret_cast<const uint4*>(dy + ((static_cast<size_t>(b) * H + y) * W + x0 + px) * 64) + q);
            *reinterpret_cast<uint4*>(&sdy[px][q * 8]) = v;
        }
        __syncthreads();
        const float* xr = &sx[g * 4 + icl][0][0];
#pragma unroll 1
        for (int p0 = 0; p0 < 64; p0 += 8) {
            float d[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) d[p] = __bfloat162float(sdy[p0 + p][oc]);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float4 a = *reinterpret_cast<const float4*>(xr + ky * 68 + p0);
                const float4 c4 = *reinterpret_cast<const float4*>(xr + ky * 68 + p0 + 4);
                const float2 e2 = *reinterpret_cast<const float2*>(xr + ky * 68 + p0 + 8);
                const float v[10] = {a.x, a.y, a.z, a.w, c4.x, c4.y, c4.z, c4.w, e2.x, e2.y};
#pragma unroll
                for (int p = 0; p < 8; ++p)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) acc[ky * 3 + kx] = fmaf(d[p], v[p + kx], acc[ky * 3 + kx]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k)
        if (acc[k] != 0.f) atomicAdd(&dw[(oc * 4 + icl) * 9 + k], acc[k]);
}

}  // namespace p2i

using namespace p2i;

extern "C" int p2i_head_bwd(const float* dout, const float* out, const void* x, const float* w, void* dx, float* dw, int B,
                            int H, int W, void* stream) {
    P2I_CHECK_ARG(dout && out && x && w && dx && dw, "head_bwd: null pointer");
    const long long npix = static_cast<long long>(B) * H * W;
    P2I_CHECK_ARG(npix * 64 < (1ll << 31), "head_bwd: tensor too large for 32-bit indexing");
    long long blocks = (npix + 63) / 64;
    if (blocks > sm_count() * 4) blocks = sm_count() * 4;
    head_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(
        dout, out, static_cast<const __nv_bfloat16*>(x), w, static_cast<__nv_bfloat16*>(dx), dw, npix, H * W);
    P2I_CHECK_LAUNCH("head_bwd_kernel");
    return P2I_OK;
}

extern "C" int p2i_upmod_bwd(const void* z, const float* pos, const float* bias, const void* dout, void* g_scratch, void* dz,
                             float* dbias, float* dpos, int B, int h, int w, int C, void* stream) {
    P2I_CHECK_ARG(z && pos && bias && dout && g_scratch && dz && dbias && dpos, "upmod_bwd: null pointer");
    P2I_CHECK_ARG(C % 64 == 0 && (C & (C - 1)) == 0 && C <= 1024, "upmod_bwd: C=%d must be a power of two in [64, 1024]", C);
    P2I_CHECK_ARG(static_cast<long long>(B) * 4 * h * w * (C / 8) < (1ll << 31), "upmod_bwd: tensor too large for 32-bit indexing");
    // channels per thread of the high-resolution pass: 8, 16 or 32 (P2I_UPMOD_CPT; read per call so that an A/B can switch it),
    // raised until a pixel's threads fit one warp
    int cpt = UPMOD_CPT_DEFAULT;
    if (const char* e = getenv("P2I_UPMOD_CPT")) { const int v = atoi(e); if (v == 8 || v == 16 || v == 32) cpt = v; }
    while (C / cpt > 32) cpt *= 2;
    const long long total = static_cast<long long>(B) * 4 * h * w * (C / cpt);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    // the grid stride (blocks * 256) is a multiple of C / cpt, so a thread keeps its channel slice for its whole loop
    const auto hi = cpt == 8 ? upmod_bwd_hi_kernel<1> : cpt == 16 ? upmod_bwd_hi_kernel<2> : upmod_bwd_hi_kernel<4>;
    hi<<<static_cast<unsigned>(blocks), 256, C * sizeof(float), as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(z), pos, bias, static_cast<const __nv_bfloat16*>(dout),
        static_cast<__nv_bfloat16*>(g_scratch), dbias, dpos, B, h, w, C);
    P2I_CHECK_LAUNCH("upmod_bwd_hi_kernel");
    const long long tl = static_cast<long long>(B) * h * w * (C / 32);
    upmod_bwd_lo_kernel<<<static_cast<unsigned>((tl + 255) / 256), 256, 0, as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(g_scratch), static_cast<__nv_bfloat16*>(dz), B, h, w, C);
    P2I_CHECK_LAUNCH("upmod_bwd_lo_kernel");
    return P2I_OK;
}

extern "C" int p2i_pyramid_bwd(const void* stem, const void* dx4, const void* dx8, void* dstem, int B, int H, int W,
                               void* stream) {
    P2I_CHECK_ARG(stem && dx4 && dx8 && dstem, "pyramid_bwd: null pointer");
    P2I_CHECK_ARG(H % 8 == 0 && W % 8 == 0, "pyramid_bwd: H, W must be multiples of 8");
    const int warps = B * (H / 8) * (W / 8);
    pyramid_bwd_kernel<<<cdiv(warps, 8), 256, 0, as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(stem), static_cast<const __nv_bfloat16*>(dx4), static_cast<const __nv_bfloat16*>(dx8),
        static_cast<__nv_bfloat16*>(dstem), B, H, W);
    P2I_CHECK_LAUNCH("pyramid_bwd_kernel");
    return P2I_OK;
}

extern "C" int p2i_stem_bwd(const void* dy, const float* x, const float* w, float* dx, float* dw, int B, int H, int W,
                            void* stream) {
    P2I_CHECK_ARG(dy && x && w && (dx || dw), "stem_bwd: null pointer");
    if (dx && W % 4 == 0) {       // dx == NULL / dw == NULL: only the other gradient (two independent kernels)
        dim3 grid(cdiv(W, 128), cdiv(H, 4), B);
        stem_bwd_dx4_kernel<<<grid, 128, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(dy), w, dx, H, W);
        P2I_CHECK_LAUNCH("stem_bwd_dx4_kernel");
    } else if (dx) {
        dim3 grid(cdiv(W, 32), cdiv(H, 4), B);
        stem_bwd_dx_kernel<<<grid, 128, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(dy), w, dx, H, W);
        P2I_CHECK_LAUNCH("stem_bwd_dx_kernel");
    }
    if (dw) {
        long long tiles = static_cast<long long>(B) * H * ((W + 63) / 64);
        P2I_CHECK_ARG(tiles < (1ll << 31), "stem_bwd: tensor too large");
        if (tiles > sm_count() * 4) tiles = sm_count() * 4;
        stem_bwd_dw2_kernel<<<static_cast<unsigned>(tiles), 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(dy), x, dw, B, H, W);
        P2I_CHECK_LAUNCH("stem_bwd_dw2_kernel");
    }
    return P2I_OK;
}
