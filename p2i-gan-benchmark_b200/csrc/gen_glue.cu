// HBM-bound glue of the generator trunk (p2igan_bench/models/p2igan.py:72-112): grouped stem,
// max-pool/duplicate pyramid, bilinear upsample + positional modulation, grouped 1x1 head + tanh,
// and NCHW f32 <-> NHWC bf16 layout changes.  Vectorised (16 B per thread) and coalesced.
#include "common.h"
#include "ptx.cuh"

namespace p2i {

// ------------------------------------------------------------------------------------------------
// Convsin: grouped 3x3 conv 16->64 (groups 4) + x.repeat_interleave(4, dim=1)   (p2igan.py:79)
// one thread per pixel, lanes along x; 64 bf16 outputs (128 B) written per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       __nv_bfloat16* __restrict__ y, int H, int W) {
    __shared__ float sw[64 * 36];
    for (int i = threadIdx.x; i < 64 * 36; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int b = blockIdx.z, yy = blockIdx.y;
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= W) return;
    const size_t HW = static_cast<size_t>(H) * W;
    const float* xb = x + static_cast<size_t>(b) * 16 * HW;
    uint4* out = reinterpret_cast<uint4*>(y + ((static_cast<size_t>(b) * H + yy) * W + xx) * 64);
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
        float in[36];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int iy = yy + ky - 1, ix = xx + kx - 1;
                    in[ci * 9 + ky * 3 + kx] =
                        (iy >= 0 && iy < H && ix >= 0 && ix < W) ? xb[(g * 4 + ci) * HW + static_cast<size_t>(iy) * W + ix] : 0.f;
                }
        float acc[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) {
            const float* wp = sw + (g * 16 + o) * 36;
            float a = in[(o >> 2) * 9 + 4];  // repeat_interleave: channel (16g+o)/4 = 4g + o/4, centre tap
#pragma unroll
            for (int k = 0; k < 36; ++k) a = fmaf(wp[k], in[k], a);
            acc[o] = a;
        }
        uint4 o0, o1;
        o0.x = pack_bf16x2(acc[0], acc[1]);   o0.y = pack_bf16x2(acc[2], acc[3]);
        o0.z = pack_bf16x2(acc[4], acc[5]);   o0.w = pack_bf16x2(acc[6], acc[7]);
        o1.x = pack_bf16x2(acc[8], acc[9]);   o1.y = pack_bf16x2(acc[10], acc[11]);
        o1.z = pack_bf16x2(acc[12], acc[13]); o1.w = pack_bf16x2(acc[14], acc[15]);
        out[g * 2] = o0;
        out[g * 2 + 1] = o1;
    }
}

// ------------------------------------------------------------------------------------------------
// DownsampleDuplicateChannels x3 (layer.py:200-214): out[c] = maxpool2(in)[c/2] per level, so
//   x4[c] = max over 4x4 of stem[c/4],  x8[c] = max over 8x8 of stem[c/8].
// One warp per 8x8 pixel block, lane = channel pair.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pyramid_fwd_kernel(const __nv_bfloat16* __restrict__ stem,
                                                          __nv_bfloat16* __restrict__ x4, __nv_bfloat16* __restrict__ x8,
                                                          int B, int H, int W) {
    const int lane = threadIdx.x & 31;
    const int wid = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int bw = W >> 3, bh = H >> 3;
    if (wid >= B * bh * bw) return;
    const int b = wid / (bh * bw), r = wid - b * bh * bw, by = r / bw, bx = r - by * bw;
    const uint32_t* in = reinterpret_cast<const uint32_t*>(stem) + ((static_cast<size_t>(b) * H + by * 8) * W + bx * 8) * 32 + lane;
    float2 m4[2][2];
#pragma unroll
    for (int sy = 0; sy < 2; ++sy)
#pragma unroll
        for (int sx = 0; sx < 2; ++sx) {
            float2 m = make_float2(-INFINITY, -INFINITY);
#pragma unroll
            for (int dy = 0; dy < 4; ++dy)
#pragma unroll
                for (int dx = 0; dx < 4; ++dx) {
                    const float2 v = unpack_bf16x2(__ldg(in + (static_cast<size_t>(sy * 4 + dy) * W + sx * 4 + dx) * 32));
                    m.x = fmaxf(m.x, v.x);
                    m.y = fmaxf(m.y, v.y);
                }
            m4[sy][sx] = m;
        }
    const int H4 = H >> 2, W4 = W >> 2;
#pragma unroll
    for (int sy = 0; sy < 2; ++sy)
#pragma unroll
        for (int sx = 0; sx < 2; ++sx) {
            const uint32_t a = pack_bf16x2(m4[sy][sx].x, m4[sy][sx].x), c = pack_bf16x2(m4[sy][sx].y, m4[sy][sx].y);
            uint4* o = reinterpret_cast<uint4*>(x4 + ((static_cast<size_t>(b) * H4 + by * 2 + sy) * W4 + bx * 2 + sx) * 256) + lane;
            *o = make_uint4(a, a, c, c);
        }
    const float mx = fmaxf(fmaxf(m4[0][0].x, m4[0][1].x), fmaxf(m4[1][0].x, m4[1][1].x));
    const float my = fmaxf(fmaxf(m4[0][0].y, m4[0][1].y), fmaxf(m4[1][0].y, m4[1][1].y));
    const uint32_t a = pack_bf16x2(mx, mx), c = pack_bf16x2(my, my);
    uint4* o = reinterpret_cast<uint4*>(x8 + ((static_cast<size_t>(b) * bh + by) * bw + bx) * 512) + lane * 2;
    o[0] = make_uint4(a, a, a, a);
    o[1] = make_uint4(c, c, c, c);
}

// ------------------------------------------------------------------------------------------------
// UPPos tail (layer.py:392-399) after the 1x1 projection was hoisted below the upsample:
//   out = relu(2*sigmoid(pos) * bilinear_x2(z) + bias) [+ skip]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upmod_fwd_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ pos,
                                                        const float* __restrict__ bias, const __nv_bfloat16* __restrict__ skip,
                                                        __nv_bfloat16* __restrict__ out, int B, int h, int w, int C) {
    const int cg = C >> 3;
    const int total = B * 4 * h * w * cg;                      // < 2^31 (checked on the host): 32-bit index arithmetic
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int pix = idx / cg, c8 = idx - pix * cg;
    const int H2 = 2 * h, W2 = 2 * w;
    const int r_ = pix / W2, X = pix - r_ * W2, b = r_ / H2, Y = r_ - b * H2;
    const float sy = (h > 1) ? static_cast<float>(h - 1) / static_cast<float>(H2 - 1) : 0.f;
    const float sx = (w > 1) ? static_cast<float>(w - 1) / static_cast<float>(W2 - 1) : 0.f;
    const float fy = sy * Y, fx = sx * X;
    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0;
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const float s = 2.f / (1.f + __expf(-pos[static_cast<size_t>(Y) * W2 + X]));
    const uint4* zb = reinterpret_cast<const uint4*>(z + static_cast<size_t>(b) * h * w * C) + c8;
    const uint4 a = __ldg(zb + (static_cast<size_t>(y0) * w + x0) * cg), bq = __ldg(zb + (static_cast<size_t>(y0) * w + x1) * cg);
    const uint4 c = __ldg(zb + (static_cast<size_t>(y1) * w + x0) * cg), d = __ldg(zb + (static_cast<size_t>(y1) * w + x1) * cg);
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
    const uint32_t cv[4] = {c.x, c.y, c.z, c.w}, dv[4] = {d.x, d.y, d.z, d.w};
    uint32_t sv[4] = {0, 0, 0, 0};
    const size_t o = (static_cast<size_t>(pix)) * cg + c8;
    if (skip) {
        const uint4 k = __ldg(reinterpret_cast<const uint4*>(skip) + o);
        sv[0] = k.x; sv[1] = k.y; sv[2] = k.z; sv[3] = k.w;
    }
    const float4 bia0 = __ldg(reinterpret_cast<const float4*>(bias) + c8 * 2);
    const float4 bia1 = __ldg(reinterpret_cast<const float4*>(bias) + c8 * 2 + 1);
    const float bb[8] = {bia0.x, bia0.y, bia0.z, bia0.w, bia1.x, bia1.y, bia1.z, bia1.w};
    uint32_t ov[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 fa = unpack_bf16x2(av[i]), fb = unpack_bf16x2(bv[i]), fc = unpack_bf16x2(cv[i]), fd = unpack_bf16x2(dv[i]);
        float r0 = fmaxf(fmaf(s, w00 * fa.x + w01 * fb.x + w10 * fc.x + w11 * fd.x, bb[2 * i]), 0.f);
        float r1 = fmaxf(fmaf(s, w00 * fa.y + w01 * fb.y + w10 * fc.y + w11 * fd.y, bb[2 * i + 1]), 0.f);
        if (skip) {
            const float2 fs = unpack_bf16x2(sv[i]);
            r0 += fs.x;
            r1 += fs.y;
        }
        ov[i] = pack_bf16x2(r0, r1);
    }
    reinterpret_cast<uint4*>(out)[o] = make_uint4(ov[0], ov[1], ov[2], ov[3]);
}

// ------------------------------------------------------------------------------------------------
// ConvsOut: 1x1 conv, groups 4, 64->16 (weights W [16,16,1] used directly, deconv_pytorch.py:126-127) + tanh.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) head_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                       float* __restrict__ out, float* __restrict__ pre, long long npix, int HW) {
    __shared__ float sw[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const int b = static_cast<int>(p / HW), pix = static_cast<int>(p - static_cast<long long>(b) * HW);
    const uint4* xp = reinterpret_cast<const uint4*>(x + p * 64);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint4 u0 = __ldg(xp + g * 2), u1 = __ldg(xp + g * 2 + 1);
        const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
        float in[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 f = unpack_bf16x2(uu[i]);
            in[2 * i] = f.x;
            in[2 * i + 1] = f.y;
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int oc = g * 4 + o;
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) a = fmaf(sw[oc * 16 + i], in[i], a);
            const size_t oi = (static_cast<size_t>(b) * 16 + oc) * HW + pix;
            if (pre) pre[oi] = a;
            out[oi] = tanhf(a);
        }
    }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int C, int HW, long long total) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = static_cast<int>(i % C);
    const long long p = i / C;
    const long long b = p / HW, pix = p - b * HW;
    y[i] = __float2bfloat16(x[(b * C + c) * HW + pix]);
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int C, int HW, long long total) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long pix = i % HW;
    const long long bc = i / HW;
    const long long b = bc / C, c = bc - b * C;
    y[i] = __bfloat162float(x[(b * HW + pix) * C + c]);
}

}  // namespace p2i

using namespace p2i;

extern "C" int p2i_stem_fwd(const float* x, const float* w, void* y, int B, int H, int W, void* stream) {
    P2I_CHECK_ARG(x && w && y && B > 0 && H > 0 && W > 0, "stem_fwd: bad arguments");
    dim3 grid(cdiv(W, 128), H, B);
    stem_fwd_kernel<<<grid, 128, 0, as_stream(stream)>>>(x, w, static_cast<__nv_bfloat16*>(y), H, W);
    P2I_CHECK_LAUNCH("stem_fwd_kernel");
    return P2I_OK;
}

extern "C" int p2i_pyramid_fwd(const void* stem, void* x4, void* x8, int B, int H, int W, void* stream) {
    P2I_CHECK_ARG(stem && x4 && x8, "pyramid_fwd: null pointer");
    P2I_CHECK_ARG(H % 8 == 0 && W % 8 == 0 && H > 0 && W > 0, "pyramid_fwd: H=%d W=%d must be multiples of 8", H, W);
    const int warps = B * (H / 8) * (W / 8);
    pyramid_fwd_kernel<<<cdiv(warps, 8), 256, 0, as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(stem), static_cast<__nv_bfloat16*>(x4), static_cast<__nv_bfloat16*>(x8), B, H, W);
    P2I_CHECK_LAUNCH("pyramid_fwd_kernel");
    return P2I_OK;
}

extern "C" int p2i_upmod_fwd(const void* z, const float* pos, const float* bias, const void* skip, void* out, int B, int h,
                             int w, int C, void* stream) {
    P2I_CHECK_ARG(z && pos && bias && out, "upmod_fwd: null pointer");
    P2I_CHECK_ARG(C % 8 == 0 && C > 0, "upmod_fwd: C=%d must be a multiple of 8", C);
    const long long total = static_cast<long long>(B) * 4 * h * w * (C / 8);
    P2I_CHECK_ARG(total < (1ll << 31), "upmod_fwd: tensor too large for 32-bit indexing");
    upmod_fwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(z), pos, bias, static_cast<const __nv_bfloat16*>(skip),
        static_cast<__nv_bfloat16*>(out), B, h, w, C);
    P2I_CHECK_LAUNCH("upmod_fwd_kernel");
    return P2I_OK;
}

extern "C" int p2i_head_fwd(const void* x, const float* w, float* out, float* pre, int B, int H, int W, void* stream) {
    P2I_CHECK_ARG(x && w && out, "head_fwd: null pointer");
    const long long npix = static_cast<long long>(B) * H * W;
    head_fwd_kernel<<<static_cast<unsigned>((npix + 127) / 128), 128, 0, as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), w, out, pre, npix, H * W);
    P2I_CHECK_LAUNCH("head_fwd_kernel");
    return P2I_OK;
}

extern "C" int p2i_nchw_f32_to_nhwc_bf16(const float* x, void* y, int B, int C, int H, int W, void* stream) {
    P2I_CHECK_ARG(x && y, "nchw_f32_to_nhwc_bf16: null pointer");
    const long long total = static_cast<long long>(B) * C * H * W;
    nchw_to_nhwc_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, as_stream(stream)>>>(
        x, static_cast<__nv_bfloat16*>(y), C, H * W, total);
    P2I_CHECK_LAUNCH("nchw_to_nhwc_kernel");
    return P2I_OK;
}

extern "C" int p2i_nhwc_bf16_to_nchw_f32(const void* x, float* y, int B, int C, int H, int W, void* stream) {
    P2I_CHECK_ARG(x && y, "nhwc_bf16_to_nchw_f32: null pointer");
    const long long total = static_cast<long long>(B) * C * H * W;
    nhwc_to_nchw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), y, C, H * W, total);
    P2I_CHECK_LAUNCH("nhwc_to_nchw_kernel");
    return P2I_OK;
}

// ------------------------------------------------------------------------------------------------
// Sliding-window blend of scripts/infer.py:237-245: frame l of an event of L frames is the mean of the predictions of
// every window [s, s+stride) that covers it (windows start every `step` frames; tail windows were padded by repeating
// the last frame and contribute only their valid part), scaled by output_scale and clipped at 0.
// preds f32 [n_win, stride, HW] -> out f32 [L, HW]
// ------------------------------------------------------------------------------------------------
namespace p2i {
__global__ void __launch_bounds__(256) window_blend_kernel(const float* __restrict__ preds, float* __restrict__ out, int L, int HW,
                                                           int stride, int step, int n_win, float scale) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(L) * HW) return;
    const int l = static_cast<int>(i / HW), pix = static_cast<int>(i - static_cast<long long>(l) * HW);
    float acc = 0.f, cnt = 0.f;
    for (int w = 0; w < n_win; ++w) {
        const int s = w * step;
        if (l >= s && l < s + stride) {
            acc += preds[(static_cast<size_t>(w) * stride + (l - s)) * HW + pix];
            cnt += 1.f;
        }
    }
    out[i] = fmaxf(acc / fmaxf(cnt, 1e-5f) * scale, 0.f);
}
}  // namespace p2i

extern "C" int p2i_window_blend(const float* preds, float* out, int L, int HW, int stride, int step, int n_win, float scale,
                                void* stream) {
    P2I_CHECK_ARG(preds && out && L > 0 && HW > 0 && stride > 0 && step > 0 && n_win > 0, "window_blend: bad arguments");
    const long long n = static_cast<long long>(L) * HW;
    p2i::window_blend_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, p2i::as_stream(stream)>>>(preds, out, L, HW, stride, step,
                                                                                                        n_win, scale);
    P2I_CHECK_LAUNCH("window_blend_kernel");
    return P2I_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Standalone DownsampleDuplicateChannels (layer.py:205-214) on the reference's own NCHW f32 layout: max_pool2d(2,2), then
// view [B*T, C/T] -> repeat_interleave(2, dim=1) -> view [B, 2C], i.e. input channel c feeds output channels 2c and 2c+1.
// The generator never calls this (its three levels are fused into pyramid_fwd/bwd); it exists so that the module's own
// forward/backward work as in the reference.  HBM-bound: 4 B read + 2 B written per input pixel.
namespace p2i {
__global__ void __launch_bounds__(256) downsample_dup_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                 long long n_out, int C, int h, int w) {
    // one thread per (b, c, yo, xo): reads a 2x2 window, writes the two duplicate channels
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const int xo = static_cast<int>(i % w);
    const int yo = static_cast<int>((i / w) % h);
    const long long bc = i / (static_cast<long long>(w) * h);
    const int c = static_cast<int>(bc % C);
    const long long b = bc / C;
    const float2* r0 = reinterpret_cast<const float2*>(x + ((bc * 2 * h + 2 * yo) * 2 * w + 2 * xo));
    const float2* r1 = reinterpret_cast<const float2*>(x + ((bc * 2 * h + 2 * yo + 1) * 2 * w + 2 * xo));
    const float2 a = __ldg(r0), d = __ldg(r1);
    const float m = fmaxf(fmaxf(a.x, a.y), fmaxf(d.x, d.y));
    const long long o = ((b * 2 * C + 2 * c) * h + yo) * w + xo;
    y[o] = m;
    y[o + static_cast<long long>(h) * w] = m;
}

__global__ void __launch_bounds__(256) downsample_dup_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                 float* __restrict__ dx, long long n_out, int C, int h, int w) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const int xo = static_cast<int>(i % w);
    const int yo = static_cast<int>((i / w) % h);
    const long long bc = i / (static_cast<long long>(w) * h);
    const int c = static_cast<int>(bc % C);
    const long long b = bc / C;
    const long long i0 = (bc * 2 * h + 2 * yo) * 2 * w + 2 * xo, i1 = i0 + 2 * w;
    const float2 a = __ldg(reinterpret_cast<const float2*>(x + i0)), d = __ldg(reinterpret_cast<const float2*>(x + i1));
    const long long o = ((b * 2 * C + 2 * c) * h + yo) * w + xo;
    const float g = dy[o] + dy[o + static_cast<long long>(h) * w];
    // max_pool2d backward: the FIRST maximum in window scan order receives the gradient (NaN counts as maximal, as in ATen)
    float m = a.x; int k = 0;
    if (a.y > m || a.y != a.y) { m = a.y; k = 1; }
    if (d.x > m || d.x != d.x) { m = d.x; k = 2; }
    if (d.y > m || d.y != d.y) { m = d.y; k = 3; }
    *reinterpret_cast<float2*>(dx + i0) = make_float2(k == 0 ? g : 0.f, k == 1 ? g : 0.f);
    *reinterpret_cast<float2*>(dx + i1) = make_float2(k == 2 ? g : 0.f, k == 3 ? g : 0.f);
}
}  // namespace p2i

extern "C" int p2i_downsample_dup_fwd(const float* x, float* y, int B, int C, int H, int W, void* stream) {
    P2I_CHECK_ARG(x && y, "downsample_dup_fwd: null pointer");
    P2I_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "downsample_dup_fwd: H=%d W=%d must be even", H, W);
    const long long n = static_cast<long long>(B) * C * (H / 2) * (W / 2);
    p2i::downsample_dup_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, p2i::as_stream(stream)>>>(x, y, n, C, H / 2, W / 2);
    P2I_CHECK_LAUNCH("downsample_dup_fwd_kernel");
    return P2I_OK;
}

extern "C" int p2i_downsample_dup_bwd(const float* x, const float* dy, float* dx, int B, int C, int H, int W, void* stream) {
    P2I_CHECK_ARG(x && dy && dx, "downsample_dup_bwd: null pointer");
    P2I_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "downsample_dup_bwd: H=%d W=%d must be even", H, W);
    const long long n = static_cast<long long>(B) * C * (H / 2) * (W / 2);
    p2i::downsample_dup_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, p2i::as_stream(stream)>>>(x, dy, dx, n, C, H / 2, W / 2);
    P2I_CHECK_LAUNCH("downsample_dup_bwd_kernel");
    return P2I_OK;
}
