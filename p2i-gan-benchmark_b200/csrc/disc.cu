// CUDA-core pieces of the dual-branch patch discriminator (p2igan_bench/models/p2igan.py:115-173):
// spectral normalisation (torch.nn.utils.spectral_norm, call sites layer.py:402-407, p2igan.py:141), weight
// packing into the implicit-GEMM operand layouts, the thin first/last layers (Cin = 1 / Cout = 1) and the fused
// tail (1x1x1 conv + mean over depth + bilinear resize + sigmoid(alpha2d) fusion, p2igan.py:164-173).
// The wide layers run on the tensor cores (conv_igemm.cu / conv_wgrad.cu).
#include "common.h"
#include "ptx.cuh"

namespace p2i {

__device__ __forceinline__ float blk_sum(float v, float* sh) {   // result valid in ALL threads
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float t = (lane < nw) ? sh[lane] : 0.f;
    t = warp_sum(t);
    return t;
}

// ------------------------------------------------------------------------------------------------
// Spectral norm (torch.nn.utils.spectral_norm pre-forward hook), all layers in three grid-parallel launches:
//   K1  t1 = W^T u            (training only)     ss1 += |t1|^2
//   K2  t2' = W t1  (eval: W v)                   ss2 += |t2'|^2
//   K3  n1 = max(|t1|,eps); v = t1/n1; t2 = t2'/n1; n2 = max(|t2|,eps); u = t2/n2; sigma = u . t2   (eval: sigma = u . t2')
// scratch per layer: [0] ss1, [1] ss2, [2 .. 2+rows) t2'   (K3 leaves ss1, ss2 at zero for the next call)
// ------------------------------------------------------------------------------------------------
// block = 32 columns x 8 row slices (coalesced 128-B row reads, 8 independent partial sums per column)
__global__ void __launch_bounds__(1024) sn_wtu_kernel(const P2iSnLayer* __restrict__ table) {
    // block = 32 columns x 32 row lanes: R / 32 <= 8 loads per thread, all issued together (one DRAM round trip; with 8 row
    // lanes the 32 dependent-latency iterations made this 20-us kernel pure long-scoreboard stall)
    const P2iSnLayer L = table[blockIdx.y];
    __shared__ float part[32][33];
    const int R = L.rows, K = L.cols;
    if (blockIdx.x * 32 >= K) return;
    const int jl = threadIdx.x & 31, rs = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + jl;
    float s = 0.f;
    if (j < K) {
#pragma unroll 8
        for (int i = rs; i < R; i += 32) s = fmaf(__ldg(L.W + static_cast<size_t>(i) * K + j), __ldg(L.u + i), s);
    }
    part[rs][jl] = s;
    __syncthreads();
    if (rs == 0) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 32; ++q) t += part[q][jl];
        if (j < K) L.v[j] = t;                          // un-normalised; K3 divides by n1
        const float ss = warp_sum(j < K ? t * t : 0.f);
        if (jl == 0) atomicAdd(&L.scratch[0], ss);
    }
}

// block = 2 rows x 128 threads
__global__ void __launch_bounds__(256) sn_wv_kernel(const P2iSnLayer* __restrict__ table) {
    const P2iSnLayer L = table[blockIdx.y];
    __shared__ float part[8];
    const int R = L.rows, K = L.cols;
    if (blockIdx.x * 2 >= R) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 2 + (threadIdx.x >> 7), t = threadIdx.x & 127;
    float s = 0.f;
    if (i < R) {
#pragma unroll 4
        for (int j = t; j < K; j += 128) s = fmaf(__ldg(L.W + static_cast<size_t>(i) * K + j), L.v[j], s);
    }
    s = warp_sum(s);
    if (lane == 0) part[warp] = s;
    __syncthreads();
    if (t == 0 && i < R) {
        const int w0 = (threadIdx.x >> 7) * 4;
        const float r = part[w0] + part[w0 + 1] + part[w0 + 2] + part[w0 + 3];
        L.scratch[2 + i] = r;
        atomicAdd(&L.scratch[1], r * r);
    }
}

__global__ void __launch_bounds__(256) sn_finish_kernel(const P2iSnLayer* __restrict__ table, int training) {
    const P2iSnLayer L = table[blockIdx.x];
    __shared__ float sh[32];
    const int R = L.rows, K = L.cols;
    float sigma;
    if (training) {
        const float n1 = fmaxf(sqrtf(L.scratch[0]), 1e-12f);
        const float n2 = fmaxf(sqrtf(L.scratch[1]) / n1, 1e-12f);
        for (int j = threadIdx.x; j < K; j += blockDim.x) {
            const float vn = L.v[j] / n1;
            L.v[j] = vn;
            if (L.v_snap) L.v_snap[j] = vn;
        }
        float dot = 0.f;
        for (int i = threadIdx.x; i < R; i += blockDim.x) {
            const float t2 = L.scratch[2 + i] / n1;
            const float un = t2 / n2;
            L.u[i] = un;
            if (L.u_snap) L.u_snap[i] = un;
            dot += un * t2;
        }
        sigma = blk_sum(dot, sh);
    } else {
        float dot = 0.f;
        for (int i = threadIdx.x; i < R; i += blockDim.x) {
            dot += L.u[i] * L.scratch[2 + i];
            if (L.u_snap) L.u_snap[i] = L.u[i];
        }
        if (L.v_snap)
            for (int j = threadIdx.x; j < K; j += blockDim.x) L.v_snap[j] = L.v[j];
        sigma = blk_sum(dot, sh);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *L.sigma = sigma;
        L.scratch[0] = 0.f;
        L.scratch[1] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------------
// weight_orig / sigma -> bf16 GEMM operands.  W_orig [Cout][Cin][KT][3][3] (or [Cout][Cin][1][1][1]).
//   s2 == 0: tap' = (kt*k + ky)*k + kx, ci' = ci                      (Cin' = cin_pad >= Cin)
//   s2 == 1: spatial stride 2 -> k=2 conv on the space-to-depth input: ky -> (dy, py) = (0,1),(1,0),(1,1)
//            tap' = (kt*2 + dy)*2 + dx, ci' = (py*2 + px)*Cin + ci      (Cin' = 4*Cin)
//   out   [tap'][Cout][Cin']            forward operand
//   out_t [tapT][Cin'][Cout]            data-gradient operand, spatial taps flipped; temporal taps flipped unless
//                                       keep_t (the temporally transposed mode indexes kt directly)
// Buffers are zero-initialised once by the caller; unused slots are never written.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_taps(const P2iPackLayer& L, int r, int ci, int& tap, int& tapT, int& cip, int& cinp) {
    const int k = L.ksize, kk = k * k;
    const int kt = r / kk, ky = (r - kt * kk) / k, kx = r - kt * kk - ky * k;
    if (L.s2) {
        const int dy = ky == 0 ? 0 : 1, py = ky == 1 ? 0 : 1, dx = kx == 0 ? 0 : 1, px = kx == 1 ? 0 : 1;
        tap = (kt * 2 + dy) * 2 + dx;
        tapT = ((L.keep_t ? kt : L.KT - 1 - kt) * 2 + (1 - dy)) * 2 + (1 - dx);
        cip = (py * 2 + px) * L.Cin + ci;
        cinp = 4 * L.Cin;
    } else {
        tap = (kt * k + ky) * k + kx;
        tapT = ((L.keep_t ? kt : L.KT - 1 - kt) * k + (k - 1 - ky)) * k + (k - 1 - kx);
        cip = ci;
        cinp = L.cin_pad;
    }
}

// Two passes over the (L2-resident) fp32 weights, each ordered so that the WRITES of a warp are contiguous:
// pass 1 runs input channels fastest (forward operand), pass 2 output channels fastest (data-gradient operand).
// 32-bit index arithmetic (a layer has < 2^31 weights).
__global__ void __launch_bounds__(256) disc_pack_weight_kernel(const P2iPackLayer* __restrict__ table) {
    const P2iPackLayer L = table[blockIdx.y];
    const int per = L.KT * L.ksize * L.ksize;
    const int total = L.Cout * L.Cin * per;
    const float inv = L.sigma ? 1.f / *L.sigma : 1.f;
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(L.out);
    __nv_bfloat16* ot = static_cast<__nv_bfloat16*>(L.out_t);
    const int cc = L.Cout * L.Cin;
    if (o)
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
            const int r = e / cc, rem = e - r * cc, co = rem / L.Cin, ci = rem - co * L.Cin;
            int tap, tapT, cip, cinp;
            pack_taps(L, r, ci, tap, tapT, cip, cinp);
            o[(static_cast<size_t>(tap) * L.Cout + co) * cinp + cip] = __float2bfloat16(__ldg(L.W + (co * L.Cin + ci) * per + r) * inv);
        }
    if (ot)
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
            const int r = e / cc, rem = e - r * cc, ci = rem / L.Cout, co = rem - ci * L.Cout;
            int tap, tapT, cip, cinp;
            pack_taps(L, r, ci, tap, tapT, cip, cinp);
            ot[(static_cast<size_t>(tapT) * cinp + cip) * L.Cout + co] = __float2bfloat16(__ldg(L.W + (co * L.Cin + ci) * per + r) * inv);
        }
}

// ------------------------------------------------------------------------------------------------
// x f32 [B,C,H,W] -> bf16 [B,H,W,64] with channels >= C zero (input of the first 2-D layer, C = 16).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) disc_pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int C,
                                                              int HW, long long npix) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // (pixel, 8-channel group)
    if (t >= npix * 8) return;
    const int cg = static_cast<int>(t & 7);
    const long long p = t >> 3;
    const long long b = p / HW, pix = p - b * HW;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = cg * 8 + i;
        f[i] = (c < C) ? x[(b * C + c) * HW + pix] : 0.f;
    }
    reinterpret_cast<uint4*>(y)[t] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                                pack_bf16x2(f[6], f[7]));
}

// ------------------------------------------------------------------------------------------------
// d3d.0: Conv3d(1 -> 32, k 3x3x3, stride (1,2,2), pad 1) + bias + LeakyReLU, written in space-to-depth layout
// [B, T, H/4, W/4, 4*32] for the next (stride-2) layer.  x f32 [B,T,H,W]; w f32 [32][27] (already / sigma).
// one thread per output pixel (t, y, x) of the H/2 x W/2 grid.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) d3d_first_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ sigma, const float* __restrict__ bias,
                                                            __nv_bfloat16* __restrict__ y, int B, int T, int H, int W) {
    __shared__ float sw[32 * 27], sb[32];
    const float inv = 1.f / *sigma;
    for (int i = threadIdx.x; i < 32 * 27; i += blockDim.x) sw[i] = w[i] * inv;
    if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int Ho = H >> 1, Wo = W >> 1;
    const int total = B * T * Ho * Wo;          // < 2^31 (checked on the host)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int xo = idx % Wo, r1 = idx / Wo, yo = r1 % Ho, r2 = r1 / Ho, t = r2 % T, b = r2 / T;
    float in[27];
#pragma unroll
    for (int kt = 0; kt < 3; ++kt)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ti = t + kt - 1, yi = 2 * yo + ky - 1, xi = 2 * xo + kx - 1;
                in[(kt * 3 + ky) * 3 + kx] = (ti >= 0 && ti < T && yi >= 0 && yi < H && xi >= 0 && xi < W)
                                                 ? __ldg(x + ((static_cast<size_t>(b) * T + ti) * H + yi) * W + xi) : 0.f;
            }
    // s2d address: pixel (yo, xo), channel c -> [yo/2][xo/2][(yo&1)*2 + (xo&1)][c]
    __nv_bfloat16* o = y + ((((static_cast<size_t>(b) * T + t) * (Ho >> 1) + (yo >> 1)) * (Wo >> 1) + (xo >> 1)) * 4 + ((yo & 1) * 2 + (xo & 1))) * 32;
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 8) {
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float s = sb[c0 + i];
            const float* wp = sw + (c0 + i) * 27;
#pragma unroll
            for (int k = 0; k < 27; ++k) s = fmaf(wp[k], in[k], s);
            a[i] = s > 0.f ? s : 0.2f * s;
        }
        *reinterpret_cast<uint4*>(o + c0) = make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]),
                                                       pack_bf16x2(a[6], a[7]));
    }
}


// Register-blocked variant (W % 8 == 0): thread = 4 consecutive output pixels x 8 output channels.  The 9 input
// values of a (frame, row) feed 3 horizontal taps x 4 pixels; the 8 weights of a tap come as two broadcast LDS.128
// ([tap][channel] layout), so one shared-memory access feeds 16 FMAs instead of 1.
__global__ void __launch_bounds__(128) d3d_first_fwd4_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ sigma, const float* __restrict__ bias,
                                                             __nv_bfloat16* __restrict__ y, int B, int T, int H, int W) {
    __shared__ __align__(16) float sw[27 * 32];
    __shared__ float sb[32];
    const float inv = 1.f / *sigma;
    for (int i = threadIdx.x; i < 32 * 27; i += blockDim.x) {
        const int c = i / 27, tap = i - c * 27;
        sw[tap * 32 + c] = w[i] * inv;
    }
    if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int Ho = H >> 1, Wo = W >> 1, Wg = Wo >> 2;
    const int ngroups = B * T * Ho * Wg;
    const int gid = blockIdx.x * 32 + (threadIdx.x >> 2);
    const int cq = threadIdx.x & 3;
    if (gid >= ngroups) return;
    const int j = gid % Wg, r1 = gid / Wg, yo = r1 % Ho, r2 = r1 / Ho, t = r2 % T, b = r2 / T;
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[p][c] = sb[cq * 8 + c];
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
        const int ti = t + kt - 1;
        if (ti < 0 || ti >= T) continue;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yi = 2 * yo + ky - 1;
            if (yi < 0 || yi >= H) continue;
            const float* row = x + ((static_cast<size_t>(b) * T + ti) * H + yi) * W + 8 * j;
            float v[9];
            v[0] = (j > 0) ? __ldg(row - 1) : 0.f;
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(row)), q1 = __ldg(reinterpret_cast<const float4*>(row) + 1);
            v[1] = q0.x; v[2] = q0.y; v[3] = q0.z; v[4] = q0.w; v[5] = q1.x; v[6] = q1.y; v[7] = q1.z; v[8] = q1.w;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4* wp = reinterpret_cast<const float4*>(sw + ((kt * 3 + ky) * 3 + kx) * 32 + cq * 8);
                const float4 w0 = wp[0], w1 = wp[1];
                const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[p][c] = fmaf(v[2 * p + kx], ww[c], acc[p][c]);
            }
        }
    }
    // s2d address: pixel (yo, xo), channel c -> [yo/2][xo/2][(yo&1)*2 + (xo&1)][c]
    __nv_bfloat16* cell = y + (((static_cast<size_t>(b) * T + t) * (Ho >> 1) + (yo >> 1)) * (Wo >> 1) + 2 * j) * 128 + (yo & 1) * 64 + cq * 8;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        float a[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) a[c] = acc[p][c] > 0.f ? acc[p][c] : 0.2f * acc[p][c];
        *reinterpret_cast<uint4*>(cell + (p >> 1) * 128 + (p & 1) * 32) =
            make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
    }
}


// (A warp-MMA variant -- im2col fragments gathered from a staged patch, 8 mma.sync.m16n8k16 per 16 pixels -- was built and
// measured at 108 us against 78 us for the kernel above: 618 instructions per 16-pixel tile and only 20 resident warps
// left it bound by the global-load latency of the patch; it was dropped.)

// ------------------------------------------------------------------------------------------------
// d2d.8: Conv2d(256 -> 1, 3x3, pad 1) + bias, no activation.  y bf16 [B,H,W,C]; w f32 [C][9] (weight_orig),
// out f32 [B,H,W].  One warp per output pixel, lanes over channels (8 per lane per 256).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) d2d_last_fwd_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ w,
                                                           const float* __restrict__ sigma, const float* __restrict__ bias,
                                                           float* __restrict__ out, int B, int H, int W, int C) {
    extern __shared__ float sw[];   // [9][C]
    const float inv = 1.f / *sigma;
    for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
        const int tap = i / C, c = i - tap * C;
        sw[i] = w[c * 9 + tap] * inv;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long pix = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (pix >= static_cast<long long>(B) * H * W) return;
    const int xx = static_cast<int>(pix % W), yy = static_cast<int>((pix / W) % H), b = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float acc = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = yy + ky - 1;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = xx + kx - 1;
            if (ix < 0 || ix >= W) continue;
            const __nv_bfloat16* yp = y + ((static_cast<size_t>(b) * H + iy) * W + ix) * C;
            const float* wp = sw + (ky * 3 + kx) * C;
            for (int c = lane * 8; c < C; c += 256) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(yp + c));
                const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = unpack_bf16x2(qq[i]);
                    acc = fmaf(f.x, wp[c + 2 * i], acc);
                    acc = fmaf(f.y, wp[c + 2 * i + 1], acc);
                }
            }
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[pix] = acc + bias[0];
}


// C == 256: lane l owns channels 8l..8l+7 and keeps their 9 x 8 weights in registers; warps stride over pixels
// (persistent grid), so the strided weight gather happens once per warp instead of once per 8 pixels.
__global__ void __launch_bounds__(256) d2d_last_fwd256_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ w,
                                                              const float* __restrict__ sigma, const float* __restrict__ bias,
                                                              float* __restrict__ out, int B, int H, int W) {
    constexpr int C = 256;
    const int lane = threadIdx.x & 31;
    const float inv = 1.f / *sigma;
    float wr[9][8];
    {
        const float* wp = w + lane * 72;          // w[c][tap], c = 8*lane + i  ->  72 contiguous floats
        float tmp[72];
#pragma unroll
        for (int q = 0; q < 18; ++q) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(wp) + q);
            tmp[4 * q] = f.x; tmp[4 * q + 1] = f.y; tmp[4 * q + 2] = f.z; tmp[4 * q + 3] = f.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int t = 0; t < 9; ++t) wr[t][i] = tmp[i * 9 + t] * inv;
    }
    const float b0 = bias[0];
    const int npix = B * H * W;
    const int nwarps = gridDim.x * 8;
    for (int pix = blockIdx.x * 8 + (threadIdx.x >> 5); pix < npix; pix += nwarps) {
        const int xx = pix % W, yy = (pix / W) % H;
        float acc = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = yy + ky - 1;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = xx + kx - 1;
                if (ix < 0 || ix >= W) continue;
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(y + static_cast<size_t>(pix + (ky - 1) * W + (kx - 1)) * C) + lane);
                const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = unpack_bf16x2(qq[i]);
                    acc = fmaf(f.x, wr[ky * 3 + kx][2 * i], acc);
                    acc = fmaf(f.y, wr[ky * 3 + kx][2 * i + 1], acc);
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) out[pix] = acc + b0;
    }
}

// ------------------------------------------------------------------------------------------------
// Tail (p2igan.py:164-173):  o3[b,t,y,x] = w3 . z[b,t,y,x,:] + b3 ;  m = mean_t o3 ;
//   fused = sigmoid(alpha2d) * out2d + bilinear(m -> out2d grid, align_corners=False)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) disc_tail_mean_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ w3,
                                                             const float* __restrict__ sigma, const float* __restrict__ b3,
                                                             float* __restrict__ m, int B, int T, int HW, int C) {
    const int lane = threadIdx.x & 31;
    const long long pix = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);   // (b, y, x)
    if (pix >= static_cast<long long>(B) * HW) return;
    const int b = static_cast<int>(pix / HW), p = static_cast<int>(pix - static_cast<long long>(b) * HW);
    const float inv = 1.f / *sigma;
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
        const __nv_bfloat16* zp = z + ((static_cast<size_t>(b) * T + t) * HW + p) * C;
        for (int c = lane * 2; c < C; c += 64) {
            const float2 f = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(zp + c)));
            acc = fmaf(f.x, w3[c], acc);
            acc = fmaf(f.y, w3[c + 1], acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) m[pix] = acc * inv / static_cast<float>(T) + b3[0];
}

// PyTorch bilinear, align_corners=False: src = (dst + 0.5) * (in/out) - 0.5, clamped at 0; idx1 = min(idx0+1, in-1)
__device__ __forceinline__ void bil_src(int dst, float scale, int in, int& i0, int& i1, float& l1) {
    float s = (dst + 0.5f) * scale - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = static_cast<int>(s);
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = s - i0;
}

__global__ void __launch_bounds__(256) disc_tail_fuse_kernel(const float* __restrict__ m, const float* __restrict__ out2d,
                                                             const float* __restrict__ alpha, float* __restrict__ fused, int B, int h,
                                                             int w, int H2, int W2) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(B) * H2 * W2) return;
    const int X = static_cast<int>(idx % W2), Y = static_cast<int>((idx / W2) % H2), b = static_cast<int>(idx / (static_cast<long long>(W2) * H2));
    float up;
    if (h == H2 && w == W2) {
        up = m[idx];
    } else {
        int y0, y1, x0, x1;
        float ly, lx;
        bil_src(Y, static_cast<float>(h) / H2, h, y0, y1, ly);
        bil_src(X, static_cast<float>(w) / W2, w, x0, x1, lx);
        const float* mb = m + static_cast<size_t>(b) * h * w;
        up = (1.f - ly) * ((1.f - lx) * mb[y0 * w + x0] + lx * mb[y0 * w + x1]) + ly * ((1.f - lx) * mb[y1 * w + x0] + lx * mb[y1 * w + x1]);
    }
    const float s = 1.f / (1.f + expf(-alpha[0]));
    fused[idx] = s * out2d[idx] + up;
}

}  // namespace p2i

using namespace p2i;

extern "C" int p2i_spectral_norm(const P2iSnLayer* table_dev, int n_layers, int max_rows, int max_cols, int training,
                                 void* stream) {
    P2I_CHECK_ARG(table_dev && n_layers > 0 && max_rows > 0 && max_cols > 0, "spectral_norm: bad arguments");
    if (training) {
        sn_wtu_kernel<<<dim3(cdiv(max_cols, 32), n_layers), 1024, 0, as_stream(stream)>>>(table_dev);
        P2I_CHECK_LAUNCH("sn_wtu_kernel");
    }
    sn_wv_kernel<<<dim3(cdiv(max_rows, 2), n_layers), 256, 0, as_stream(stream)>>>(table_dev);
    P2I_CHECK_LAUNCH("sn_wv_kernel");
    sn_finish_kernel<<<n_layers, 256, 0, as_stream(stream)>>>(table_dev, training);
    P2I_CHECK_LAUNCH("sn_finish_kernel");
    return P2I_OK;
}

extern "C" int p2i_disc_pack_weights(const P2iPackLayer* table_dev, int n_layers, void* stream) {
    P2I_CHECK_ARG(table_dev && n_layers > 0, "disc_pack_weights: empty table");
    dim3 grid(128, n_layers);
    disc_pack_weight_kernel<<<grid, 256, 0, as_stream(stream)>>>(table_dev);
    P2I_CHECK_LAUNCH("disc_pack_weight_kernel");
    return P2I_OK;
}

extern "C" int p2i_disc_pack_input(const float* x, void* y, int B, int C, int H, int W, void* stream) {
    P2I_CHECK_ARG(x && y && C > 0 && C <= 64, "disc_pack_input: bad arguments");
    const long long npix = static_cast<long long>(B) * H * W;
    disc_pack_input_kernel<<<static_cast<unsigned>((npix * 8 + 255) / 256), 256, 0, as_stream(stream)>>>(
        x, static_cast<__nv_bfloat16*>(y), C, H * W, npix);
    P2I_CHECK_LAUNCH("disc_pack_input_kernel");
    return P2I_OK;
}

extern "C" int p2i_d3d_first_fwd(const float* x, const float* w, const float* sigma, const float* bias, void* y, int B, int T,
                                 int H, int W, void* stream) {
    P2I_CHECK_ARG(x && w && sigma && bias && y, "d3d_first_fwd: null pointer");
    P2I_CHECK_ARG(H % 4 == 0 && W % 4 == 0, "d3d_first_fwd: H, W must be multiples of 4");
    P2I_CHECK_ARG(static_cast<long long>(B) * T * H * W < (1ll << 31), "d3d_first_fwd: tensor too large for 32-bit indexing");
    const long long total = static_cast<long long>(B) * T * (H / 2) * (W / 2);
    if (d3d_first_mma_ok(T, H, W)) return d3d_first_fwd_mma(x, w, sigma, bias, y, B, T, H, W, as_stream(stream));
    if (W % 8 == 0) {
        d3d_first_fwd4_kernel<<<static_cast<unsigned>((total / 4 + 31) / 32), 128, 0, as_stream(stream)>>>(
            x, w, sigma, bias, static_cast<__nv_bfloat16*>(y), B, T, H, W);
        P2I_CHECK_LAUNCH("d3d_first_fwd4_kernel");
        return P2I_OK;
    }
    d3d_first_fwd_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, as_stream(stream)>>>(
        x, w, sigma, bias, static_cast<__nv_bfloat16*>(y), B, T, H, W);
    P2I_CHECK_LAUNCH("d3d_first_fwd_kernel");
    return P2I_OK;
}

extern "C" int p2i_d2d_last_fwd(const void* y, const float* w, const float* sigma, const float* bias, float* out, int B, int H,
                                int W, int C, void* stream) {
    P2I_CHECK_ARG(y && w && sigma && bias && out && C % 8 == 0, "d2d_last_fwd: bad arguments");
    const long long npix = static_cast<long long>(B) * H * W;
    if (C == 256 && npix < (1ll << 31)) {
        long long blocks = (npix + 7) / 8;
        if (blocks > sm_count() * 2) blocks = sm_count() * 2;
        d2d_last_fwd256_kernel<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(y), w, sigma,
                                                                                            bias, out, B, H, W);
        P2I_CHECK_LAUNCH("d2d_last_fwd256_kernel");
        return P2I_OK;
    }
    d2d_last_fwd_kernel<<<static_cast<unsigned>((npix + 7) / 8), 256, 9 * C * sizeof(float), as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(y), w, sigma, bias, out, B, H, W, C);
    P2I_CHECK_LAUNCH("d2d_last_fwd_kernel");
    return P2I_OK;
}

extern "C" int p2i_disc_tail_fwd(const void* z, const float* w3, const float* sigma3, const float* b3, const float* out2d,
                                 const float* alpha, float* m_scratch, float* fused, int B, int T, int h, int w, int C, int H2,
                                 int W2, void* stream) {
    P2I_CHECK_ARG(z && w3 && sigma3 && b3 && out2d && alpha && m_scratch && fused, "disc_tail_fwd: null pointer");
    const long long npix = static_cast<long long>(B) * h * w;
    disc_tail_mean_kernel<<<static_cast<unsigned>((npix + 7) / 8), 256, 0, as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(z), w3, sigma3, b3, m_scratch, B, T, h * w, C);
    P2I_CHECK_LAUNCH("disc_tail_mean_kernel");
    const long long n = static_cast<long long>(B) * H2 * W2;
    disc_tail_fuse_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream)>>>(m_scratch, out2d, alpha, fused, B, h,
                                                                                                  w, H2, W2);
    P2I_CHECK_LAUNCH("disc_tail_fuse_kernel");
    return P2I_OK;
}
