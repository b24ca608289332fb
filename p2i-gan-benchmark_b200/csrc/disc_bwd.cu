// Backward of the CUDA-core pieces of the discriminator (see disc.cu): fused tail, thin first/last layers,
// bias reductions, and the spectral-norm backward that also un-packs the tensor-core weight gradients.
#include "common.h"
#include "ptx.cuh"

namespace p2i {

__device__ __forceinline__ float blk_sum_all(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float t = (lane < nw) ? sh[lane] : 0.f;
    return warp_sum(t);
}

__device__ __forceinline__ void bil_src_b(int dst, float scale, int in, int& i0, int& i1, float& l1) {
    float s = (dst + 0.5f) * scale - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = static_cast<int>(s);
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = s - i0;
}

// ------------------------------------------------------------------------------------------------
// Tail backward, part 1: fused = s*out2d + up(m), s = sigmoid(alpha)
//   d_out2d = s * dfused ;  dalpha += s(1-s) * sum dfused*out2d
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tail_bwd_fuse_kernel(const float* __restrict__ dfused, const float* __restrict__ out2d,
                                                            const float* __restrict__ alpha, float* __restrict__ d_out2d,
                                                            float* __restrict__ dalpha, long long n) {
    __shared__ float sh[32];
    const float s = 1.f / (1.f + expf(-alpha[0]));
    float acc = 0.f;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float g = dfused[i];
        d_out2d[i] = s * g;
        acc += g * out2d[i];
    }
    acc = blk_sum_all(acc, sh);
    if (threadIdx.x == 0 && dalpha) atomicAdd(dalpha, acc * s * (1.f - s));
}

// part 2: one warp per low-res pixel (b, y, x): d_m = transpose-bilinear(dfused); for every frame t:
//   d_o3 = d_m / T ;  dpre[b,t,y,x,c] = lmask(z) * d_o3 * w3[c]/sigma ;  dW3[c] += d_o3 * z ;  db3 += d_o3
__global__ void __launch_bounds__(256) tail_bwd_mean_kernel(const float* __restrict__ dfused, const __nv_bfloat16* __restrict__ z,
                                                            const float* __restrict__ w3, const float* __restrict__ sigma,
                                                            __nv_bfloat16* __restrict__ dpre, float* __restrict__ dW3, float* __restrict__ db3,
                                                            int B, int T, int h, int w, int C, int H2, int W2) {
    extern __shared__ float s_dw[];   // [C]
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_dw[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long pix = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const bool live = pix < static_cast<long long>(B) * h * w;
    if (live) {
        const int x = static_cast<int>(pix % w), y = static_cast<int>((pix / w) % h), b = static_cast<int>(pix / (static_cast<long long>(w) * h));
        float dm = 0.f;
        const float* gb = dfused + static_cast<size_t>(b) * H2 * W2;
        if (h == H2 && w == W2) {
            dm = gb[y * w + x];
        } else {
            const float sy = static_cast<float>(h) / H2, sx = static_cast<float>(w) / W2;
            const int Ya = max(0, 2 * y - 3), Yb = min(H2 - 1, 2 * y + 4), Xa = max(0, 2 * x - 3), Xb = min(W2 - 1, 2 * x + 4);
            for (int Y = Ya; Y <= Yb; ++Y) {
                int y0, y1; float ly;
                bil_src_b(Y, sy, h, y0, y1, ly);
                const float wy = (y0 == y ? 1.f - ly : 0.f) + (y1 == y ? ly : 0.f);
                if (wy == 0.f) continue;
                for (int X = Xa + lane; X <= Xb; X += 32) {
                    int x0, x1; float lx;
                    bil_src_b(X, sx, w, x0, x1, lx);
                    const float wx = (x0 == x ? 1.f - lx : 0.f) + (x1 == x ? lx : 0.f);
                    dm += wy * wx * gb[Y * W2 + X];
                }
            }
            dm = warp_sum(dm);
        }
        const float do3 = dm / static_cast<float>(T);
        const float inv = 1.f / *sigma;
        if (lane == 0 && db3) atomicAdd(db3, dm);
        for (int t = 0; t < T; ++t) {
            const size_t o = ((static_cast<size_t>(b) * T + t) * h * w + static_cast<size_t>(y) * w + x) * C;
            for (int c = lane * 2; c < C; c += 64) {
                const float2 f = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(z + o + c)));
                const float g0 = do3 * w3[c] * inv * (f.x > 0.f ? 1.f : 0.2f);
                const float g1 = do3 * w3[c + 1] * inv * (f.y > 0.f ? 1.f : 0.2f);
                *reinterpret_cast<uint32_t*>(dpre + o + c) = pack_bf16x2(g0, g1);
                if (dW3) { atomicAdd(&s_dw[c], do3 * f.x); atomicAdd(&s_dw[c + 1], do3 * f.y); }
            }
        }
    }
    __syncthreads();
    if (dW3)
        for (int i = threadIdx.x; i < C; i += blockDim.x)
            if (s_dw[i] != 0.f) atomicAdd(&dW3[i], s_dw[i]);
}

// ------------------------------------------------------------------------------------------------
// d2d.8 backward: dpre[b,y,x,c] = lmask(y4) * sum_taps d_o[y-ky+1, x-kx+1] * w[c][ky,kx]/sigma   (warp per pixel)
//                 dW[c][tap] += sum_pix d_o[pix] * y4[pix+tap][c] ;  db += sum d_o
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) d2d_last_bwd_kernel(const float* __restrict__ d_o, const __nv_bfloat16* __restrict__ y4,
                                                           const float* __restrict__ w, const float* __restrict__ sigma,
                                                           __nv_bfloat16* __restrict__ dpre, float* __restrict__ dW, float* __restrict__ db,
                                                           int B, int H, int W, int C) {
    extern __shared__ float sm[];            // sw[9*C] then sdw[9*C]
    float* sw = sm;
    float* sdw = sm + 9 * C;
    const float inv = 1.f / *sigma;
    for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
        const int tap = i / C, c = i - tap * C;
        sw[i] = w[c * 9 + tap] * inv;
        sdw[i] = 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long pix = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (pix < static_cast<long long>(B) * H * W) {
        const int xx = static_cast<int>(pix % W), yy = static_cast<int>((pix / W) % H), b = static_cast<int>(pix / (static_cast<long long>(W) * H));
        const float* gb = d_o + static_cast<size_t>(b) * H * W;
        float g[9];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int oy = yy - ky + 1, ox = xx - kx + 1;
                g[ky * 3 + kx] = (oy >= 0 && oy < H && ox >= 0 && ox < W) ? gb[oy * W + ox] : 0.f;
            }
        const float g0 = gb[yy * W + xx];
        if (lane == 0 && db) atomicAdd(db, g0);
        const size_t o = static_cast<size_t>(pix) * C;
        for (int c = lane * 2; c < C; c += 64) {
            const float2 f = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(y4 + o + c)));
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t) { a0 = fmaf(g[t], sw[t * C + c], a0); a1 = fmaf(g[t], sw[t * C + c + 1], a1); }
            *reinterpret_cast<uint32_t*>(dpre + o + c) = pack_bf16x2(a0 * (f.x > 0.f ? 1.f : 0.2f), a1 * (f.y > 0.f ? 1.f : 0.2f));
        }
        if (dW) {   // this pixel's y4 feeds output pixels (yy-ky+1, xx-kx+1) through tap (ky,kx)
            for (int c = lane * 2; c < C; c += 64) {
                const float2 f = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(y4 + o + c)));
#pragma unroll
                for (int t = 0; t < 9; ++t)
                    if (g[t] != 0.f) { atomicAdd(&sdw[t * C + c], g[t] * f.x); atomicAdd(&sdw[t * C + c + 1], g[t] * f.y); }
            }
        }
    }
    __syncthreads();
    if (dW)
        for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
            const int tap = i / C, c = i - tap * C;
            if (sdw[i] != 0.f) atomicAdd(&dW[c * 9 + tap], sdw[i]);
        }
}

// ------------------------------------------------------------------------------------------------
// column sums of a bf16 [rows, C] gradient -> f32 [C]  (bias gradients)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ g, float* __restrict__ out, long long rows, int C) {
    extern __shared__ float s_acc[];   // [C]
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const int cg = C >> 3;
    const long long total = rows * cg;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // stride is a multiple of cg so each thread keeps the same channel group
    const long long stride = (static_cast<long long>(gridDim.x) * blockDim.x / cg) * cg;
    long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int c8 = static_cast<int>(i % cg);
    const bool live = i < stride;
    {
        for (; live && i < total; i += stride) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(g) + i);
            const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = unpack_bf16x2(qq[k]);
                acc[2 * k] += f.x;
                acc[2 * k + 1] += f.y;
            }
        }
    }
    // lanes l, l+cg, l+2cg.. of a warp hold the same channel group when cg < 32 (the stride is a multiple of cg and
    // blockDim of 32): fold them with shuffles so that at most one lane per (warp, channel) touches shared memory
    const bool fold = (cg & (cg - 1)) == 0 && cg < 32;      // power-of-two group count: lane ^ off keeps the group
    if (fold)
        for (int off = 16; off >= cg; off >>= 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
        }
    if (!fold || (threadIdx.x & 31) < cg) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(&s_acc[c8 * 8 + k], acc[k]);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < C; k += blockDim.x) atomicAdd(&out[k], s_acc[k]);
}

// ------------------------------------------------------------------------------------------------
// d3d.0 backward.  dpre bf16 [B,T,H/2,W/2,32] (natural layout), x f32 [B,T,H,W].
//   dW[c][tap] += sum dpre * x_patch ; db[c] += sum dpre                     (one thread per output pixel)
//   dx[b,t,y,x]  = sum_c sum_taps dpre[b, t-kt+1, (y-ky+1)/2, (x-kx+1)/2, c] * w[c][tap]/sigma   (gather, parity)
// ------------------------------------------------------------------------------------------------
// thread = (pixel sub-stream s in 0..15, channel octet cq in 0..3, temporal tap kt in 0..2): 8 channels x 9 spatial
// taps accumulated in registers over the thread's pixels, one shared-memory + one global reduction per block.
__global__ void __launch_bounds__(192) d3d_first_bwd_w_kernel(const __nv_bfloat16* __restrict__ dpre, const float* __restrict__ x,
                                                              float* __restrict__ dW, float* __restrict__ db, int B, int T, int H, int W) {
    __shared__ float sdw[32 * 27], sdb[32];
    for (int i = threadIdx.x; i < 32 * 27; i += blockDim.x) sdw[i] = 0.f;
    if (threadIdx.x < 32) sdb[threadIdx.x] = 0.f;
    __syncthreads();
    const int kt = threadIdx.x % 3, cq = (threadIdx.x / 3) & 3, s = threadIdx.x / 12;
    const int Ho = H >> 1, Wo = W >> 1;
    const int total = B * T * Ho * Wo;          // < 2^31 (checked on the host)
    float acc[8][9];
    float bsum[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        bsum[c] = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[c][k] = 0.f;
    }
    for (int idx = blockIdx.x * 16 + s; idx < total; idx += gridDim.x * 16) {
        const int xo = idx % Wo, r1 = idx / Wo, yo = r1 % Ho, r2 = r1 / Ho, t = r2 % T, b = r2 / T;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(dpre + static_cast<size_t>(idx) * 32 + cq * 8));
        const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
        float g[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = unpack_bf16x2(qq[k]);
            g[2 * k] = f.x; g[2 * k + 1] = f.y;
        }
        if (kt == 0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) bsum[c] += g[c];
        }
        const int ti = t + kt - 1;
        if (ti < 0 || ti >= T) continue;
        const float* xp = x + (static_cast<size_t>(b) * T + ti) * H * W;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int yi = 2 * yo + ky - 1, xi = 2 * xo + kx - 1;
                const float v = (yi >= 0 && yi < H && xi >= 0 && xi < W) ? __ldg(xp + static_cast<size_t>(yi) * W + xi) : 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c][ky * 3 + kx] = fmaf(g[c], v, acc[c][ky * 3 + kx]);
            }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int k = 0; k < 9; ++k) atomicAdd(&sdw[(cq * 8 + c) * 27 + kt * 9 + k], acc[c][k]);
        if (kt == 0) atomicAdd(&sdb[cq * 8 + c], bsum[c]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 27; i += blockDim.x)
        if (sdw[i] != 0.f) atomicAdd(&dW[i], sdw[i]);
    if (threadIdx.x < 32 && sdb[threadIdx.x] != 0.f) atomicAdd(&db[threadIdx.x], sdb[threadIdx.x]);
}

__global__ void __launch_bounds__(128) d3d_first_bwd_x_kernel(const __nv_bfloat16* __restrict__ dpre, const float* __restrict__ w,
                                                              const float* __restrict__ sigma, float* __restrict__ dx, int B, int T, int H, int W) {
    __shared__ float sw[32 * 27];
    const float inv = 1.f / *sigma;
    for (int i = threadIdx.x; i < 32 * 27; i += blockDim.x) sw[i] = w[i] * inv;
    __syncthreads();
    const int Ho = H >> 1, Wo = W >> 1;
    const int total = B * T * H * W;            // < 2^31 (checked on the host)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int xx = idx % W, r1 = idx / W, yy = r1 % H, r2 = r1 / H, t = r2 % T, b = r2 / T;
    float acc = 0.f;
    for (int kt = 0; kt < 3; ++kt) {
        const int to = t - kt + 1;
        if (to < 0 || to >= T) continue;
        for (int ky = 0; ky < 3; ++ky) {
            const int sy = yy - ky + 1;
            if (sy < 0 || (sy & 1) || (sy >> 1) >= Ho) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int sx = xx - kx + 1;
                if (sx < 0 || (sx & 1) || (sx >> 1) >= Wo) continue;
                const uint4* gp = reinterpret_cast<const uint4*>(dpre + (((static_cast<size_t>(b) * T + to) * Ho + (sy >> 1)) * Wo + (sx >> 1)) * 32);
                const int tap = (kt * 3 + ky) * 3 + kx;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const uint4 q = __ldg(gp + q4);
                    const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 f = unpack_bf16x2(qq[k]);
                        acc = fmaf(f.x, sw[(q4 * 8 + 2 * k) * 27 + tap], acc);
                        acc = fmaf(f.y, sw[(q4 * 8 + 2 * k + 1) * 27 + tap], acc);
                    }
                }
            }
        }
    }
    dx[idx] = acc;
}


// ------------------------------------------------------------------------------------------------
// Register-blocked variants of the thin-layer backward kernels.  Reductions go warp-shuffle -> one shared-memory
// pass -> one global atomic per (block, element): shared-memory float atomics are CAS loops and serialise badly.
// ------------------------------------------------------------------------------------------------

// d2d.8 backward, C == 256: lane l owns channels 8l..8l+7 (weights and dW accumulators in registers), warps stride
// over pixels.
__global__ void __launch_bounds__(256) d2d_last_bwd256_kernel(const float* __restrict__ d_o, const __nv_bfloat16* __restrict__ y4,
                                                              const float* __restrict__ w, const float* __restrict__ sigma,
                                                              __nv_bfloat16* __restrict__ dpre, float* __restrict__ dW,
                                                              float* __restrict__ db, int B, int H, int W) {
    constexpr int C = 256;
    __shared__ float sdw[9 * C];
    __shared__ float sdb[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inv = 1.f / *sigma;
    float wr[9][8], aw[9][8];
    {
        const float* wp = w + lane * 72;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int t = 0; t < 9; ++t) { wr[t][i] = __ldg(wp + i * 9 + t) * inv; aw[t][i] = 0.f; }
    }
    float bsum = 0.f;
    const int npix = B * H * W;
    const int nwarps = gridDim.x * 8;
    for (int pix = blockIdx.x * 8 + warp; pix < npix; pix += nwarps) {
        const int xx = pix % W, yy = (pix / W) % H;
        float g[9];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int oy = yy - ky + 1, ox = xx - kx + 1;
                g[ky * 3 + kx] = (oy >= 0 && oy < H && ox >= 0 && ox < W) ? __ldg(d_o + pix - (ky - 1) * W - (kx - 1)) : 0.f;
            }
        bsum += g[4];
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(y4 + static_cast<size_t>(pix) * C) + lane);
        const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
        float f[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t2 = unpack_bf16x2(qq[i]);
            f[2 * i] = t2.x; f[2 * i + 1] = t2.y;
        }
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float s = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                s = fmaf(g[t], wr[t][i], s);
                aw[t][i] = fmaf(g[t], f[i], aw[t][i]);
            }
            a[i] = s * (f[i] > 0.f ? 1.f : 0.2f);
        }
        reinterpret_cast<uint4*>(dpre + static_cast<size_t>(pix) * C)[lane] =
            make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
    }
    if (!dW) return;
    // block reduction of the 8 warps: each warp adds its registers into shared memory in turn
    for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) sdw[i] = 0.f;
    if (lane == 0) sdb[warp] = bsum;
    __syncthreads();
    for (int wv = 0; wv < 8; ++wv) {
        if (warp == wv) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int t = 0; t < 9; ++t) sdw[lane * 72 + i * 9 + t] += aw[t][i];
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 9 * C; i += blockDim.x)
        if (sdw[i] != 0.f) atomicAdd(&dW[i], sdw[i]);
    if (threadIdx.x == 0 && db) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += sdb[i];
        atomicAdd(db, s);
    }
}

// d3d.0 weight gradient, W % 8 == 0.  Warp = 8 groups of 4 consecutive output pixels (lane >> 2) x 4 channel octets
// (lane & 3); the warp's temporal tap is warp % 3.  72 accumulators (8 channels x 9 spatial taps) per thread.
__global__ void __launch_bounds__(192) d3d_first_bwd_w4_kernel(const __nv_bfloat16* __restrict__ dpre, const float* __restrict__ x,
                                                               float* __restrict__ dW, float* __restrict__ db, int B, int T, int H, int W) {
    __shared__ float sdw[32 * 27], sdb[32];
    for (int i = threadIdx.x; i < 32 * 27; i += blockDim.x) sdw[i] = 0.f;
    if (threadIdx.x < 32) sdb[threadIdx.x] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kt = warp % 3, half = warp / 3;
    const int cq = lane & 3, gl = lane >> 2;
    const int Ho = H >> 1, Wo = W >> 1, Wg = Wo >> 2;
    const int ngroups = B * T * Ho * Wg;
    float acc[8][9], bsum[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        bsum[c] = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[c][k] = 0.f;
    }
    for (int g0 = (blockIdx.x * 2 + half) * 8; g0 < ngroups; g0 += gridDim.x * 16) {
        const int gid = g0 + gl;
        if (gid >= ngroups) continue;
        const int j = gid % Wg, r1 = gid / Wg, yo = r1 % Ho, r2 = r1 / Ho, t = r2 % T, b = r2 / T;
        const int ti = t + kt - 1;
        const bool tok = ti >= 0 && ti < T;
        if (!tok && kt != 0) continue;
        const uint4* gp = reinterpret_cast<const uint4*>(dpre + static_cast<size_t>(gid) * 128 + cq * 8);   // 4 pixels x 32 ch
        float v[3][9];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yi = 2 * yo + ky - 1;
            if (tok && yi >= 0 && yi < H) {
                const float* row = x + ((static_cast<size_t>(b) * T + ti) * H + yi) * W + 8 * j;
                v[ky][0] = (j > 0) ? __ldg(row - 1) : 0.f;
                const float4 q0 = __ldg(reinterpret_cast<const float4*>(row)), q1 = __ldg(reinterpret_cast<const float4*>(row) + 1);
                v[ky][1] = q0.x; v[ky][2] = q0.y; v[ky][3] = q0.z; v[ky][4] = q0.w;
                v[ky][5] = q1.x; v[ky][6] = q1.y; v[ky][7] = q1.z; v[ky][8] = q1.w;
            } else {
#pragma unroll
                for (int k = 0; k < 9; ++k) v[ky][k] = 0.f;
            }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const uint4 q = __ldg(gp + p * 4);
            const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
            float g[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = unpack_bf16x2(qq[k]);
                g[2 * k] = f.x; g[2 * k + 1] = f.y;
            }
            if (kt == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) bsum[c] += g[c];
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[c][ky * 3 + kx] = fmaf(g[c], v[ky][2 * p + kx], acc[c][ky * 3 + kx]);
        }
    }
    // reduce over the 8 pixel groups of the warp (lane bits 2..4), then one shared-memory add per (warp, element)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            float a = acc[c][k];
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            a += __shfl_xor_sync(0xffffffffu, a, 8);
            a += __shfl_xor_sync(0xffffffffu, a, 16);
            acc[c][k] = a;
        }
        float s = bsum[c];
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        bsum[c] = s;
    }
    for (int h = 0; h < 2; ++h) {               // the two warps that share a temporal tap take turns
        if (half == h && gl == 0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
#pragma unroll
                for (int k = 0; k < 9; ++k) sdw[(cq * 8 + c) * 27 + kt * 9 + k] += acc[c][k];
                if (kt == 0) sdb[cq * 8 + c] += bsum[c];
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 32 * 27; i += blockDim.x)
        if (sdw[i] != 0.f) atomicAdd(&dW[i], sdw[i]);
    if (threadIdx.x < 32 && sdb[threadIdx.x] != 0.f) atomicAdd(&db[threadIdx.x], sdb[threadIdx.x]);
}

// d3d.0 input gradient (transposed stride-2 conv), W % 4 == 0.  A group of 4 lanes = two adjacent 2x2 blocks of dx
// (rows 2m, 2m+1, columns 4j .. 4j+3), fed by the 2 x 3 gradient pixels (m..m+1, 2j..2j+2) of three frames; lane q of
// the group handles channels 8q..8q+7, so a warp's 16-byte loads cover 8 pixels x 64 contiguous bytes (4 cache lines
// per request instead of 32 with one pixel per lane: the first version of this kernel was LSU-wavefront bound).  The
// four partial sums are folded with two shuffle steps.  Weights: broadcast LDS.128 from a [kt][chunk][c][12] layout
// whose chunk pitch (100 floats) puts the four chunks of a warp on disjoint banks.
__global__ void __launch_bounds__(256) d3d_first_bwd_x2_kernel(const __nv_bfloat16* __restrict__ dpre, const float* __restrict__ w,
                                                               const float* __restrict__ sigma, float* __restrict__ dx, int B, int T, int H, int W) {
    __shared__ __align__(16) float sw[3 * 4 * 100];
    const float inv = 1.f / *sigma;
    for (int i = threadIdx.x; i < 3 * 4 * 100; i += blockDim.x) {
        const int kt = i / 400, r = i - kt * 400, ch = r / 100, r2 = r - ch * 100, c = r2 / 12, k = r2 - c * 12;
        sw[i] = (c < 8 && k < 9) ? w[(ch * 8 + c) * 27 + kt * 9 + k] * inv : 0.f;
    }
    __syncthreads();
    const int Ho = H >> 1, Wo = W >> 1, Wp = Wo >> 1;
    const int total = B * T * Ho * Wp;
    const int gidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int ch = threadIdx.x & 3;
    const bool live = gidx < total;
    const int idx = live ? gidx : total - 1;
    const int j = idx % Wp, r1 = idx / Wp, m = r1 % Ho, r2 = r1 / Ho, t = r2 % T, b = r2 / T;
    float o[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[r][c] = 0.f;
    const bool row1 = m + 1 < Ho, col2 = 2 * j + 2 < Wo;
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
        const int to = t - kt + 1;
        if (to < 0 || to >= T) continue;
        const __nv_bfloat16* base = dpre + (((static_cast<size_t>(b) * T + to) * Ho + m) * Wo + 2 * j) * 32 + ch * 8;
        float d[2][3][8];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                const bool ok = (r == 0 || row1) && (cc < 2 || col2);
                uint4 q = make_uint4(0u, 0u, 0u, 0u);
                if (ok) q = __ldg(reinterpret_cast<const uint4*>(base + (static_cast<size_t>(r) * Wo + cc) * 32));
                const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = unpack_bf16x2(qq[k]);
                    d[r][cc][2 * k] = f.x; d[r][cc][2 * k + 1] = f.y;
                }
            }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4* wp = reinterpret_cast<const float4*>(sw + kt * 400 + ch * 100 + c * 12);
            const float4 wa = wp[0], wb = wp[1], wc = wp[2];
            // w[ky][kx]: wa = (00 01 02 10), wb = (11 12 20 21), wc.x = 22
            const float w00 = wa.x, w01 = wa.y, w02 = wa.z, w10 = wa.w, w11 = wb.x, w12 = wb.y, w20 = wb.z, w21 = wb.w, w22 = wc.x;
#pragma unroll
            for (int n = 0; n < 2; ++n) {       // 2x2 block n: gradient columns n, n+1 -> dx columns 2n, 2n+1
                const float d00 = d[0][n][c], d01 = d[0][n + 1][c], d10 = d[1][n][c], d11 = d[1][n + 1][c];
                o[0][2 * n] = fmaf(d00, w11, o[0][2 * n]);
                o[0][2 * n + 1] = fmaf(d00, w12, fmaf(d01, w10, o[0][2 * n + 1]));
                o[1][2 * n] = fmaf(d00, w21, fmaf(d10, w01, o[1][2 * n]));
                o[1][2 * n + 1] = fmaf(d00, w22, fmaf(d01, w20, fmaf(d10, w02, fmaf(d11, w00, o[1][2 * n + 1]))));
            }
        }
    }
    // fold the four channel chunks (lane bits 0..1); afterwards lane q of the group stores quarter q: row q>>1, columns 2(q&1)..
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float v = o[r][c];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            o[r][c] = v;
        }
    if (!live) return;
    const int rr = ch >> 1, c0 = (ch & 1) * 2;
    float* op = dx + ((static_cast<size_t>(b) * T + t) * H + 2 * m + rr) * W + 4 * j + c0;
    const float v0 = rr ? (c0 ? o[1][2] : o[1][0]) : (c0 ? o[0][2] : o[0][0]);
    const float v1 = rr ? (c0 ? o[1][3] : o[1][1]) : (c0 ? o[0][3] : o[0][1]);
    *reinterpret_cast<float2*>(op) = make_float2(v0, v1);
}

// dx[b,c,y,x] += g[b,y,x,c] for c < C (unpacks the padded 64-channel input gradient of d2d.0)
__global__ void __launch_bounds__(256) disc_unpack_input_grad_kernel(const __nv_bfloat16* __restrict__ g, float* __restrict__ dx, int C,
                                                                     int HW, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // over [B,C,HW]
    if (i >= n) return;
    const long long pix = i % HW;
    const long long bc = i / HW;
    const long long b = bc / C, c = bc - b * C;
    dx[i] += __bfloat162float(g[(b * HW + pix) * 64 + c]);
}

// ------------------------------------------------------------------------------------------------
// Spectral-norm backward (+ un-packing of tensor-core weight gradients), one block per layer:
//   G = dL/dW_sn (packed [tap'][Cout][Cin'] or plain [Cout][K]);  dW_orig = G/sigma - (<G, W_orig>/sigma^2) u v^T
// ------------------------------------------------------------------------------------------------
constexpr int SN_BWD_BLOCKS = 64;   // blocks per layer
constexpr int SN_ROW_MAX = 3456;    // largest Cin * taps of a layer (d3d.6: 128 * 27)

// One output-channel row of G (all Cin x taps entries of `co`) -> shared memory in weight_orig order [ci][r].  The packed
// forms are read plane by plane, i.e. with consecutive threads on consecutive ci (coalesced), and transposed on the way
// in; round 1 gathered G element by element in weight_orig order (one 4-byte read per tap plane per thread, two 32-bit
// divisions and the un-packing branch per element: 148 instructions per weight, 35 us per launch).
__device__ __forceinline__ void sn_stage_row(const P2iSnGrad& L, int co, float* __restrict__ sG) {
    const int per = L.KT * L.ksize * L.ksize;
    const int K = L.Cin * per;
    if (!L.packed) {
        for (int i = threadIdx.x; i < K; i += blockDim.x) sG[i] = L.G[static_cast<size_t>(co) * K + i];
        return;
    }
    const int lc = __ffs(L.Cin) - 1;                       // Cin is a power of two (16 .. 256)
    if (!L.s2) {
        const int n = per << lc;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int tap = i >> lc, ci = i & (L.Cin - 1);
            sG[ci * per + tap] = L.G[(static_cast<size_t>(tap) * L.Cout + co) * L.cin_pad + ci];
        }
        return;
    }
    // space-to-depth layers: plane tap' = (kt*2 + dy)*2 + dx, row = [(py*2 + px)][ci]; (dy,py) = (0,1),(1,0),(1,1) <-> ky = 0,1,2
    const int n = (L.KT * 4 * 4) << lc;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ci = i & (L.Cin - 1), q = (i >> lc) & 3, tp = i >> (lc + 2);
        const int kt = tp >> 2, dy = (tp >> 1) & 1, dx = tp & 1, py = q >> 1, px = q & 1;
        const int ky = dy ? (py ? 2 : 1) : (py ? 0 : -1), kx = dx ? (px ? 2 : 1) : (px ? 0 : -1);
        if (ky < 0 || kx < 0) continue;                     // the (0,0) sub-position carries no weight
        sG[ci * per + (kt * 3 + ky) * 3 + kx] = L.G[(static_cast<size_t>(tp) * L.Cout + co) * (4 * L.Cin) + (q << lc) + ci];
    }
}

// pass 1: inner[layer] += <G, W_orig>   pass 2: dW = G/sigma - inner/sigma^2 * u v^T      (block = rows co, co + 64, ...)
__global__ void __launch_bounds__(256) sn_bwd_inner_kernel(const P2iSnGrad* __restrict__ table, float* __restrict__ inner) {
    const P2iSnGrad L = table[blockIdx.y];
    __shared__ float sG[SN_ROW_MAX];
    __shared__ float sh[32];
    const int K = L.Cin * L.KT * L.ksize * L.ksize;
    float acc = 0.f;
    for (int co = blockIdx.x; co < L.Cout; co += gridDim.x) {
        sn_stage_row(L, co, sG);
        __syncthreads();
        const float* w = L.W + static_cast<size_t>(co) * K;
        for (int i = threadIdx.x; i < K; i += blockDim.x) acc = fmaf(sG[i], w[i], acc);
        __syncthreads();
    }
    acc = blk_sum_all(acc, sh);
    if (threadIdx.x == 0 && acc != 0.f) atomicAdd(&inner[blockIdx.y], acc);
}
__global__ void __launch_bounds__(256) sn_bwd_apply_kernel(const P2iSnGrad* __restrict__ table, const float* __restrict__ inner) {
    const P2iSnGrad L = table[blockIdx.y];
    __shared__ float sG[SN_ROW_MAX];
    const int K = L.Cin * L.KT * L.ksize * L.ksize;
    const float sig = *L.sigma;
    const float c1 = 1.f / sig, c2 = inner[blockIdx.y] / (sig * sig);
    for (int co = blockIdx.x; co < L.Cout; co += gridDim.x) {
        sn_stage_row(L, co, sG);
        __syncthreads();
        float* dw = L.dW + static_cast<size_t>(co) * K;
        const float cu = c2 * L.u[co];
        for (int i = threadIdx.x; i < K; i += blockDim.x) dw[i] += sG[i] * c1 - cu * L.v[i];
        __syncthreads();
    }
}

}  // namespace p2i

using namespace p2i;

extern "C" int p2i_disc_tail_bwd(const float* dfused, const float* out2d, const float* alpha, const void* z, const float* w3,
                                 const float* sigma3, float* d_out2d, float* dalpha, void* dpre, float* dW3, float* db3, int B, int T,
                                 int h, int w, int C, int H2, int W2, void* stream) {
    P2I_CHECK_ARG(dfused && out2d && alpha && z && w3 && sigma3 && d_out2d && dpre, "disc_tail_bwd: null pointer");
    const long long n = static_cast<long long>(B) * H2 * W2;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    tail_bwd_fuse_kernel<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(dfused, out2d, alpha, d_out2d, dalpha, n);
    P2I_CHECK_LAUNCH("tail_bwd_fuse_kernel");
    const long long npix = static_cast<long long>(B) * h * w;
    tail_bwd_mean_kernel<<<static_cast<unsigned>((npix + 7) / 8), 256, C * sizeof(float), as_stream(stream)>>>(
        dfused, static_cast<const __nv_bfloat16*>(z), w3, sigma3, static_cast<__nv_bfloat16*>(dpre), dW3, db3, B, T, h, w, C, H2, W2);
    P2I_CHECK_LAUNCH("tail_bwd_mean_kernel");
    return P2I_OK;
}

extern "C" int p2i_d2d_last_bwd(const float* d_out, const void* y4, const float* w, const float* sigma, void* dpre, float* dW,
                                float* db, int B, int H, int W, int C, void* stream) {
    P2I_CHECK_ARG(d_out && y4 && w && sigma && dpre, "d2d_last_bwd: null pointer");
    const long long npix = static_cast<long long>(B) * H * W;
    if (C == 256 && npix < (1ll << 31)) {
        long long blocks = (npix + 7) / 8;
        if (blocks > sm_count()) blocks = sm_count();
        d2d_last_bwd256_kernel<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(
            d_out, static_cast<const __nv_bfloat16*>(y4), w, sigma, static_cast<__nv_bfloat16*>(dpre), dW, db, B, H, W);
        P2I_CHECK_LAUNCH("d2d_last_bwd256_kernel");
        return P2I_OK;
    }
    d2d_last_bwd_kernel<<<static_cast<unsigned>((npix + 7) / 8), 256, 18 * C * sizeof(float), as_stream(stream)>>>(
        d_out, static_cast<const __nv_bfloat16*>(y4), w, sigma, static_cast<__nv_bfloat16*>(dpre), dW, db, B, H, W, C);
    P2I_CHECK_LAUNCH("d2d_last_bwd_kernel");
    return P2I_OK;
}

extern "C" int p2i_colsum_bf16(const void* g, float* out, long long rows, int C, void* stream) {
    P2I_CHECK_ARG(g && out && C % 8 == 0 && C <= 2048 && rows > 0, "colsum_bf16: bad arguments");
    long long blocks = (rows * (C / 8) + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    while (blocks * 256 < C / 8) ++blocks;
    colsum_kernel<<<static_cast<unsigned>(blocks), 256, C * sizeof(float), as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(g), out,
                                                                                                 rows, C);
    P2I_CHECK_LAUNCH("colsum_kernel");
    return P2I_OK;
}

extern "C" int p2i_d3d_first_bwd(const void* dpre, const float* x, const float* w, const float* sigma, float* dW, float* db,
                                 float* dx, int B, int T, int H, int W, void* stream) {
    P2I_CHECK_ARG(dpre && x && w && sigma, "d3d_first_bwd: null pointer");
    P2I_CHECK_ARG(static_cast<long long>(B) * T * H * W < (1ll << 31), "d3d_first_bwd: tensor too large for 32-bit indexing");
    if (dW && db) {
        const long long total = static_cast<long long>(B) * T * (H / 2) * (W / 2);
        if (d3d_first_mma_ok(T, H, W)) {
            const int rc = d3d_first_bwd_w_mma(dpre, x, dW, db, B, T, H, W, as_stream(stream));
            if (rc) return rc;
        } else if (W % 8 == 0) {
            long long blocks = (total / 4 + 15) / 16;
            if (blocks > sm_count() * 2) blocks = sm_count() * 2;
            d3d_first_bwd_w4_kernel<<<static_cast<unsigned>(blocks), 192, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(dpre), x,
                                                                                                   dW, db, B, T, H, W);
            P2I_CHECK_LAUNCH("d3d_first_bwd_w4_kernel");
        } else {
        long long blocks = (total + 15) / 16;
        if (blocks > 148 * 4) blocks = 148 * 4;
        d3d_first_bwd_w_kernel<<<static_cast<unsigned>(blocks), 192, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(dpre), x, dW,
                                                                                              db, B, T, H, W);
        P2I_CHECK_LAUNCH("d3d_first_bwd_w_kernel");
        }
    }
    if (dx) {
        const long long total = static_cast<long long>(B) * T * H * W;
        if (W % 4 == 0 && H % 2 == 0) {
            d3d_first_bwd_x2_kernel<<<static_cast<unsigned>((total / 2 + 255) / 256), 256, 0, as_stream(stream)>>>(
                static_cast<const __nv_bfloat16*>(dpre), w, sigma, dx, B, T, H, W);
            P2I_CHECK_LAUNCH("d3d_first_bwd_x2_kernel");
            return P2I_OK;
        }
        d3d_first_bwd_x_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, as_stream(stream)>>>(
            static_cast<const __nv_bfloat16*>(dpre), w, sigma, dx, B, T, H, W);
        P2I_CHECK_LAUNCH("d3d_first_bwd_x_kernel");
    }
    return P2I_OK;
}

extern "C" int p2i_disc_unpack_input_grad(const void* g, float* dx, int B, int C, int H, int W, void* stream) {
    P2I_CHECK_ARG(g && dx, "disc_unpack_input_grad: null pointer");
    const long long n = static_cast<long long>(B) * C * H * W;
    disc_unpack_input_grad_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(g), dx, C, H * W, n);
    P2I_CHECK_LAUNCH("disc_unpack_input_grad_kernel");
    return P2I_OK;
}

extern "C" int p2i_spectral_norm_bwd(const P2iSnGrad* table_dev, int n_layers, float* inner_scratch, void* stream) {
    P2I_CHECK_ARG(table_dev && n_layers > 0 && inner_scratch, "spectral_norm_bwd: bad arguments");
    // (layers are the discriminator's: Cin a power of two, Cin * taps <= SN_ROW_MAX; the table lives on the device, so this is
    //  the caller's contract -- disc_bwd.py builds it from the module's own shapes)
    dim3 grid(SN_BWD_BLOCKS, n_layers);
    sn_bwd_inner_kernel<<<grid, 256, 0, as_stream(stream)>>>(table_dev, inner_scratch);
    P2I_CHECK_LAUNCH("sn_bwd_inner_kernel");
    sn_bwd_apply_kernel<<<grid, 256, 0, as_stream(stream)>>>(table_dev, inner_scratch);
    P2I_CHECK_LAUNCH("sn_bwd_apply_kernel");
    return P2I_OK;
}
