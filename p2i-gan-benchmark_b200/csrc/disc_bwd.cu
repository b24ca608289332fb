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
    if (i < stride) {
        for (; i < total; i += stride) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(g) + i);
            const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = unpack_bf16x2(qq[k]);
                acc[2 * k] += f.x;
                acc[2 * k + 1] += f.y;
            }
        }
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
__device__ __forceinline__ size_t packed_index(const P2iSnGrad& L, int co, int ci, int r) {
    if (!L.packed) return (static_cast<size_t>(co) * L.Cin + ci) * (L.KT * L.ksize * L.ksize) + r;
    const int k = L.ksize, kk = k * k;
    const int kt = r / kk, ky = (r % kk) / k, kx = r % k;
    if (L.s2) {
        const int dy = ky == 0 ? 0 : 1, py = ky == 1 ? 0 : 1, dx = kx == 0 ? 0 : 1, px = kx == 1 ? 0 : 1;
        const int tap = (kt * 2 + dy) * 2 + dx;
        return (static_cast<size_t>(tap) * L.Cout + co) * (4 * L.Cin) + (py * 2 + px) * L.Cin + ci;
    }
    const int tap = (kt * k + ky) * k + kx;
    return (static_cast<size_t>(tap) * L.Cout + co) * L.cin_pad + ci;
}

constexpr int SN_BWD_BLOCKS = 32;   // blocks per layer
// pass 1: inner[layer] += <G, W_orig>   pass 2: dW = G/sigma - inner/sigma^2 * u v^T
__global__ void __launch_bounds__(256) sn_bwd_inner_kernel(const P2iSnGrad* __restrict__ table, float* __restrict__ inner) {
    const P2iSnGrad L = table[blockIdx.y];
    __shared__ float sh[32];
    const int per = L.KT * L.ksize * L.ksize;
    const int K = L.Cin * per;
    const long long total = static_cast<long long>(L.Cout) * K;
    float acc = 0.f;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(e % per), ci = static_cast<int>((e / per) % L.Cin), co = static_cast<int>(e / K);
        acc += L.G[packed_index(L, co, ci, r)] * L.W[e];
    }
    acc = blk_sum_all(acc, sh);
    if (threadIdx.x == 0 && acc != 0.f) atomicAdd(&inner[blockIdx.y], acc);
}
__global__ void __launch_bounds__(256) sn_bwd_apply_kernel(const P2iSnGrad* __restrict__ table, const float* __restrict__ inner) {
    const P2iSnGrad L = table[blockIdx.y];
    const int per = L.KT * L.ksize * L.ksize;
    const int K = L.Cin * per;
    const long long total = static_cast<long long>(L.Cout) * K;
    const float sig = *L.sigma;
    const float c1 = 1.f / sig, c2 = inner[blockIdx.y] / (sig * sig);
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(e % per), ci = static_cast<int>((e / per) % L.Cin), co = static_cast<int>(e / K);
        L.dW[e] += L.G[packed_index(L, co, ci, r)] * c1 - c2 * L.u[co] * L.v[e - static_cast<long long>(co) * K];
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
        long long blocks = (total + 15) / 16;
        if (blocks > 148 * 4) blocks = 148 * 4;
        d3d_first_bwd_w_kernel<<<static_cast<unsigned>(blocks), 192, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(dpre), x, dW,
                                                                                              db, B, T, H, W);
        P2I_CHECK_LAUNCH("d3d_first_bwd_w_kernel");
    }
    if (dx) {
        const long long total = static_cast<long long>(B) * T * H * W;
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
    dim3 grid(SN_BWD_BLOCKS, n_layers);
    sn_bwd_inner_kernel<<<grid, 256, 0, as_stream(stream)>>>(table_dev, inner_scratch);
    P2I_CHECK_LAUNCH("sn_bwd_inner_kernel");
    sn_bwd_apply_kernel<<<grid, 256, 0, as_stream(stream)>>>(table_dev, inner_scratch);
    P2I_CHECK_LAUNCH("sn_bwd_apply_kernel");
    return P2I_OK;
}
