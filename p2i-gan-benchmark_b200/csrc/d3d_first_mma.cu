// d3d.0 -- Conv3d(1 -> 32, k 3x3x3, stride (1,2,2), pad 1) of the discriminator's 3-D branch (reference:
// p2igan_bench/models/p2igan.py:139-142) -- forward and weight gradient as warp-level tensor-core GEMMs.
//
// One input channel makes this layer a [pixels x 27] x [27 x 32] product: far too thin for a tcgen05 tile pipeline (K = 27,
// N = 32; the operand would have to be materialised as an im2col matrix in HBM), and the CUDA-core kernels it had in round 1
// (disc.cu: d3d_first_fwd4_kernel, disc_bwd.cu: d3d_first_bwd_w4_kernel) ran at 79 / 128 us per launch, alone on the GPU (no
// other kernel of the step can co-reside with them or has work at that point).  Here the im2col operand is gathered from a bf16
// shared-memory copy of the 3 input rows x (T + 2) frames that one output row needs, into mma.sync.m16n8k16 fragments
// (K padded to 32): 8 MMAs per 16 output pixels.  Both kernels are then bound by their HBM traffic (67 MB of bf16 activations /
// gradients per launch at B = 16, 128 x 128).
//
//   forward:  D[pixel, channel]  = sum_tap im2col[pixel, tap] * w[channel, tap] / sigma  (+ bias, LeakyReLU 0.2, s2d layout)
//   weights:  dW[channel, tap]  += sum_pixel dpre[pixel, channel] * im2col[pixel, tap];  tap 27 is a column of ones -> db
//
// Fragment maps (PTX ISA, mma.m16n8k16 with .bf16): g = lane >> 2, q = lane & 3
//   A (16 x 16, row):  a0 = (row g, k 2q..2q+1)  a1 = (row g+8, k 2q..)  a2 = (row g, k 2q+8..)  a3 = (row g+8, k 2q+8..)
//   B (16 x 8,  col):  b0 = (k 2q..2q+1, n g)    b1 = (k 2q+8.., n g)
//   C (16 x 8):        c0, c1 = (row g, n 2q, 2q+1)   c2, c3 = (row g+8, n 2q, 2q+1)
#include "common.h"
#include "ptx.cuh"

#include <cstdlib>

namespace p2i {

namespace {

constexpr int D3_THREADS = 256, D3_WARPS = 8;
constexpr int D3_STAGE_PITCH = 80;                      // bytes per pixel row of a warp's 16 x 32 bf16 staging tile (64 + 16 pad)
constexpr int D3_STAGE_BYTES = 16 * D3_STAGE_PITCH;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}

__device__ __forceinline__ unsigned short bf16_bits(float v) { return __bfloat16_as_ushort(__float2bfloat16(v)); }

// offset of tap k = (kt, ky, kx) inside the staged rows, relative to (frame t, column 2 * xo); -1 for the padding taps
__device__ __forceinline__ int tap_offset(int k, int WP) {
    if (k >= 27) return -1;
    const int kt = k / 9, r = k - kt * 9, ky = r / 3, kx = r - ky * 3;
    return (kt * 3 + ky) * WP + kx;
}

// Stages the input rows of output row yo as bf16:  xs[(f * 3 + r) * WP + c] = x[b, f - 1, 2 yo + r - 1, c - 1]  (0 outside),
// f in [0, T + 2), r in [0, 3), c in [0, W + 1);  WP = W + 2.
__device__ __forceinline__ void stage_rows(const float* __restrict__ x, unsigned short* xs, int b, int yo, int T, int H, int W,
                                           int WP) {
    const int W4 = W >> 2;
    const int nvec = (T + 2) * 3 * W4;
    for (int i0 = threadIdx.x; i0 < nvec; i0 += 4 * D3_THREADS) {      // four independent 16-byte loads in flight per thread
        float4 v[4];
        int dst[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * D3_THREADS;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            dst[u] = -1;
            if (i < nvec) {
                const int c4 = i % W4, fr = i / W4, f = fr / 3, r = fr - f * 3;
                const int ti = f - 1, yi = 2 * yo + r - 1;
                dst[u] = fr * WP + 4 * c4 + 1;
                if (ti >= 0 && ti < T && yi >= 0 && yi < H)
                    v[u] = __ldg(reinterpret_cast<const float4*>(x + ((static_cast<size_t>(b) * T + ti) * H + yi) * W) + c4);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (dst[u] >= 0) {
                unsigned short* d = xs + dst[u];
                d[0] = bf16_bits(v[u].x); d[1] = bf16_bits(v[u].y); d[2] = bf16_bits(v[u].z); d[3] = bf16_bits(v[u].w);
            }
    }
    for (int i = threadIdx.x; i < (T + 2) * 3; i += D3_THREADS) { xs[i * WP] = 0; xs[i * WP + W + 1] = 0; }
}

__host__ __device__ inline int xs_bytes(int T, int W) { return (((T + 2) * 3 * (W + 2) * 2) + 15) & ~15; }

}  // namespace

// grid = B * H/2 blocks (one output row each, all frames), 256 threads; warp tile = 16 consecutive output pixels of one frame
__global__ void __launch_bounds__(D3_THREADS) d3d_first_fwd_mma_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                       const float* __restrict__ sigma, const float* __restrict__ bias,
                                                                       __nv_bfloat16* __restrict__ y, int B, int T, int H, int W) {
    extern __shared__ __align__(16) unsigned char d3_smem[];
    const int WP = W + 2, Ho = H >> 1, Wo = W >> 1;
    unsigned short* xs = reinterpret_cast<unsigned short*>(d3_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    unsigned char* stage = d3_smem + xs_bytes(T, W) + warp * D3_STAGE_BYTES;
    const int b = blockIdx.x / Ho, yo = blockIdx.x - b * Ho;

    // B fragments (the whole 32 x 32 weight matrix, constant per thread), bias, and the 8 taps this lane gathers
    const float inv = 1.f / __ldg(sigma);
    uint32_t bw[2][4][2];
    int toff[2][2][2];
    float bia[4][2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k0 = 16 * s + 8 * h + 2 * q;
            toff[s][h][0] = tap_offset(k0, WP);
            toff[s][h][1] = tap_offset(k0 + 1, WP);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = 8 * j + g;
                const float w0 = (k0 < 27) ? __ldg(w + n * 27 + k0) * inv : 0.f;
                const float w1 = (k0 + 1 < 27) ? __ldg(w + n * 27 + k0 + 1) * inv : 0.f;
                bw[s][j][h] = pack_bf16x2(w0, w1);
            }
        }
#pragma unroll
    for (int j = 0; j < 4; ++j) { bia[j][0] = __ldg(bias + 8 * j + 2 * q); bia[j][1] = __ldg(bias + 8 * j + 2 * q + 1); }

    stage_rows(x, xs, b, yo, T, H, W, WP);
    __syncthreads();

    const int tiles_x = Wo >> 4, ntiles = T * tiles_x;
    for (int tile = warp; tile < ntiles; tile += D3_WARPS) {
        const int t = tile / tiles_x, x0 = (tile - t * tiles_x) << 4;
        const unsigned short* base = xs + t * 3 * WP + 2 * (x0 + g);           // output pixel x0 + g; pixel x0 + g + 8 is 16 further
        uint32_t a[2][4];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int o0 = toff[s][h][0], o1 = toff[s][h][1];
                const uint32_t lo0 = (o0 >= 0) ? base[o0] : 0u, hi0 = (o1 >= 0) ? base[o1] : 0u;
                const uint32_t lo1 = (o0 >= 0) ? base[o0 + 16] : 0u, hi1 = (o1 >= 0) ? base[o1 + 16] : 0u;
                a[s][2 * h] = lo0 | (hi0 << 16);
                a[s][2 * h + 1] = lo1 | (hi1 << 16);
            }
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc[j][0] = bia[j][0]; acc[j][1] = bia[j][1]; acc[j][2] = bia[j][0]; acc[j][3] = bia[j][1];
#pragma unroll
            for (int s = 0; s < 2; ++s) mma_bf16_16816(acc[j], a[s], bw[s][j][0], bw[s][j][1]);
        }
        // LeakyReLU, bf16, through the warp's staging tile so that every lane stores 16 contiguous bytes
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = acc[j][i] > 0.f ? acc[j][i] : 0.2f * acc[j][i];
            *reinterpret_cast<uint32_t*>(stage + g * D3_STAGE_PITCH + (8 * j + 2 * q) * 2) = pack_bf16x2(v[0], v[1]);
            *reinterpret_cast<uint32_t*>(stage + (g + 8) * D3_STAGE_PITCH + (8 * j + 2 * q) * 2) = pack_bf16x2(v[2], v[3]);
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int pix = (lane >> 2) + 8 * it, chunk = lane & 3;
            const uint4 v = *reinterpret_cast<const uint4*>(stage + pix * D3_STAGE_PITCH + chunk * 16);
            const int xo = x0 + pix;
            // s2d address: pixel (yo, xo), channel c -> [yo/2][xo/2][(yo&1)*2 + (xo&1)][c]
            __nv_bfloat16* o = y + ((((static_cast<size_t>(b) * T + t) * (Ho >> 1) + (yo >> 1)) * (Wo >> 1) + (xo >> 1)) * 4 +
                                    ((yo & 1) * 2 + (xo & 1))) * 32 + chunk * 8;
            *reinterpret_cast<uint4*>(o) = v;
        }
    }
}

// dpre bf16 [B, T, H/2, W/2, 32] (natural layout); dW f32 [32][27] and db f32 [32] are accumulated with atomics.
__global__ void __launch_bounds__(D3_THREADS) d3d_first_bwd_w_mma_kernel(const __nv_bfloat16* __restrict__ dpre,
                                                                         const float* __restrict__ x, float* __restrict__ dW,
                                                                         float* __restrict__ db, int B, int T, int H, int W) {
    extern __shared__ __align__(16) unsigned char d3_smem[];
    const int WP = W + 2, Ho = H >> 1, Wo = W >> 1;
    unsigned short* xs = reinterpret_cast<unsigned short*>(d3_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    unsigned char* stage = d3_smem + xs_bytes(T, W) + warp * D3_STAGE_BYTES;
    float* red = reinterpret_cast<float*>(d3_smem + xs_bytes(T, W) + D3_WARPS * D3_STAGE_BYTES);       // [32][33]
    const int b = blockIdx.x / Ho, yo = blockIdx.x - b * Ho;

    for (int i = threadIdx.x; i < 32 * 33; i += D3_THREADS) red[i] = 0.f;
    int toff[4];                                                             // taps 8 j + g of the B fragments
#pragma unroll
    for (int j = 0; j < 4; ++j) toff[j] = tap_offset(8 * j + g, WP);
    const bool ones = (g == 3);                                              // n-tile 3: tap 27 is the bias column
    stage_rows(x, xs, b, yo, T, H, W, WP);
    __syncthreads();

    float acc[2][4][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[m][j][i] = 0.f;

    // ldmatrix row address of this lane: matrix i = lane >> 3 covers pixels 8 (i >> 1) .. + 7, channels 8 (i & 1) .. + 7 (+ 16 m)
    const uint32_t lm_addr = smem_u32(stage) + (8 * (lane >> 4) + (lane & 7)) * D3_STAGE_PITCH + ((lane >> 3) & 1) * 16;

    const int tiles_x = Wo >> 4, ntiles = T * tiles_x;
    // gradient tile of (frame t, pixels x0 .. x0 + 15): 1 KB contiguous, two 16-byte pieces per lane; the next tile's pieces are
    // requested before the current tile is consumed
    auto tile_src = [&](int tile) {
        const int t = tile / tiles_x, x0 = (tile - t * tiles_x) << 4;
        return reinterpret_cast<const uint4*>(dpre + (((static_cast<size_t>(b) * T + t) * Ho + yo) * Wo + x0) * 32);
    };
    uint4 cur[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
    if (warp < ntiles) { const uint4* src = tile_src(warp); cur[0] = __ldg(src + lane); cur[1] = __ldg(src + lane + 32); }
    for (int tile = warp; tile < ntiles; tile += D3_WARPS) {
        const int t = tile / tiles_x, x0 = (tile - t * tiles_x) << 4;
        uint4 nxt[2] = {cur[0], cur[1]};
        if (tile + D3_WARPS < ntiles) { const uint4* src = tile_src(tile + D3_WARPS); nxt[0] = __ldg(src + lane); nxt[1] = __ldg(src + lane + 32); }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int idx = lane + 32 * it;
            *reinterpret_cast<uint4*>(stage + (idx >> 2) * D3_STAGE_PITCH + (idx & 3) * 16) = cur[it];
        }
        cur[0] = nxt[0]; cur[1] = nxt[1];
        __syncwarp();
        uint32_t a[2][4];
        ldmatrix_x4_trans(a[0], lm_addr);                                    // channels 0..15
        ldmatrix_x4_trans(a[1], lm_addr + 32);                               // channels 16..31
        const unsigned short* base = xs + t * 3 * WP + 2 * x0;               // + 2 * pixel + tap offset
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t b0 = 0u, b1 = 0u;
            const int o = toff[j];
            if (o >= 0) {
                const unsigned short* p = base + o + 4 * q;                  // pixels 2q, 2q+1 (and + 8)
                b0 = static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[2]) << 16);
                b1 = static_cast<uint32_t>(p[16]) | (static_cast<uint32_t>(p[18]) << 16);
            } else if (j == 3 && ones) {
                b0 = b1 = 0x3F803F80u;                                       // bf16 (1, 1)
            }
#pragma unroll
            for (int m = 0; m < 2; ++m) mma_bf16_16816(acc[m][j], a[m], b0, b1);
        }
    }
    // block reduction: the warps add their 32 x 32 accumulator tiles in turn, then one atomic per element
    for (int wv = 0; wv < D3_WARPS; ++wv) {
        if (warp == wv) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c0 = 16 * m + g, k0 = 8 * j + 2 * q;
                    red[c0 * 33 + k0] += acc[m][j][0];
                    red[c0 * 33 + k0 + 1] += acc[m][j][1];
                    red[(c0 + 8) * 33 + k0] += acc[m][j][2];
                    red[(c0 + 8) * 33 + k0 + 1] += acc[m][j][3];
                }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 32 * 28; i += D3_THREADS) {
        const int c = i / 28, k = i - c * 28;
        const float v = red[c * 33 + k];
        if (v != 0.f) atomicAdd(k < 27 ? &dW[c * 27 + k] : &db[c], v);
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------
static bool d3d_mma_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("P2I_D3D_MMA");                               // A/B switch: 0 = the round-1 CUDA-core kernels
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

static size_t d3d_fwd_smem(int T, int W) { return static_cast<size_t>(xs_bytes(T, W)) + D3_WARPS * D3_STAGE_BYTES; }
static size_t d3d_bww_smem(int T, int W) { return d3d_fwd_smem(T, W) + 32 * 33 * sizeof(float); }

bool d3d_first_mma_ok(int T, int H, int W) {
    return d3d_mma_enabled() && H % 4 == 0 && W % 32 == 0 && T >= 1 && d3d_bww_smem(T, W) <= 48 * 1024;
}

int d3d_first_fwd_mma(const float* x, const float* w, const float* sigma, const float* bias, void* y, int B, int T, int H, int W,
                      cudaStream_t stream) {
    d3d_first_fwd_mma_kernel<<<static_cast<unsigned>(B * (H / 2)), D3_THREADS, d3d_fwd_smem(T, W), stream>>>(
        x, w, sigma, bias, static_cast<__nv_bfloat16*>(y), B, T, H, W);
    P2I_CHECK_LAUNCH("d3d_first_fwd_mma_kernel");
    return P2I_OK;
}

int d3d_first_bwd_w_mma(const void* dpre, const float* x, float* dW, float* db, int B, int T, int H, int W, cudaStream_t stream) {
    d3d_first_bwd_w_mma_kernel<<<static_cast<unsigned>(B * (H / 2)), D3_THREADS, d3d_bww_smem(T, W), stream>>>(
        static_cast<const __nv_bfloat16*>(dpre), x, dW, db, B, T, H, W);
    P2I_CHECK_LAUNCH("d3d_first_bwd_w_mma_kernel");
    return P2I_OK;
}

}  // namespace p2i
