// DO-Conv weight composition (p2igan_bench/modules/deconv_pytorch.py:111-132):
//   DoW[o,i,m] = sum_s (D + D_diag)[i,m,s] * W[o,i,s]        (groups = 1, 3x3: m,s in 0..8)
// written straight into the bf16 operand layouts of the implicit-GEMM kernels:
//   out   [m][o][i]            forward B operand (K = input channels contiguous)
//   out_t [8-m][i][o]          dgrad   B operand (taps flipped, K = output channels contiguous)
// One launch covers a whole table of layers (33 tiny einsum launches in the reference).
#include "common.h"
#include "ptx.cuh"

namespace p2i {

// block = 32 input channels x 32 output channels (thread = input channel x 4 output channels); the 9 x 32 x 32 results
// are staged in shared memory so that BOTH operand layouts are written with contiguous 64-byte warp stores
// (out: input channels contiguous, out_t: output channels contiguous).  grid = (C/32, C/32, layer)
__global__ void __launch_bounds__(256) doconv_compose_kernel(const P2iDoLayer* __restrict__ table) {
    const P2iDoLayer L = table[blockIdx.z];
    const int C = L.channels;
    const int i0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
    if (i0 >= C || o0 >= C) return;
    __shared__ float sD[32][82];                 // (D + D_diag)[i0 + ii][m*9 + s], padded against bank conflicts
    __shared__ float sW[8][288];                 // per-warp staging of one output channel's 32 x 9 weights
    __shared__ __nv_bfloat16 sT[9][32][34];      // [m][o][i], row pitch 17 words: conflict-free both ways
    for (int e = threadIdx.x; e < 32 * 81; e += 256) {
        const int ii = e / 81, r = e - ii * 81;
        const size_t g = static_cast<size_t>(i0 + ii) * 81 + r;
        sD[ii][r] = L.D[g] + L.D_diag[g];
    }
    __syncthreads();
    const int ii = threadIdx.x & 31, oo = threadIdx.x >> 5;
    for (int q = 0; q < 4; ++q) {
        const int ol = oo + 8 * q;
        const float* wrow = L.W + (static_cast<size_t>(o0 + ol) * C + i0) * 9;      // 288 contiguous floats
#pragma unroll
        for (int k = 0; k < 9; ++k) sW[oo][ii + 32 * k] = __ldg(wrow + ii + 32 * k);
        __syncwarp();
        float w[9];
#pragma unroll
        for (int s9 = 0; s9 < 9; ++s9) w[s9] = sW[oo][ii * 9 + s9];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 9; ++m) {
            float acc = 0.f;
#pragma unroll
            for (int s9 = 0; s9 < 9; ++s9) acc = fmaf(sD[ii][m * 9 + s9], w[s9], acc);
            sT[m][ol][ii] = __float2bfloat16(acc);
        }
    }
    __syncthreads();
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(L.out);
    __nv_bfloat16* out_t = static_cast<__nv_bfloat16*>(L.out_t);
    for (int e = threadIdx.x; e < 9 * 32 * 32; e += 256) {
        const int lo = e & 31, hi = (e >> 5) & 31, m = e >> 10;
        out[(static_cast<size_t>(m) * C + o0 + hi) * C + i0 + lo] = sT[m][hi][lo];
        if (out_t) out_t[(static_cast<size_t>(8 - m) * C + i0 + hi) * C + o0 + lo] = sT[m][lo][hi];
    }
}

// Convsin: W [64,4,9] is raw-reshaped to [16,16,9] before the einsum (deconv_pytorch.py:119), which pairs
// weight row (oc, icl) with D row i = (oc % 4) * 4 + icl.  out f32 [64][4][9].
__global__ void doconv_compose_stem_kernel(const float* __restrict__ W, const float* __restrict__ D,
                                           const float* __restrict__ Dd, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (oc, icl, m)
    if (idx >= 64 * 4 * 9) return;
    const int m = idx % 9, icl = (idx / 9) % 4, oc = idx / 36;
    const int i = (oc % 4) * 4 + icl;
    const float* wp = W + (oc * 4 + icl) * 9;
    float acc = 0.f;
#pragma unroll
    for (int s = 0; s < 9; ++s) acc = fmaf(D[(i * 9 + m) * 9 + s] + Dd[(i * 9 + m) * 9 + s], wp[s], acc);
    out[idx] = acc;
}

}  // namespace p2i

using namespace p2i;

extern "C" int p2i_doconv_compose_fwd(const P2iDoLayer* table_dev, int n_layers, int max_channels, void* stream) {
    P2I_CHECK_ARG(table_dev && n_layers > 0, "doconv_compose: empty table");
    P2I_CHECK_ARG(max_channels % 64 == 0 && max_channels > 0, "doconv_compose: channels must be a multiple of 64");
    dim3 grid(max_channels / 32, max_channels / 32, n_layers);
    doconv_compose_kernel<<<grid, 256, 0, as_stream(stream)>>>(table_dev);
    P2I_CHECK_LAUNCH("doconv_compose_kernel");
    return P2I_OK;
}

extern "C" int p2i_doconv_compose_stem_fwd(const float* W, const float* D, const float* D_diag, float* out,
                                           void* stream) {
    P2I_CHECK_ARG(W && D && D_diag && out, "doconv_compose_stem: null pointer");
    doconv_compose_stem_kernel<<<cdiv(64 * 4 * 9, 256), 256, 0, as_stream(stream)>>>(W, D, D_diag, out);
    P2I_CHECK_LAUNCH("doconv_compose_stem_kernel");
    return P2I_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward of the composition.  dDoW is the fp32 wgrad output [m][o][i] (same layout as `out`).
//   dW[o,i,s] = sum_m dDoW[m,o,i] * (D + D_diag)[i,m,s]
//   dD[i,m,s] = sum_o dDoW[m,o,i] * W[o,i,s]
// ------------------------------------------------------------------------------------------------
namespace p2i {

// One kernel for both gradients.  block = 32 input channels x 64 output channels (thread = input channel x 8 output
// channels): per output channel the warp loads the 288 contiguous weights and read-modify-writes the 288 contiguous
// dW values through a per-warp staging buffer (coalesced), keeps the 9x9 dD partial of its input channel in
// registers, and the block reduces the 8 warps in turn before one global atomic per element.
__global__ void __launch_bounds__(256, 2) doconv_bwd_kernel(const P2iDoGrad* __restrict__ table) {
    const P2iDoGrad L = table[blockIdx.z];
    const int C = L.channels;
    const int i0 = blockIdx.x * 32, o0 = blockIdx.y * 64;
    if (i0 >= C || o0 >= C) return;
    __shared__ float sD[32][82];
    __shared__ float red[32][82];
    __shared__ float sW[8][288];
    for (int e = threadIdx.x; e < 32 * 81; e += 256) {
        const int a = e / 81, r = e - a * 81;
        const size_t g = static_cast<size_t>(i0 + a) * 81 + r;
        sD[a][r] = L.D[g] + L.D_diag[g];
        red[a][r] = 0.f;
    }
    __syncthreads();
    const int ii = threadIdx.x & 31, oo = threadIdx.x >> 5;
    const int i = i0 + ii;
    float acc[81];
#pragma unroll
    for (int k = 0; k < 81; ++k) acc[k] = 0.f;
    for (int q = 0; q < 8; ++q) {
        const int o = o0 + oo + 8 * q;
        const size_t rowoff = (static_cast<size_t>(o) * C + i0) * 9;
        // all 27 global loads of this output channel (gradient, weights, old dW) are issued before anything waits
        float g[9], w[9], wst[9], dwo[9];
#pragma unroll
        for (int m = 0; m < 9; ++m) g[m] = __ldg(L.dDoW + (static_cast<size_t>(m) * C + o) * C + i);
#pragma unroll
        for (int k = 0; k < 9; ++k) wst[k] = __ldg(L.W + rowoff + ii + 32 * k);
#pragma unroll
        for (int k = 0; k < 9; ++k) dwo[k] = L.dW[rowoff + ii + 32 * k];
#pragma unroll
        for (int k = 0; k < 9; ++k) sW[oo][ii + 32 * k] = wst[k];
        __syncwarp();
#pragma unroll
        for (int s9 = 0; s9 < 9; ++s9) w[s9] = sW[oo][ii * 9 + s9];
        __syncwarp();
#pragma unroll
        for (int s9 = 0; s9 < 9; ++s9) {
            float a = 0.f;
#pragma unroll
            for (int m = 0; m < 9; ++m) {
                a = fmaf(g[m], sD[ii][m * 9 + s9], a);
                acc[m * 9 + s9] = fmaf(g[m], w[s9], acc[m * 9 + s9]);
            }
            sW[oo][ii * 9 + s9] = a;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 9; ++k) L.dW[rowoff + ii + 32 * k] = dwo[k] + sW[oo][ii + 32 * k];
        __syncwarp();
    }
    for (int wv = 0; wv < 8; ++wv) {
        if (oo == wv) {
#pragma unroll
            for (int k = 0; k < 81; ++k) red[ii][k] += acc[k];
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < 32 * 81; e += 256) {
        const int a = e / 81, r = e - a * 81;
        if (red[a][r] != 0.f) atomicAdd(&L.dD[static_cast<size_t>(i0 + a) * 81 + r], red[a][r]);
    }
}

// stem: dDoW f32 [64][4][9] -> dW [64,4,9], dD [16,9,9] with the reference's raw-reshape row pairing
__global__ void doconv_bwd_stem_kernel(const float* __restrict__ W, const float* __restrict__ D, const float* __restrict__ Dd,
                                       const float* __restrict__ g, float* __restrict__ dW, float* __restrict__ dD) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 64 * 4 * 9) {            // dW[oc][icl][s] = sum_m g[oc][icl][m] * Dsum[i][m][s]
        const int s = idx % 9, icl = (idx / 9) % 4, oc = idx / 36;
        const int i = (oc % 4) * 4 + icl;
        float acc = 0.f;
#pragma unroll
        for (int m = 0; m < 9; ++m) acc = fmaf(g[(oc * 4 + icl) * 9 + m], D[(i * 9 + m) * 9 + s] + Dd[(i * 9 + m) * 9 + s], acc);
        dW[idx] += acc;
    }
    if (idx < 16 * 81) {               // dD[i][m][s] = sum over (oc, icl) with (oc%4)*4+icl == i
        const int s = idx % 9, m = (idx / 9) % 9, i = idx / 81;
        const int icl = i % 4, r = i / 4;
        float acc = 0.f;
        for (int oc = r; oc < 64; oc += 4) acc = fmaf(g[(oc * 4 + icl) * 9 + m], W[(oc * 4 + icl) * 9 + s], acc);
        dD[idx] += acc;
    }
}

}  // namespace p2i

extern "C" int p2i_doconv_compose_bwd(const P2iDoGrad* table_dev, int n_layers, int max_channels, void* stream) {
    P2I_CHECK_ARG(table_dev && n_layers > 0 && max_channels % 64 == 0, "doconv_compose_bwd: bad table");
    dim3 gw(max_channels / 32, max_channels / 64, n_layers);
    p2i::doconv_bwd_kernel<<<gw, 256, 0, p2i::as_stream(stream)>>>(table_dev);
    P2I_CHECK_LAUNCH("doconv_bwd_kernel");
    return P2I_OK;
}

extern "C" int p2i_doconv_compose_stem_bwd(const float* W, const float* D, const float* D_diag, const float* dDoW,
                                           float* dW, float* dD, void* stream) {
    P2I_CHECK_ARG(W && D && D_diag && dDoW && dW && dD, "doconv_compose_stem_bwd: null pointer");
    p2i::doconv_bwd_stem_kernel<<<p2i::cdiv(64 * 4 * 9, 256), 256, 0, p2i::as_stream(stream)>>>(W, D, D_diag, dDoW, dW, dD);
    P2I_CHECK_LAUNCH("doconv_bwd_stem_kernel");
    return P2I_OK;
}
