/* libp2i_sm100a.so -- C ABI of the B200-native P2I-GAN hot path.
 *
 * The reference (NTU-CompHydroMet-Lab/P2I-GAN-benchmark) is pure Python/PyTorch and exposes no FFI;
 * its seam is the Python module API.  Each entry point below replaces the ATen/cuDNN
 * work behind one reference call site, cited as file:line under the reference root.  The Python
 * host layer (p2i-gan-benchmark_b200/p2igan_b200) binds these with ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain device pointers + sizes; no torch types; `stream` is a cudaStream_t passed as void*.
 *   - returns 0 on success, negative P2I_ERR_* otherwise; p2i_last_error() gives the message.
 *   - never allocates, never synchronises, never throws; re-entrant per stream.
 *   - trunk activations are NHWC bf16 ("cl" = channels-last); model I/O is NCHW fp32 as in the reference.
 *   - bf16 buffers are passed as `void*` (uint16 storage).
 */
#ifndef P2I_B200_H
#define P2I_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P2I_OK 0
#define P2I_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define P2I_ERR_CUDA (-2)    /* CUDA runtime / driver error at launch */

int p2i_abi_version(void);
const char* p2i_last_error(void);
/* Number of kernel launches issued by this library since load (bench.py's `gpu_launches`). */
long long p2i_launch_count(void);
/* Debug: which kernel variant the calling thread's most recent conv entry point launched.
 *   2000000 + MB*100000 + NT*100 + RES*10 + CG : conv_halo_kernel<NT,RES,CG> (M blocking MB)
 *   1000000 + NT*100                           : first-generation conv_igemm_kernel<NT>
 *   3000000 + nt*100 + stacked*10              : conv_wgrad_kernel        4000000 + mode : conv_wgrad2_kernel
 * The parity tests assert that the shapes they run select the instantiations the benchmark's train step launches. */
int p2i_conv_last_variant(void);
/* Programmatic dependent launch of the tensor-core conv kernels (prologue + weight fetch of a launch overlap the tail of
 * its stream predecessor; CUDA-graph capturable).  Default off (measured neutral on the training step); 1 switches it on for A/B measurements. */
int p2i_set_pdl(int on);

/* ---------------------------------------------------------------------------------------------
 * InputBlock  (p2igan_bench/modules/layer.py:307-361; gate :296-304; idw_3d_knn :259-293)
 * ------------------------------------------------------------------------------------------- */

/* torch.nonzero(mask_b > 0) per sample, lexicographic (t,y,x) order  (layer.py:329).
 * masks [B,T,H,W] f32 -> pts [B,cap] int32 linear index t*H*W+y*W+x (ascending), counts [B].
 * Points beyond `cap` are dropped (counts is clamped to cap).
 * src [B] (optional, may be NULL): src[b] = 0 when sample b observes exactly the points of sample 0
 * (the 'stis' gauge mask repeats one pattern over the batch, data/sti_dataset.py:104-117), else b;
 * it lets the neighbour search below run once per distinct pattern. */
int p2i_points_extract(const float* masks, int B, int T, int H, int W, int* pts, int* counts, int* src, int cap,
                       void* stream);

/* Two AttentionBlock gates evaluated only at the observed points (layer.py:318-322,344):
 * x <- relu(x + x*(W x + b)) over the 16 frames of a pixel.  masked [B,16,H,W] f32, vals [B,cap] f32.
 * If gate_l1 != NULL it receives the layer-1 activations [B,cap,16]. */
int p2i_gate_points_fwd(const float* masked, const int* pts, const int* counts, int cap, const float* w0,
                        const float* b0, const float* w1, const float* b1, float* vals, float* gate_l1, int B, int T,
                        int H, int W, void* stream);
/* d(vals) -> d(w0,b0,w1,b1), accumulated with atomics into zero-initialised f32 buffers. */
int p2i_gate_points_bwd(const float* masked, const int* pts, const int* counts, int cap, const float* w0,
                        const float* b0, const float* w1, const float* b1, const float* dvals, float* dw0, float* db0,
                        float* dw1, float* db1, int B, int T, int H, int W, void* stream);

/* 4-nearest-neighbour inverse-distance interpolation onto the full (T,H,W) grid (layer.py:259-293,
 * k=4, rho=2).  Exact integer distance ordering, ties -> smaller point index.  out [B,T,H,W] f32.
 * nbr_idx [B,T*H*W,4] int32 / nbr_w [B,T*H*W,4] f32: neighbour table (written when search != 0,
 * reused when search == 0, e.g. while the mask is unchanged between steps).
 * Samples with zero points give zeros (layer.py:330-332). */
int p2i_idw_knn_fwd(const int* pts, const float* vals, const int* counts, const int* src, int cap, float* out,
                    int* nbr_idx, float* nbr_w, int B, int T, int H, int W, float tau, int search, const int* reuse_flag,
                    int* cache_idx, float* cache_w, void* stream);
/* Device-side cache of sample 0's neighbour table across calls (a static gauge mask makes every step search the same
 * pattern; the reference recomputes cdist + topk every forward, layer.py:280-281).  Compares sample 0's points with
 * cache_pts [cap] / cache_count [1]; writes flag[0] = 1 when identical, else 0 and replaces the cached points.  With
 * reuse_flag / cache_idx [T*H*W,4] / cache_w passed to p2i_idw_knn_fwd (all three may be NULL), sample 0's rows of
 * the table are copied from the cache when flag[0] == 1, else searched and stored into the cache.  No host sync. */
int p2i_idw_cache_check(const int* pts, const int* counts, int cap, int* cache_pts, int* cache_count, int* flag,
                        void* stream);
/* dvals[b, idx] += w * dout ; dvals [B,cap] must be zero-initialised. */
int p2i_idw_knn_bwd(const float* dout, const int* nbr_idx, const float* nbr_w, const int* counts, const int* src,
                    float* dvals, int cap, int B, int T, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DO-Conv weight composition  (p2igan_bench/modules/deconv_pytorch.py:111-132)
 * ------------------------------------------------------------------------------------------- */

typedef struct P2iDoLayer {
    const float* W;      /* [C, C, 9]  */
    const float* D;      /* [C, 9, 9]  */
    const float* D_diag; /* [C, 9, 9]  */
    void* out;           /* bf16 [9][C_out][C_in] -- the implicit-GEMM B operand (K-major) */
    void* out_t;         /* bf16 [9][C_in][C_out], taps flipped -- dgrad operand; may be NULL */
    int channels;
    int _pad;
} P2iDoLayer;

/* One launch for a whole table of groups=1, 3x3 DO-Conv layers (device-resident table). */
int p2i_doconv_compose_fwd(const P2iDoLayer* table_dev, int n_layers, int max_channels, void* stream);
typedef struct P2iDoGrad {
    const float* W;      /* [C, C, 9] */
    const float* D;      /* [C, 9, 9] */
    const float* D_diag; /* [C, 9, 9] */
    const float* dDoW;   /* f32 [9][C][C]: output of p2i_conv2d_wgrad */
    float* dW;           /* [C, C, 9] (accumulated: +=) */
    float* dD;           /* [C, 9, 9] (accumulated: +=) */
    int channels;
    int _pad;
} P2iDoGrad;

/* Backward of the composition for a table of layers (torch.einsum backward at deconv_pytorch.py:124):
 * dW[o,i,s] = sum_m dDoW[m,o,i] (D+D_diag)[i,m,s] ;  dD[i,m,s] = sum_o dDoW[m,o,i] W[o,i,s]. */
int p2i_doconv_compose_bwd(const P2iDoGrad* table_dev, int n_layers, int max_channels, void* stream);

/* Grouped stem variant (Convsin: 16->64, k3, groups 4) keeping the reference's raw-reshape row
 * pairing (deconv_pytorch.py:119-124).  out f32 [64,4,9]. */
int p2i_doconv_compose_stem_fwd(const float* W, const float* D, const float* D_diag, float* out, void* stream);
/* dDoW f32 [64,4,9] -> dW [64,4,9], dD [16,9,9] (both accumulated: +=). */
int p2i_doconv_compose_stem_bwd(const float* W, const float* D, const float* D_diag, const float* dDoW, float* dW,
                                float* dD, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Convolutions as tcgen05 implicit GEMM  (F.conv2d at deconv_pytorch.py:108; layer.py:129-135)
 * ------------------------------------------------------------------------------------------- */

/* y = mask( act( conv(x, w) + bias + residual ) );
 * x [B,H,W,Cin] bf16, w [k*k][Cout][Cin] bf16, y [B,H,W,Cout] bf16.  ksize in {1,3}, stride 1, zero padding k/2.
 * Cin % 64 == 0, Cout % 64 == 0.  residual / mask (bf16 [B,H,W,Cout]) and bias (f32 [Cout]) may be NULL.
 * mask: the output is zeroed where mask <= 0 -- the ReLU backward fused into a dgrad launch.
 * The data gradient of the same convolution is this entry point called with dY as x and the
 * transposed, tap-flipped weights ([k*k][Cin][Cout], produced by p2i_doconv_compose_fwd's out_t). */
#define P2I_CONV_RELU 1
#define P2I_CONV_LEAKY 2 /* LeakyReLU(0.2) */
int p2i_conv2d_igemm_fwd(const void* x, const void* w, const void* residual, const void* mask, const float* bias,
                         void* y, int B, int H, int W, int Cin, int Cout, int ksize, int flags, void* stream);
/* CUDA-core direct convolution with identical semantics (test/triage only; never on the product path). */
int p2i_conv2d_direct_fwd(const void* x, const void* w, const void* residual, const void* mask, const float* bias,
                          void* y, int B, int H, int W, int Cin, int Cout, int ksize, int flags, void* stream);

/* General form used by the discriminator: 5-D activations [samples, T, H, W, C] (T = 1 for 2-D), temporal taps,
 * k in {1,2,3} with explicit left padding, temporal stride, fused bias / residual / activation / gradient mask and
 * the space-to-depth output modes that turn the reference's stride-2 convs (models/p2igan.py:123-139) into
 * stride-1 k=2 convs.  Weights: bf16 [kt*k*k][Cout][Cin], tap index (kt*k + ky)*k + kx. */
typedef struct P2iConvDesc {
    int samples, T_in, T_out, H, W, Cin, Cout; /* H, W: output (= input) spatial size of the GEMM grid            */
    int kt, ksize;                             /* temporal taps (1|3), spatial taps k (1|2|3)                      */
    int pad, pad_t;                            /* x_in = x + kx - pad ; t_in = stride_t*t_out + kt - pad_t         */
    int stride_t;                              /* 1 | 2                                                            */
    int t_transposed;                          /* 1: data gradient of a temporal stride_t conv:
                                                  t_in = (t_out + pad_t - kt)/stride_t when divisible, else skipped */
    int act;                                   /* 0 none | 1 ReLU | 2 LeakyReLU(0.2)                               */
    int mask_mode;                             /* 0 none | 1 zero where mask<=0 | 2 scale by 0.2 where mask<=0     */
    int out_mode;                              /* 0 natural | 1 space-to-depth pack | 2 space-to-depth unpack      */
} P2iConvDesc;
int p2i_conv_igemm(const void* x, const void* w, const P2iConvDesc* desc, const void* residual, const void* mask,
                   const float* bias, void* y, void* stream);
int p2i_conv_wgrad(const void* x, const void* dy, float* dW, const P2iConvDesc* desc, void* stream);
/* Kernel selection for p2i_conv_igemm / p2i_conv2d_igemm_fwd (A/B measurements and tests): 0 = automatic (the halo
 * kernel of conv_igemm_halo.cu when the shape is eligible, else the first-generation kernel), 1 = first generation
 * only, 2 = halo only (ineligible shapes fail with P2I_ERR_INVALID), 3 / 4 = halo only, forced to single CTAs /
 * CTA pairs (tcgen05 cta_group::2). */
int p2i_set_conv_impl(int impl);
/* Same for p2i_conv_wgrad / p2i_conv2d_wgrad: 0 = automatic and 1 = first-generation kernel (the measured winner once
 * the K split is limited to one wave); 2 = the experimental second-generation variants where eligible (all nine taps
 * per CTA for 64 -> 64 channel 3x3 layers, flipped GEMM with 16-byte vector reductions otherwise), kept for A/B runs. */
int p2i_set_wgrad_impl(int impl);

/* Weight gradient: dW[tap][co][ci] += sum_pix dy[pix][co] * x[pix+tap][ci]  (fp32 [k*k][Cout][Cin], atomically
 * accumulated: the caller zero-fills).  x [B,H,W,Cin], dy [B,H,W,Cout] bf16.  Cin in {64} or % 128 == 0. */
int p2i_conv2d_wgrad(const void* x, const void* dy, float* dW, int B, int H, int W, int Cin, int Cout, int ksize,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * Generator glue  (p2igan_bench/models/p2igan.py:72-112)
 * ------------------------------------------------------------------------------------------- */

/* Convsin (grouped 3x3, 16->64) + x.repeat_interleave(4)  (p2igan.py:79).
 * x [B,16,H,W] f32 NCHW, w f32 [64,4,9] -> y [B,H,W,64] bf16. */
int p2i_stem_fwd(const float* x, const float* w, void* y, int B, int H, int W, void* stream);

/* DownsampleDuplicateChannels x3 fused (layer.py:200-214; p2igan.py:81-83): from the stem output
 * [B,H,W,64] produce x4 [B,H/4,W/4,256] and x8 [B,H/8,W/8,512] (x2 is never consumed, p2igan.py:100). */
int p2i_pyramid_fwd(const void* stem, void* x4, void* x8, int B, int H, int W, void* stream);
/* One standalone DownsampleDuplicateChannels level on the reference's NCHW f32 layout (layer.py:205-214): x [B,C,H,W] ->
 * y [B,2C,H/2,W/2], y[:,2c] = y[:,2c+1] = max_pool2d(x[:,c], 2, 2).  The generator uses the fused pyramid above instead. */
int p2i_downsample_dup_fwd(const float* x, float* y, int B, int C, int H, int W, void* stream);
/* dx [B,C,H,W] (overwritten) = gradient routed to the first arg-max of every 2x2 window, summed over the two duplicates. */
int p2i_downsample_dup_bwd(const float* x, const float* dy, float* dx, int B, int C, int H, int W, void* stream);

/* UPPos after hoisting the 1x1 projection below the upsample (layer.py:384-399):
 * out = relu( 2*sigmoid(pos) * bilinear_x2_align_corners(z) + bias ) [+ skip]
 * z [B,h,w,C] bf16 (projection output, no bias), pos f32 [2h,2w], bias f32 [C], out/skip [B,2h,2w,C] bf16. */
int p2i_upmod_fwd(const void* z, const float* pos, const float* bias, const void* skip, void* out, int B, int h, int w,
                  int C, void* stream);

/* ConvsOut (1x1, groups 4, 64->16) + tanh  (p2igan.py:109-111).
 * x [B,H,W,64] bf16, w f32 [16,16] -> out f32 [B,16,H,W]; pre (optional) receives the pre-tanh z. */
int p2i_head_fwd(const void* x, const float* w, float* out, float* pre, int B, int H, int W, void* stream);

/* ---- backward of the glue ---- */
/* dout, out f32 [B,16,H,W] (out = tanh output), x [B,H,W,64] bf16 -> dx [B,H,W,64] bf16; dw f32 [16,16] accumulated. */
int p2i_head_bwd(const float* dout, const float* out, const void* x, const float* w, void* dx, float* dw, int B, int H,
                 int W, void* stream);
/* UPPos tail backward. dout [B,2h,2w,C] bf16 (gradient of the pre-skip ReLU output; the skip branch receives the
 * same tensor).  g_scratch [B,2h,2w,C] bf16 workspace.  dz [B,h,w,C] bf16 (overwritten); dbias f32 [C] and
 * dpos f32 [2h,2w] are accumulated (caller zero-fills). */
int p2i_upmod_bwd(const void* z, const float* pos, const float* bias, const void* dout, void* g_scratch, void* dz,
                  float* dbias, float* dpos, int B, int h, int w, int C, void* stream);
/* d(x4) [B,H/4,W/4,256], d(x8) [B,H/8,W/8,512] -> d(stem) [B,H,W,64] bf16 (overwritten), routed to the arg-max
 * pixels of the saved stem output (max_pool2d backward + repeat_interleave backward, layer.py:210-213). */
int p2i_pyramid_bwd(const void* stem, const void* dx4, const void* dx8, void* dstem, int B, int H, int W, void* stream);
/* dy [B,H,W,64] bf16, x f32 [B,16,H,W], w f32 [64,4,9] -> dx f32 [B,16,H,W] (overwritten), dw f32 [64,4,9] accumulated.
 * dx or dw may be NULL: only the other gradient is computed (two independent kernels). */
int p2i_stem_bwd(const void* dy, const float* x, const float* w, float* dx, float* dw, int B, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Discriminator  (p2igan_bench/models/p2igan.py:115-173; C2/C3 layer.py:402-407)
 * ------------------------------------------------------------------------------------------- */

typedef struct P2iSnLayer {
    const float* W; /* weight_orig viewed as [rows = Cout][cols = Cin*kt*k*k] */
    float* u;       /* weight_u [rows]  (updated in place when training)       */
    float* v;       /* weight_v [cols]  (updated in place when training)       */
    float* sigma;   /* out: u . (W v)                                          */
    float* u_snap;  /* optional copies of the u / v used for sigma (kept for the backward of THIS call, since */
    float* v_snap;  /* the next forward updates weight_u / weight_v in place); may be NULL                    */
    float* scratch; /* f32 [2 + rows], zero-filled once by the caller (left clean after every call)           */
    int rows, cols;
} P2iSnLayer;
/* torch.nn.utils.spectral_norm's pre-forward hook for a table of layers, one launch: training != 0 runs ONE
 * power iteration (v <- normalize(W^T u), u <- normalize(W v), eps 1e-12) before sigma; eval uses stored u, v.
 * max_rows / max_cols: the largest rows / cols in the table (grid sizing). */
int p2i_spectral_norm(const P2iSnLayer* table_dev, int n_layers, int max_rows, int max_cols, int training, void* stream);

typedef struct P2iPackLayer {
    const float* W;     /* weight_orig [Cout][Cin][kt][k][k]                                   */
    const float* sigma; /* device scalar (NULL = 1)                                              */
    void* out;          /* bf16 [taps'][Cout][Cin']  forward operand (may be NULL)               */
    void* out_t;        /* bf16 [taps'][Cin'][Cout]  data-gradient operand, taps flipped (may be NULL) */
    int Cout, Cin, KT, ksize;
    int s2;             /* 1: spatial stride-2 layer -> k=2 taps over a space-to-depth input (Cin' = 4 Cin) */
    int cin_pad;        /* Cin' for s2 == 0 (>= Cin; padded input channels stay zero)            */
    int keep_t;         /* 1: do not flip temporal taps in out_t (temporally transposed dgrad)   */
    int _pad;
} P2iPackLayer;
/* weight_orig / sigma -> bf16 GEMM operands for a table of layers (buffers zero-initialised once by the caller). */
int p2i_disc_pack_weights(const P2iPackLayer* table_dev, int n_layers, void* stream);

/* x f32 [B,C,H,W] -> bf16 [B,H,W,64], channels >= C zero (input of d2d.0). */
int p2i_disc_pack_input(const float* x, void* y, int B, int C, int H, int W, void* stream);
/* d3d.0: Conv3d(1->32, 3x3x3, stride (1,2,2), pad 1) + bias + LeakyReLU(0.2) -> bf16 space-to-depth
 * [B,T,H/4,W/4,128].  x f32 [B,T,H,W]; w = weight_orig f32 [32,27], divided by *sigma in-kernel. */
int p2i_d3d_first_fwd(const float* x, const float* w, const float* sigma, const float* bias, void* y, int B, int T, int H,
                      int W, void* stream);
/* d2d.8: Conv2d(C->1, 3x3, pad 1) + bias.  y bf16 [B,H,W,C]; w = weight_orig f32 [C,9]; out f32 [B,H,W]. */
int p2i_d2d_last_fwd(const void* y, const float* w, const float* sigma, const float* bias, float* out, int B, int H, int W,
                     int C, void* stream);
/* d3d.8 (1x1x1, C->1) + mean over T + bilinear resize (align_corners=False) to [H2,W2] + sigmoid(alpha)*out2d fusion.
 * z bf16 [B,T,h,w,C]; m_scratch f32 [B,h,w]; fused f32 [B,H2,W2]. */
int p2i_disc_tail_fwd(const void* z, const float* w3, const float* sigma3, const float* b3, const float* out2d,
                      const float* alpha, float* m_scratch, float* fused, int B, int T, int h, int w, int C, int H2, int W2,
                      void* stream);

/* ---- discriminator backward ---- */
/* Tail backward.  dfused f32 [B,H2,W2] -> d_out2d f32 [B,H2,W2] (overwritten), dalpha (accumulated, may be NULL),
 * dpre bf16 [B,T,h,w,C] = gradient w.r.t. the d3d.6 pre-activation (LeakyReLU mask of z applied), dW3 f32 [C] and
 * db3 f32 [1] accumulated w.r.t. the NORMALISED 1x1x1 weight (may be NULL).  Requires H2 == 2h or H2 == h. */
int p2i_disc_tail_bwd(const float* dfused, const float* out2d, const float* alpha, const void* z, const float* w3,
                      const float* sigma3, float* d_out2d, float* dalpha, void* dpre, float* dW3, float* db3, int B, int T,
                      int h, int w, int C, int H2, int W2, void* stream);
/* d2d.8 backward: dpre bf16 [B,H,W,C] (LeakyReLU mask of y4 applied); dW f32 [C,9], db f32 [1] accumulated (may be NULL). */
int p2i_d2d_last_bwd(const float* d_out, const void* y4, const float* w, const float* sigma, void* dpre, float* dW, float* db,
                     int B, int H, int W, int C, void* stream);
/* out[c] += sum_rows g[row][c]   (bias gradients; g bf16 [rows, C]). */
int p2i_colsum_bf16(const void* g, float* out, long long rows, int C, void* stream);
/* d3d.0 backward: dpre bf16 [B,T,H/2,W/2,32]; dW f32 [32,27] / db f32 [32] accumulated (both NULL to skip);
 * dx f32 [B,T,H,W] overwritten (NULL to skip). */
int p2i_d3d_first_bwd(const void* dpre, const float* x, const float* w, const float* sigma, float* dW, float* db, float* dx,
                      int B, int T, int H, int W, void* stream);
/* dx[b,c,y,x] += g[b,y,x,c], c < C: adds the 2-D branch's (64-channel padded) input gradient to dx f32 [B,C,H,W]. */
int p2i_disc_unpack_input_grad(const void* g, float* dx, int B, int C, int H, int W, void* stream);

typedef struct P2iSnGrad {
    const float* G;     /* dL/dW_sn: packed f32 [tap'][Cout][Cin'] (wgrad output) or plain [Cout][Cin*kt*k*k] */
    const float* W;     /* weight_orig                                                                       */
    const float* u;     /* u, v, sigma as used by the forward being differentiated                            */
    const float* v;
    const float* sigma;
    float* dW;          /* gradient w.r.t. weight_orig, PyTorch layout (accumulated: +=)                      */
    int Cout, Cin, KT, ksize;
    int s2, cin_pad, packed, _pad;
} P2iSnGrad;
/* dW_orig = G/sigma - (<G, W_orig>/sigma^2) u v^T for a table of layers (spectral_norm backward with u, v detached).
 * inner_scratch: f32 [n_layers], zero-filled by the caller. */
int p2i_spectral_norm_bwd(const P2iSnGrad* table_dev, int n_layers, float* inner_scratch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Losses  (p2igan_bench/modules/losses.py:38-85, 192-253)
 * ------------------------------------------------------------------------------------------- */

/* ReconstructionLoss forward.  pred/target f32 [B,T,HW].  sums[0] += sum w(y)|p-y| (weighted L1 numerator),
 * sums[1] += sum over the B*(T-1) temporal-difference rows of KL(q||p) with p,q = softmax(diff/temperature);
 * lse f32 [B*(T-1),2] receives the per-row log-sum-exp (kept for the backward).  Caller zero-fills sums.
 * loss = sums[0]/(B*T*HW) + k1_alpha * sums[1]/B. */
int p2i_rec_loss_fwd(const float* pred, const float* target, int B, int T, int HW, float temperature, float* sums,
                     float* lse, void* stream);
/* d loss / d pred (closed form), scaled by *gscale (device scalar, NULL = 1). dpred f32 [B,T,HW] overwritten. */
int p2i_rec_loss_bwd(const float* pred, const float* target, const float* lse, const float* gscale, float k1_alpha,
                     float temperature, float* dpred, int B, int T, int HW, void* stream);
/* AdversarialLoss: *out += mean(term(x)).  mode 0: relu(1-x) (hinge, D real) | 1: relu(1+x) (hinge, D fake) |
 * 2: -x (hinge, G) | 3: BCE(x, label) on raw outputs (nsgan, losses.py:203) | 4: (x-label)^2 (lsgan). */
int p2i_gan_loss_fwd(const float* logits, long long n, int mode, float label, float* out, void* stream);
/* dlogits = (*gscale) * scale * d mean(term) / dx. */
int p2i_gan_loss_bwd(const float* logits, long long n, int mode, float label, const float* gscale, float scale,
                     float* dlogits, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Metrics  (p2igan_bench/metrics/metric.py:16-183)
 * ------------------------------------------------------------------------------------------- */

/* One RainfallMetricSuite.update(): pred/target f32 [N,H,W] (N = B*T frames).  thresholds (host, <= 4) are rain-rate
 * thresholds applied to R = 0.036*10^(x/16); scales (host, <= 4, each in {1,2,4,8}) are FSS box sizes.
 * scratch: double [50], zero-filled once by the caller (left clean after every call).
 * state: float32 [51] streaming state, layout: [0] abs_sum [1] squared_sum [2] n_obs | [3+4t+k] hits, misses, false
 * alarms, correct negatives of threshold t | [19+4t+s] FSS score_sum | [35+4t+s] FSS counts (per update call, as in
 * metric.py:168-169).  apply_transform selects whether MAE/RMSE use the transformed values (metric.py:46-48). */
int p2i_metrics_update(const float* pred, const float* target, int N, int H, int W, const float* thresholds, int n_thr,
                       const int* scales, int n_scale, int apply_transform, double* scratch, float* state, void* stream);

/* FractionalSkillScoreMetric.update (metric.py:151-169) for ONE (threshold, box size) pair with an arbitrary box size
 * 1 <= scale <= 32 (the fused pass above covers the default boxes {1,2,4,8}; MetricConfig.scales takes any, :186-191):
 * avg_pool2d(mask, scale, stride 1, pad scale/2) with the zero padding counted, (H + 2*(scale/2) - scale + 1)^2 outputs.
 * scratch: double [2] zero-filled once (left clean); state f32 [2] += {1 - mean_num/(mean_den + 1e-10), 1}. */
int p2i_fss_update(const float* pred, const float* target, int N, int H, int W, float threshold, int scale, double* scratch,
                   float* state, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimiser  (torch.optim.Adam as used by scripts/train.py:125-136)
 * ------------------------------------------------------------------------------------------- */
typedef struct P2iAdamTensor {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    long long n;
} P2iAdamTensor;
/* One launch for a whole model.  chunks_dev: int pairs (tensor index, chunk index), chunk = p2i_adam_chunk_elems()
 * elements.  step_dev points to THREE device floats {step, lr/(1-beta1^step), 1/sqrt(1-beta2^step)}: step is
 * incremented first and the two bias-correction factors are refreshed from it by a one-thread kernel (double
 * precision, as torch does on the host), so the call is CUDA-graph capturable and the update kernel's blocks only
 * read them.  grad_scale multiplies every gradient (1/world_size when the gradients hold a SUM over ranks). */
int p2i_adam_step(const P2iAdamTensor* tensors_dev, const int* chunks_dev, int n_chunks, float* step_dev, float lr,
                  float beta1, float beta2, float eps, float grad_scale, void* stream);
/* The two halves of p2i_adam_step for a BUCKETED update (the parameters of a gradient bucket are updated as soon as the
 * bucket is final, under the rest of the backward pass): p2i_adam_tick once per optimiser step (step += 1, bias corrections),
 * then p2i_adam_apply per bucket with that bucket's tensor / chunk tables.  Every parameter must be applied exactly once. */
int p2i_adam_tick(float* step_dev, float lr, float beta1, float beta2, void* stream);
int p2i_adam_apply(const P2iAdamTensor* tensors_dev, const int* chunks_dev, int n_chunks, const float* step_dev, float lr,
                   float beta1, float beta2, float eps, float grad_scale, void* stream);
int p2i_adam_chunk_elems(void);

/* SSIM of RegressionMetrics (metric.py:36,55-56,69 -> torchmetrics StructuralSimilarityIndexMeasure(data_range):
 * 11x11 gaussian window, sigma 1.5, k1 0.01, k2 0.03, per-image mean over the (H-10)x(W-10) interior, mean over images).
 * PARITY UNPINNED: torchmetrics 1.0.3 is absent from the reference tree and this image; restated from its published
 * algorithm.  pred/target f32 [N,H,W]; state f64 [2] += {sum of per-image SSIM, N}. */
int p2i_ssim_update(const float* pred, const float* target, int N, int H, int W, int apply_transform, float data_range,
                    double* state, void* stream);

/* Sliding-window blend of scripts/infer.py:237-245: preds f32 [n_win, stride, HW] (window w starts at frame w*step) ->
 * out f32 [L, HW] = clip(scale * mean over covering windows, 0). */
int p2i_window_blend(const float* preds, float* out, int L, int HW, int stride, int step, int n_win, float scale,
                     void* stream);

/* Device-side batch preparation (STIDataset.post_process, p2igan_bench/data/sti_dataset.py:203-229, and
 * Trainer._prepare_batch, scripts/train.py:468-473): frames_u8 [B,T,H0,W0] uint8 -> frames = u8/255 (IEEE fp32
 * division, as numpy), masked = frames * mask, masks in {0,1}; all three f32 [B,T,1,H,W], centre-cropped from
 * (H0,W0) to (H,W) with start (old-new)//2.  mask_u8: non-zero = observed; mask_mode 0: one [H0,W0] pattern
 * ('sti'/'stis' file masks), 1: [B,H0,W0], 2: [B,T,H0,W0]. */
int p2i_batch_prep_u8(const void* frames_u8, const void* mask_u8, float* frames, float* masked, float* masks, int B, int T,
                      int H0, int W0, int H, int W, int mask_mode, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange over NVLink peer memory (replaces the bucketed NCCL all-reduce of
 * torch DDP, which the reference would use -- it has no multi-GPU path of its own).
 * Set-up (the only entry points of this library that allocate or touch other processes): each rank allocates its flat
 * gradient buffer and a flag array with p2i_peer_alloc (cudaMalloc + zero fill), exports both with p2i_peer_export (64-byte
 * CUDA IPC handle), exchanges the handles out of band (torch.distributed.all_gather_object) and maps the peers' with
 * p2i_peer_import.  p2i_peer_allreduce then sums the N buffers in place on every rank with ONE kernel per rank
 * (two-shot: reduce-scatter + all-gather by direct peer loads, block-level flag barriers, fixed summation order ->
 * bitwise identical results on all ranks).  bufs / flags: HOST arrays of `world` device pointers as mapped in this
 * process (index = rank; the own entries are the local allocations).  epoch_dev / err_dev: local device ints (zero
 * initialised); *err_dev becomes 1 if a peer did not show up within ~10 s (the kernel then finishes instead of hanging).
 * CUDA-graph capturable; all ranks must call it the same number of times.
 * ------------------------------------------------------------------------------------------- */
int p2i_peer_alloc(void** out, long long bytes);
int p2i_peer_free(void* p);
int p2i_peer_export(const void* p, void* handle64);
int p2i_peer_import(const void* handle64, void** out);
int p2i_peer_close(void* p);
int p2i_peer_flags_bytes(void);
int p2i_peer_allreduce(void* const* bufs, void* const* flags, int rank, int world, long long n, int* epoch_dev, int* err_dev,
                       void* stream);
/* The same exchange restricted to elements [offset, offset + n) of the flat buffers (both multiples of 4) with `blocks`
 * CTAs (0 = one per SM): the bucketed gradient exchange that runs on a side stream UNDER the backward pass -- a bucket is
 * exchanged as soon as its gradients are final, with few enough CTAs to co-reside with the tensor-core kernels
 * (what torch DDP's bucketed NCCL all-reduce would do).  Calls on one set of flags must be serialised
 * (one stream) and issued in the same order on every rank. */
int p2i_peer_allreduce_range(void* const* bufs, void* const* flags, int rank, int world, long long offset, long long n,
                             int blocks, int* epoch_dev, int* err_dev, void* stream);

/* Layout helpers for the per-layer drop-in modules: NCHW f32 <-> NHWC bf16. */
int p2i_nchw_f32_to_nhwc_bf16(const float* x, void* y, int B, int C, int H, int W, void* stream);
int p2i_nhwc_bf16_to_nchw_f32(const void* x, float* y, int B, int C, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* P2I_B200_H */
