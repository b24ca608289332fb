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
                    int* nbr_idx, float* nbr_w, int B, int T, int H, int W, float tau, int search, void* stream);
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
    float* dW;           /* [C, C, 9] (overwritten) */
    float* dD;           /* [C, 9, 9] (overwritten) */
    int channels;
    int _pad;
} P2iDoGrad;

/* Backward of the composition for a table of layers (torch.einsum backward at deconv_pytorch.py:124):
 * dW[o,i,s] = sum_m dDoW[m,o,i] (D+D_diag)[i,m,s] ;  dD[i,m,s] = sum_o dDoW[m,o,i] W[o,i,s]. */
int p2i_doconv_compose_bwd(const P2iDoGrad* table_dev, int n_layers, int max_channels, void* stream);

/* Grouped stem variant (Convsin: 16->64, k3, groups 4) keeping the reference's raw-reshape row
 * pairing (deconv_pytorch.py:119-124).  out f32 [64,4,9]. */
int p2i_doconv_compose_stem_fwd(const float* W, const float* D, const float* D_diag, float* out, void* stream);
/* dDoW f32 [64,4,9] -> dW [64,4,9], dD [16,9,9]. */
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
/* dy [B,H,W,64] bf16, x f32 [B,16,H,W], w f32 [64,4,9] -> dx f32 [B,16,H,W] (overwritten), dw f32 [64,4,9] accumulated. */
int p2i_stem_bwd(const void* dy, const float* x, const float* w, float* dx, float* dw, int B, int H, int W, void* stream);

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

/* Layout helpers for the per-layer drop-in modules: NCHW f32 <-> NHWC bf16. */
int p2i_nchw_f32_to_nhwc_bf16(const float* x, void* y, int B, int C, int H, int W, void* stream);
int p2i_nhwc_bf16_to_nchw_f32(const void* x, float* y, int B, int C, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* P2I_B200_H */
