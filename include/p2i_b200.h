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
/* Grouped stem variant (Convsin: 16->64, k3, groups 4) keeping the reference's raw-reshape row
 * pairing (deconv_pytorch.py:119-124).  out f32 [64,4,9]. */
int p2i_doconv_compose_stem_fwd(const float* W, const float* D, const float* D_diag, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Convolutions as tcgen05 implicit GEMM  (F.conv2d at deconv_pytorch.py:108; layer.py:129-135)
 * ------------------------------------------------------------------------------------------- */

/* y = act(conv(x, w) [+ residual]);  x [B,H,W,Cin] bf16, w [k*k][Cout][Cin] bf16, y [B,H,W,Cout] bf16.
 * ksize in {1,3}, stride 1, zero padding k/2.  Cin % 64 == 0, Cout % 64 == 0.
 * flags: bit0 = ReLU after the residual add.  residual may be NULL. */
#define P2I_CONV_RELU 1
int p2i_conv2d_igemm_fwd(const void* x, const void* w, const void* residual, void* y, int B, int H, int W, int Cin,
                         int Cout, int ksize, int flags, void* stream);
/* CUDA-core direct convolution with identical semantics (test/triage only; never on the product path). */
int p2i_conv2d_direct_fwd(const void* x, const void* w, const void* residual, void* y, int B, int H, int W, int Cin,
                          int Cout, int ksize, int flags, void* stream);

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

/* Layout helpers for the per-layer drop-in modules: NCHW f32 <-> NHWC bf16. */
int p2i_nchw_f32_to_nhwc_bf16(const float* x, void* y, int B, int C, int H, int W, void* stream);
int p2i_nhwc_bf16_to_nchw_f32(const void* x, float* y, int B, int C, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* P2I_B200_H */
