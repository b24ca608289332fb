// Device-side batch preparation: what STIDataset.post_process does on the host per sample
// (p2igan_bench/data/sti_dataset.py:203-229) and Trainer._prepare_batch does per batch (scripts/train.py:468-473):
//   frames = uint8 / 255 (float32) ; masked = frames * mask ; centre crop to (H, W) ; layout [B, T, 1, H, W].
// The DataLoader can then ship uint8 frames (1 byte/pixel) plus one mask instead of three float32 tensors
// (12 bytes/pixel): 12x less host->device traffic.  HBM-bound: 1-2 bytes read, 12 bytes written per pixel.
#include "common.h"
#include "ptx.cuh"

namespace p2i {

// mask_mode: 0 = one [H0, W0] pattern for every frame and sample ('stis' / 'sti' file masks),
//            1 = per sample [B, H0, W0], 2 = per frame [B, T, H0, W0]
__global__ void __launch_bounds__(256) batch_prep_u8_kernel(const uint8_t* __restrict__ src, const uint8_t* __restrict__ mask,
                                                            float* __restrict__ frames, float* __restrict__ masked,
                                                            float* __restrict__ masks, int BT, int T, int H0, int W0, int H, int W,
                                                            int y0, int x0, int mask_mode) {
    const int wq = W >> 2;                                   // 4 output pixels per thread (W % 4 == 0)
    const long long total = static_cast<long long>(BT) * H * wq;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int xq = static_cast<int>(i % wq);
        const long long r = i / wq;
        const int y = static_cast<int>(r % H);
        const int bt = static_cast<int>(r / H);
        const size_t so = (static_cast<size_t>(bt) * H0 + y0 + y) * W0 + x0 + 4 * xq;
        const int mi = mask_mode == 0 ? 0 : (mask_mode == 1 ? bt / T : bt);
        const size_t mo = (static_cast<size_t>(mi) * H0 + y0 + y) * W0 + x0 + 4 * xq;
        float f[4], m[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            f[k] = __fdiv_rn(static_cast<float>(src[so + k]), 255.f);      // numpy: astype(float32) / 255.0
            m[k] = mask[mo + k] ? 1.f : 0.f;
        }
        const size_t o = (static_cast<size_t>(bt) * H + y) * W + 4 * xq;
        *reinterpret_cast<float4*>(frames + o) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(masked + o) = make_float4(f[0] * m[0], f[1] * m[1], f[2] * m[2], f[3] * m[3]);
        *reinterpret_cast<float4*>(masks + o) = make_float4(m[0], m[1], m[2], m[3]);
    }
}

}  // namespace p2i

extern "C" int p2i_batch_prep_u8(const void* frames_u8, const void* mask_u8, float* frames, float* masked, float* masks, int B,
                                 int T, int H0, int W0, int H, int W, int mask_mode, void* stream) {
    P2I_CHECK_ARG(frames_u8 && mask_u8 && frames && masked && masks, "batch_prep_u8: null pointer");
    P2I_CHECK_ARG(B > 0 && T > 0 && H > 0 && W > 0 && H <= H0 && W <= W0, "batch_prep_u8: bad shape (crop must fit the source)");
    P2I_CHECK_ARG(W % 4 == 0, "batch_prep_u8: W must be a multiple of 4");
    P2I_CHECK_ARG(mask_mode >= 0 && mask_mode <= 2, "batch_prep_u8: mask_mode 0 (one pattern) | 1 (per sample) | 2 (per frame)");
    // STIDataset._crop_center: start = max((old - new) // 2, 0)
    const int y0 = (H0 - H) / 2, x0 = (W0 - W) / 2;
    const long long total = static_cast<long long>(B) * T * H * (W / 4);
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(p2i::sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    p2i::batch_prep_u8_kernel<<<static_cast<unsigned>(blocks), 256, 0, p2i::as_stream(stream)>>>(
        static_cast<const uint8_t*>(frames_u8), static_cast<const uint8_t*>(mask_u8), frames, masked, masks, B * T, T, H0, W0, H, W, y0, x0,
        mask_mode);
    P2I_CHECK_LAUNCH("batch_prep_u8_kernel");
    return P2I_OK;
}
