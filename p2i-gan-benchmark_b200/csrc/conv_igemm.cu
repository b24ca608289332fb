// Convolution as a tcgen05 implicit GEMM for sm_100a: spatial stride 1, taps k in {1,2,3} (x KT in {1,3} frames).
// FIRST GENERATION: since conv_igemm_halo.cu exists this kernel only takes the shapes that one declines (channel counts
// that are not multiples of 64, the thin direct-conv path, p2i_set_conv_impl(1) for A/B runs); the dispatch lives in run_igemm.
//
//   reference call sites
//     F.conv2d in DOConv2d._conv_forward (p2igan_bench/modules/deconv_pytorch.py:104-109) with the ReLU /
//       residual of BasicConv_do / ResBlock_do fused (p2igan_bench/modules/layer.py:84-94,134-135);
//     the 1x1 projection of UPPos (layer.py:390,398);
//     the spectral-norm Conv2d / Conv3d + LeakyReLU stacks of P2IDiscriminator (models/p2igan.py:120-142):
//       their stride-2 convs run here as stride-1 k=2 convs on a space-to-depth input (out_mode 1 writes that
//       layout for the next layer; out_mode 2 undoes it in the data-gradient pass), temporal taps are extra
//       TMA coordinates (5-D tensor map [C, W, H, T, B]).
//   The data gradient of every conv is this same kernel on dY with transposed, tap-flipped weights; the ReLU /
//   LeakyReLU backward is fused as an output mask.
//
// GEMM view: M = frames*H*W output pixels, N = Cout, K = KT*k*k*Cin.
//   A (activations) : NHWC bf16. One CTA tile = Ht x Wt = 128 output pixels of one frame. For every temporal tap,
//       64-channel block and horizontal tap kx ONE TMA box of (Ht+k-1) x Wt pixels x 64 channels is loaded
//       (out-of-image pixels / frames are zero-filled by the TMA unit = the conv's zero padding). The k vertical
//       taps reuse that box: tap ky starts ky*Wt rows further down, a multiple of 1024 B because Wt % 8 == 0, so
//       every shifted view is still a valid 128B-swizzled K-major UMMA operand. A is fetched k (not k*k) times.
//   B (weights)     : bf16 [tap][Cout][Cin] (K-major), one TMA box of NT x 64 per (tap, channel block).
//   D (accumulator) : fp32 in TMEM, 128 lanes x NT columns, double buffered (2*NT <= 512 columns) so the
//       epilogue of tile i overlaps the MMAs of tile i+1.
// Warp roles (256 threads, persistent over tiles): warp0 = TMA producer, warp1 = MMA issuer (one elected
// thread), warp2 = TMEM allocator, warps 4-7 = epilogue (TMEM -> regs -> bias/residual/act/mask -> bf16 -> global).
#include "common.h"
#include "ptx.cuh"

namespace p2i {

struct ConvParams {
    int F;                 // output frames = samples * T_out
    int T_out, T_in;       // frames per sample (1, 1 for 2-D)
    int H, W, Cin, Cout;
    int KT, KH, KW;
    int pad, pad_t, st;    // x_in = x + kx - pad ; t_in = st*t_out + kt - pad_t  (tmode 0)
    int tmode;             // 1: transposed in t (dgrad of a temporal stride-st conv): t_in = (t_out + pad_t - kt)/st if divisible
    int Ht, Wt, tiles_x, tiles_y, m_tiles, total_tiles;
    int act;               // 0 none, 1 ReLU, 2 LeakyReLU(0.2)
    int mask_mode;         // 0 none, 1: zero where mask <= 0, 2: x0.2 where mask <= 0
    int out_mode;          // 0 natural [F,H,W,Cout]; 1 space-to-depth pack -> [F,H/2,W/2,4*Cout]; 2 unpack -> [F,2H,2W,Cout/4]
    const __nv_bfloat16* residual;   // natural layout, added before the activation
    const __nv_bfloat16* mask;       // natural layout
    const float* bias;               // fp32 [Cout]
    __nv_bfloat16* y;
};

template <int NT>
struct ConvCfg {
    static constexpr int SA = 3;
    static constexpr int SB = (NT == 256) ? 4 : (NT == 128 ? 6 : 8);
    static constexpr int A_STAGE = 20480;  // max over tile shapes: (8+2)*16*128 B
    static constexpr int B_STAGE = NT * 128;
    static constexpr int NBAR = 2 * SA + 2 * SB + 4;
    static constexpr int SMEM = 1024 + SA * A_STAGE + SB * B_STAGE + NBAR * 8 + 16;
    static constexpr int TMEM_COLS = 2 * NT;
};

// source frame of temporal tap kt for output frame t_out; -1 = tap skipped (transposed mode, parity mismatch)
__device__ __forceinline__ int src_frame(const ConvParams& p, int t_out, int kt, bool& skip) {
    skip = false;
    if (p.tmode == 0) return p.st * t_out + kt - p.pad_t;
    const int s = t_out + p.pad_t - kt;
    if (s % p.st != 0) { skip = true; return 0; }
    return s / p.st;   // negative / too large values are zero-filled by TMA (s < 0 with st == 2: s/st rounds to 0 only for s == -1, which is odd -> skipped)
}

template <int NT>
__global__ void __launch_bounds__(256, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const ConvParams p) {
    using Cfg = ConvCfg<NT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = sA + Cfg::SA * Cfg::A_STAGE;
    uint64_t* fullA = reinterpret_cast<uint64_t*>(sB + Cfg::SB * Cfg::B_STAGE);
    uint64_t* emptyA = fullA + Cfg::SA;
    uint64_t* fullB = emptyA + Cfg::SA;
    uint64_t* emptyB = fullB + Cfg::SB;
    uint64_t* tfull = emptyB + Cfg::SB;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::SA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < Cfg::SB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int cblocks = p.Cin >> 6;
    const uint32_t a_bytes = static_cast<uint32_t>((p.Ht + p.KH - 1) * p.Wt * 128);
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        if (elect_one()) {
            // ---------------------------------------------------------------- TMA producer
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int nt = tile / p.m_tiles, mt = tile - nt * p.m_tiles;
                const int f = mt / tiles_per_img, r = mt - f * tiles_per_img;
                const int smp = f / p.T_out, t_out = f - smp * p.T_out;
                const int y0 = (r / p.tiles_x) * p.Ht, x0 = (r % p.tiles_x) * p.Wt;
                for (int kt = 0; kt < p.KT; ++kt) {
                    bool skip;
                    const int t_in = src_frame(p, t_out, kt, skip);
                    if (skip) continue;
                    for (int cb = 0; cb < cblocks; ++cb) {
                        for (int kx = 0; kx < p.KW; ++kx) {
                            mbar_wait(&emptyA[sa], pa ^ 1);
                            mbar_expect_tx(&fullA[sa], a_bytes);
                            tma_load_5d(sA + sa * Cfg::A_STAGE, &tmA, &fullA[sa], cb * 64, x0 + kx - p.pad, y0 - p.pad, t_in, smp);
                            if (++sa == Cfg::SA) { sa = 0; pa ^= 1; }
                            for (int ky = 0; ky < p.KH; ++ky) {
                                mbar_wait(&emptyB[sb], pb ^ 1);
                                mbar_expect_tx(&fullB[sb], Cfg::B_STAGE);
                                tma_load_3d(sB + sb * Cfg::B_STAGE, &tmB, &fullB[sb], cb * 64, nt * NT,
                                            (kt * p.KH + ky) * p.KW + kx);
                                if (++sb == Cfg::SB) { sb = 0; pb ^= 1; }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ---------------------------------------------------------------- MMA issuer
            constexpr uint32_t idesc = make_idesc_bf16(128, NT);
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const int mt = tile % p.m_tiles;
                const int t_out = (mt / tiles_per_img) % p.T_out;
                mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * NT;
                uint32_t acc = 0;
                for (int kt = 0; kt < p.KT; ++kt) {
                    bool skip;
                    (void)src_frame(p, t_out, kt, skip);
                    if (skip) continue;
                    for (int cb = 0; cb < cblocks; ++cb) {
                        for (int kx = 0; kx < p.KW; ++kx) {
                            mbar_wait(&fullA[sa], pa);
                            const uint32_t a_base = smem_u32(sA + sa * Cfg::A_STAGE);
                            for (int ky = 0; ky < p.KH; ++ky) {
                                mbar_wait(&fullB[sb], pb);
                                tc_fence_after();
                                const uint32_t a_addr = a_base + ky * p.Wt * 128;
                                const uint32_t b_addr = smem_u32(sB + sb * Cfg::B_STAGE);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    umma_bf16(d_tmem, make_sw128_desc(a_addr + k * 32, 16, 1024),
                                              make_sw128_desc(b_addr + k * 32, 16, 1024), idesc, acc);
                                    acc = 1;
                                }
                                umma_commit(&emptyB[sb]);
                                if (++sb == Cfg::SB) { sb = 0; pb ^= 1; }
                            }
                            umma_commit(&emptyA[sa]);
                            if (++sa == Cfg::SA) { sa = 0; pa ^= 1; }
                        }
                    }
                }
                umma_commit(&tfull[as]);
            }
        }
    } else if (warp >= 4) {
        // -------------------------------------------------------------------- epilogue
        const int ew = warp - 4;
        const int row = ew * 32 + lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const int nt = tile / p.m_tiles, mt = tile - nt * p.m_tiles;
            const int f = mt / tiles_per_img, r = mt - f * tiles_per_img;
            const int y = (r / p.tiles_x) * p.Ht + row / p.Wt;
            const int x = (r % p.tiles_x) * p.Wt + row % p.Wt;
            const bool valid = (y < p.H) && (x < p.W);
            const size_t off = ((static_cast<size_t>(f) * p.H + y) * p.W + x) * p.Cout + nt * NT;   // natural
            mbar_wait(&tfull[as], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + as * NT + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < NT; c += 16) {
                uint32_t v[16];
                tmem_ld16(t_addr + c, v);
                tmem_ld_wait();
                if (valid) {
                    float fv[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) fv[i] = __uint_as_float(v[i]);
                    if (p.bias != nullptr) {
                        const float4* bp = reinterpret_cast<const float4*>(p.bias + nt * NT + c);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 bv = __ldg(bp + i);
                            fv[4 * i] += bv.x; fv[4 * i + 1] += bv.y; fv[4 * i + 2] += bv.z; fv[4 * i + 3] += bv.w;
                        }
                    }
                    if (p.residual != nullptr) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off + c);
                        uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float2 t = unpack_bf16x2(rr[i]);
                            fv[2 * i] += t.x;
                            fv[2 * i + 1] += t.y;
                        }
                    }
                    if (p.act == 1) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) fv[i] = fmaxf(fv[i], 0.f);
                    } else if (p.act == 2) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) fv[i] = fv[i] > 0.f ? fv[i] : 0.2f * fv[i];
                    }
                    if (p.mask_mode != 0) {
                        const uint4* mp = reinterpret_cast<const uint4*>(p.mask + off + c);
                        uint4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
                        const uint32_t mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
                        const float neg = (p.mask_mode == 2) ? 0.2f : 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float2 t = unpack_bf16x2(mm[i]);
                            if (!(t.x > 0.f)) fv[2 * i] *= neg;
                            if (!(t.y > 0.f)) fv[2 * i + 1] *= neg;
                        }
                    }
                    uint4 o0, o1;
                    o0.x = pack_bf16x2(fv[0], fv[1]);   o0.y = pack_bf16x2(fv[2], fv[3]);
                    o0.z = pack_bf16x2(fv[4], fv[5]);   o0.w = pack_bf16x2(fv[6], fv[7]);
                    o1.x = pack_bf16x2(fv[8], fv[9]);   o1.y = pack_bf16x2(fv[10], fv[11]);
                    o1.z = pack_bf16x2(fv[12], fv[13]); o1.w = pack_bf16x2(fv[14], fv[15]);
                    size_t o = off + c;
                    if (p.out_mode == 1) {          // pack: pixel (y,x), channel n -> [y/2][x/2][(y&1)*2+(x&1)][n]
                        const int n = nt * NT + c;
                        o = (((static_cast<size_t>(f) * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1)) * 4 + ((y & 1) * 2 + (x & 1))) *
                                p.Cout + n;
                    } else if (p.out_mode == 2) {   // unpack: channel n = q*C + ch -> pixel (2y + q/2, 2x + q%2), channel ch
                        const int C = p.Cout >> 2;
                        const int n = nt * NT + c, q = n / C, ch = n - q * C;
                        o = ((static_cast<size_t>(f) * (2 * p.H) + 2 * y + (q >> 1)) * (2 * p.W) + 2 * x + (q & 1)) * C + ch;
                    }
                    uint4* op = reinterpret_cast<uint4*>(p.y + o);
                    op[0] = o0;
                    op[1] = o1;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int NT>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t st) {
    using Cfg = ConvCfg<NT>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::SMEM);
        if (e != cudaSuccess) return fail(P2I_ERR_CUDA, "conv_igemm smem attribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
    conv_igemm_kernel<NT><<<grid, 256, Cfg::SMEM, st>>>(tmA, tmB, p);
    P2I_CHECK_LAUNCH("conv_igemm_kernel");
    return P2I_OK;
}

// --------------------------------------------------------------------------------------------
// CUDA-core direct convolution (2-D, natural layout): used by tests to triage the tensor-core kernel.
// --------------------------------------------------------------------------------------------
__global__ void conv_direct_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                   const __nv_bfloat16* __restrict__ res, const __nv_bfloat16* __restrict__ mask,
                                   const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int B, int H,
                                   int W, int Cin, int Cout, int K, int relu) {
    const long long total = static_cast<long long>(B) * H * W * Cout;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i % Cout);
        long long pix = i / Cout;
        const int xx = static_cast<int>(pix % W);
        const int yy = static_cast<int>((pix / W) % H);
        const int b = static_cast<int>(pix / (static_cast<long long>(W) * H));
        float acc = 0.f;
        for (int ky = 0; ky < K; ++ky) {
            const int iy = yy + ky - K / 2;
            if (iy < 0 || iy >= H) continue;
            for (int kx = 0; kx < K; ++kx) {
                const int ix = xx + kx - K / 2;
                if (ix < 0 || ix >= W) continue;
                const __nv_bfloat16* xp = x + ((static_cast<size_t>(b) * H + iy) * W + ix) * Cin;
                const __nv_bfloat16* wp = w + (static_cast<size_t>(ky * K + kx) * Cout + co) * Cin;
                for (int ci = 0; ci < Cin; ++ci) acc += __bfloat162float(xp[ci]) * __bfloat162float(wp[ci]);
            }
        }
        if (bias) acc += bias[co];
        if (res) acc += __bfloat162float(res[i]);
        if (relu == 1) acc = fmaxf(acc, 0.f);
        else if (relu == 2) acc = acc > 0.f ? acc : 0.2f * acc;
        if (mask && !(__bfloat162float(mask[i]) > 0.f)) acc = 0.f;
        y[i] = __float2bfloat16(acc);
    }
}

// conv_igemm_halo.cu: second-generation kernel (single halo box per tile, resident / M-blocked weights, TMA-store
// epilogue).  Returns +1 when the shape is not eligible.
int run_igemm_halo(const void* x, const void* w, const P2iConvDesc& d, const void* residual, const void* mask,
                   const float* bias, void* y, void* stream);

int set_halo_cg(int cg);
static std::atomic<int> g_conv_impl{0};   // 0 auto (halo when eligible), 1 legacy only, 2 halo only (3: single CTA, 4: CTA pairs)

static int run_igemm(const void* x, const void* w, const P2iConvDesc& d, const void* residual, const void* mask,
                     const float* bias, void* y, void* stream) {
    P2I_CHECK_ARG(x && w && y, "conv_igemm: null pointer");
    P2I_CHECK_ARG(d.ksize >= 1 && d.ksize <= 3, "conv_igemm: ksize %d unsupported (1..3)", d.ksize);
    P2I_CHECK_ARG(d.kt == 1 || d.kt == 3, "conv_igemm: kt %d unsupported (1 or 3)", d.kt);
    P2I_CHECK_ARG(d.samples > 0 && d.H > 0 && d.W > 0 && d.T_in > 0 && d.T_out > 0, "conv_igemm: bad shape");
    P2I_CHECK_ARG(d.Cin % 64 == 0 && d.Cout % 64 == 0 && d.Cin > 0 && d.Cout > 0,
                  "conv_igemm: Cin=%d Cout=%d must be positive multiples of 64", d.Cin, d.Cout);
    P2I_CHECK_ARG(d.stride_t == 1 || d.stride_t == 2, "conv_igemm: temporal stride %d unsupported", d.stride_t);
    P2I_CHECK_ARG(d.out_mode >= 0 && d.out_mode <= 2 && d.mask_mode >= 0 && d.mask_mode <= 2, "conv_igemm: bad mode");
    P2I_CHECK_ARG(d.mask_mode == 0 || mask, "conv_igemm: mask_mode set without a mask");
    P2I_CHECK_ARG(d.out_mode != 1 || (d.H % 2 == 0 && d.W % 2 == 0), "conv_igemm: s2d pack needs even H, W");
    P2I_CHECK_ARG(d.out_mode != 2 || (d.Cout % 256 == 0 || (d.Cout / 4) % 16 == 0), "conv_igemm: s2d unpack needs Cout/4 %% 16 == 0");
    const int impl = g_conv_impl.load(std::memory_order_relaxed);
    if (impl != 1) {
        const int rc = run_igemm_halo(x, w, d, residual, mask, bias, y, stream);
        if (rc <= 0) return rc;
        if (impl == 2) return fail(P2I_ERR_INVALID, "conv_igemm: shape not eligible for the halo kernel");
    }
    ConvParams p;
    p.F = d.samples * d.T_out; p.T_out = d.T_out; p.T_in = d.T_in;
    p.H = d.H; p.W = d.W; p.Cin = d.Cin; p.Cout = d.Cout;
    p.KT = d.kt; p.KH = d.ksize; p.KW = d.ksize;
    p.pad = d.pad; p.pad_t = d.pad_t; p.st = d.stride_t; p.tmode = d.t_transposed;
    p.Wt = (d.W >= 16) ? 16 : 8;
    p.Ht = 128 / p.Wt;
    p.tiles_x = cdiv(d.W, p.Wt);
    p.tiles_y = cdiv(d.H, p.Ht);
    p.m_tiles = p.F * p.tiles_x * p.tiles_y;
    p.act = d.act; p.mask_mode = d.mask_mode; p.out_mode = d.out_mode;
    p.residual = static_cast<const __nv_bfloat16*>(residual);
    p.mask = static_cast<const __nv_bfloat16*>(mask);
    p.bias = bias;
    p.y = static_cast<__nv_bfloat16*>(y);
    const int NT = (d.Cout % 256 == 0) ? 256 : ((d.Cout % 128 == 0) ? 128 : 64);
    p.total_tiles = p.m_tiles * (d.Cout / NT);

    CUtensorMap tmA, tmB;
    {
        const uint64_t C = d.Cin, W = d.W, H = d.H, T = d.T_in;
        const uint64_t dims[5] = {C, W, H, T, uint64_t(d.samples)};
        const uint64_t strides[5] = {0, C * 2, W * C * 2, H * W * C * 2, T * H * W * C * 2};
        const uint32_t box[5] = {64, uint32_t(p.Wt), uint32_t(p.Ht + d.ksize - 1), 1, 1};
        int rc = encode_tmap_bf16(&tmA, x, 5, dims, strides, box, nullptr, true);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {uint64_t(d.Cin), uint64_t(d.Cout), uint64_t(d.kt * d.ksize * d.ksize)};
        const uint64_t strides[3] = {0, uint64_t(d.Cin) * 2, uint64_t(d.Cout) * d.Cin * 2};
        const uint32_t box[3] = {64, uint32_t(NT), 1};
        int rc = encode_tmap_bf16(&tmB, w, 3, dims, strides, box, nullptr, true);
        if (rc) return rc;
    }
    cudaStream_t st = as_stream(stream);
    set_last_variant(1000000 + NT * 100);
    if (NT == 256) return launch_conv<256>(tmA, tmB, p, st);
    if (NT == 128) return launch_conv<128>(tmA, tmB, p, st);
    return launch_conv<64>(tmA, tmB, p, st);
}

}  // namespace p2i

using namespace p2i;

extern "C" int p2i_set_conv_impl(int impl) {
    P2I_CHECK_ARG(impl >= 0 && impl <= 4, "set_conv_impl: 0 auto | 1 legacy | 2 halo | 3 halo single-CTA | 4 halo CTA pairs");
    set_halo_cg(impl == 3 ? 1 : (impl == 4 ? 2 : 0));
    g_conv_impl.store(impl >= 2 ? 2 : impl);
    return P2I_OK;
}

extern "C" int p2i_conv_igemm(const void* x, const void* w, const P2iConvDesc* desc, const void* residual, const void* mask,
                              const float* bias, void* y, void* stream) {
    P2I_CHECK_ARG(desc, "conv_igemm: null descriptor");
    return run_igemm(x, w, *desc, residual, mask, bias, y, stream);
}

extern "C" int p2i_conv2d_igemm_fwd(const void* x, const void* w, const void* residual, const void* mask,
                                    const float* bias, void* y, int B, int H, int W, int Cin, int Cout, int ksize,
                                    int flags, void* stream) {
    P2I_CHECK_ARG(ksize == 1 || ksize == 3, "conv2d_igemm: ksize %d unsupported (1 or 3)", ksize);
    P2iConvDesc d;
    d.samples = B; d.T_in = 1; d.T_out = 1; d.H = H; d.W = W; d.Cin = Cin; d.Cout = Cout;
    d.kt = 1; d.ksize = ksize; d.pad = ksize / 2; d.pad_t = 0; d.stride_t = 1; d.t_transposed = 0;
    d.act = (flags & P2I_CONV_RELU) ? 1 : ((flags & P2I_CONV_LEAKY) ? 2 : 0);
    d.mask_mode = mask ? 1 : 0;
    d.out_mode = 0;
    return run_igemm(x, w, d, residual, mask, bias, y, stream);
}

extern "C" int p2i_conv2d_direct_fwd(const void* x, const void* w, const void* residual, const void* mask,
                                     const float* bias, void* y, int B, int H, int W, int Cin, int Cout, int ksize,
                                     int flags, void* stream) {
    P2I_CHECK_ARG(x && w && y, "conv2d_direct: null pointer");
    P2I_CHECK_ARG(ksize == 1 || ksize == 3, "conv2d_direct: ksize %d unsupported", ksize);
    const long long total = static_cast<long long>(B) * H * W * Cout;
    int grid = static_cast<int>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    conv_direct_kernel<<<grid, 256, 0, as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w),
        static_cast<const __nv_bfloat16*>(residual), static_cast<const __nv_bfloat16*>(mask), bias,
        static_cast<__nv_bfloat16*>(y), B, H, W, Cin, Cout, ksize,
        (flags & P2I_CONV_RELU) ? 1 : ((flags & P2I_CONV_LEAKY) ? 2 : 0));
    P2I_CHECK_LAUNCH("conv_direct_kernel");
    return P2I_OK;
}
