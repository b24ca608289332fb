// Weight gradient of the stride-1 k x k convolutions as a tcgen05 GEMM whose reduction (K) axis is the
// PIXEL axis:   dW[tap][co][ci] = sum_pix dY[pix][co] * X[pix + tap][ci]
// (backward of F.conv2d at p2igan_bench/modules/deconv_pytorch.py:108 w.r.t. its weight, and of the
//  UPPos 1x1 projection, layer.py:390).
//
// Both operands are NHWC bf16 tiles fetched by TMA exactly as in the forward kernel ([pixels][64 channels],
// 128 B per pixel, 128B swizzle).  With pixels as K they are "MN-major" UMMA operands: the 64 channels of
// a pixel are the contiguous M/N atom, 8 pixels form a 1024-B K group (SBO), further 64-channel atoms are
// LBO bytes apart.
//   A (M side) = X halo tile: (Ht+k-1) x Wt pixels.  The vertical taps are row-shifted views (ky*Wt pixels,
//                a multiple of 1024 B) of the same box, as in the forward kernel.
//   B (N side) = dY tile: Ht x Wt pixels, NT output channels.
//   D[ci][co]  = one fp32 TMEM accumulator per vertical tap, accumulated over all pixel tiles of the CTA's
//                K split, then reduced into dW with coalesced fp32 atomics (lanes = consecutive ci).
// Work item = (kx, ci tile, co tile, K split).  C = 64 layers stack two vertical taps into one M = 128
// instruction (second atom = the view one tile row further down: LBO = Wt*128 B).
#include "common.h"
#include "ptx.cuh"

namespace p2i {

struct WgradParams {
    int F, T_out, T_in, H, W, Cin, Cout;   // F = samples * T_out frames of dY
    int KT, KH, KW;
    int pad, pad_t, st;                    // x_in = x + kx - pad ; t_in = st*t_out + kt - pad_t
    int Ht, Wt, tiles_x, tiles_y, pix_tiles;
    int ci_tiles, co_tiles, ksplit;
    int stacked;     // 1: Cin tile = 64 channels, two vertical taps per MMA
    int nt;          // co tile width (64 or 128)
    float* dW;       // [KH*KW][Cout][Cin] fp32, accumulated with atomics (caller zero-fills)
};

constexpr int WG_STAGES = 3;
constexpr int WG_X_ATOM = 20480;   // (8+2)*16*128 B, also >= (16+2)*8*128
constexpr int WG_Y_ATOM = 16384;
constexpr int WG_STAGE_BYTES = 2 * WG_X_ATOM + 2 * WG_Y_ATOM;   // 73728
constexpr int WG_SMEM = 1024 + WG_STAGES * WG_STAGE_BYTES + 64;

__global__ void __launch_bounds__(256, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* done = empty + WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ngroups = p.stacked ? 2 : p.KH;             // accumulators per CTA
    const uint32_t tmem_cols = (ngroups * p.nt <= 128) ? 128 : ((ngroups * p.nt <= 256) ? 256 : 512);

    if (threadIdx.x == 0) {
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmY); }
    if (warp == 2) { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work item decode
    int item = blockIdx.x;
    const int ks = item % p.ksplit; item /= p.ksplit;
    const int cot = item % p.co_tiles; item /= p.co_tiles;
    const int cit = item % p.ci_tiles; item /= p.ci_tiles;
    const int kx = item % p.KW;
    const int kt = item / p.KW;
    const int ci_atoms = p.stacked ? 1 : 2;
    const int co_atoms = p.nt >> 6;
    const int ci0 = cit * (p.stacked ? 64 : 128), co0 = cot * p.nt;
    const int t_begin = static_cast<int>(static_cast<long long>(p.pix_tiles) * ks / p.ksplit);
    const int t_end = static_cast<int>(static_cast<long long>(p.pix_tiles) * (ks + 1) / p.ksplit);
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const uint32_t x_bytes = static_cast<uint32_t>((p.Ht + p.KH - 1) * p.Wt * 128);
    const uint32_t stage_tx = ci_atoms * x_bytes + co_atoms * WG_Y_ATOM;

    if (warp == 0) {
        if (elect_one()) {
            griddep_wait();                 // both operands come from preceding kernels (ptx.cuh: programmatic dependent launch)
            griddep_launch_dependents();
            uint32_t s = 0, ph = 0;
            for (int t = t_begin; t < t_end; ++t) {
                const int f = t / tiles_per_img, r = t - f * tiles_per_img;
                const int smp = f / p.T_out, t_out = f - smp * p.T_out;
                const int t_in = p.st * t_out + kt - p.pad_t;      // out-of-range frames are zero-filled by TMA
                const int y0 = (r / p.tiles_x) * p.Ht, x0 = (r % p.tiles_x) * p.Wt;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], stage_tx);
                uint8_t* st = smem + s * WG_STAGE_BYTES;
                for (int a = 0; a < ci_atoms; ++a)
                    tma_load_5d(st + a * WG_X_ATOM, &tmX, &full[s], ci0 + a * 64, x0 + kx - p.pad, y0 - p.pad, t_in, smp);
                for (int a = 0; a < co_atoms; ++a)
                    tma_load_5d(st + 2 * WG_X_ATOM + a * WG_Y_ATOM, &tmY, &full[s], co0 + a * 64, x0, y0, t_out, smp);
                if (++s == WG_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, p.nt, 1, 1);
            const uint32_t row_shift = static_cast<uint32_t>(p.Wt * 128);
            uint32_t s = 0, ph = 0;
            for (int t = t_begin; t < t_end; ++t) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t xb = smem_u32(smem + s * WG_STAGE_BYTES);
                const uint32_t yb = xb + 2 * WG_X_ATOM;
                const uint32_t acc = (t > t_begin) ? 1u : 0u;
                for (int g = 0; g < ngroups; ++g) {
                    // stacked: group g covers vertical taps 2g, 2g+1 (atoms one tile-row apart)
                    const uint32_t a0 = xb + (p.stacked ? 2 * g : g) * row_shift;
                    const uint32_t a_lbo = p.stacked ? row_shift : static_cast<uint32_t>(WG_X_ATOM);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        umma_bf16(tmem_base + g * p.nt, make_sw128_desc(a0 + k * 2048, a_lbo, 1024),
                                  make_sw128_desc(yb + k * 2048, WG_Y_ATOM, 1024), idesc, acc | (k > 0 ? 1u : 0u));
                    }
                }
                umma_commit(&empty[s]);
                if (++s == WG_STAGES) { s = 0; ph ^= 1; }
            }
            umma_commit(done);
        }
    } else if (warp >= 4 && t_end > t_begin) {
        const int ew = warp - 4;
        const int row = ew * 32 + lane;
        griddep_wait();
        mbar_wait(done, 0);
        tc_fence_after();
        for (int g = 0; g < ngroups; ++g) {
            int ky, ci;
            bool ok = true;
            if (p.stacked) {
                ky = 2 * g + (row >> 6);
                ci = ci0 + (row & 63);
                ok = ky < p.KH;
            } else {
                ky = g;
                ci = ci0 + row;
            }
            const int tap = (kt * p.KH + ky) * p.KW + kx;
            float* dst = p.dW + (static_cast<size_t>(tap) * p.Cout + co0) * p.Cin + ci;
            const uint32_t t_addr = tmem_base + g * p.nt + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < p.nt; c += 16) {
                uint32_t v[16];
                tmem_ld16(t_addr + c, v);
                tmem_ld_wait();
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(dst + static_cast<size_t>(c + i) * p.Cin, __uint_as_float(v[i]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}


// ------------------------------------------------------------------------------------------------
// Second generation.  The first-generation epilogue issues one scalar fp32 RED per accumulator element and lane
// (lane = input channel, so a thread's consecutive registers are Cin*4 bytes apart in dW); at ~1.3 clk per lane-RED
// and SM the split-K reduction cost as much as the MMA main loop.  Here the GEMM is FLIPPED: M = output channels
// (dY is the A operand), N = input channels (X is the B operand), so a thread owns one `co` row and its registers
// are consecutive `ci` -> 16-byte vector REDs (red.global.add.v4.f32, REDG.E.ADD.F32x4), 4x fewer reduction ops.
//   mode 0 (FLIP): Cout % 128 == 0, Cin == 64 or Cin % 128 == 0.  Work item = (kt, kx, ci tile, co tile, K split);
//                  one accumulator [128 co][N ci] per vertical tap (row-shifted views of the X box).
//   mode 1 (TAPS): Cin == Cout == 64, 3x3.  ALL nine taps in one CTA from one X box and one dY box per pixel tile:
//                  N = 3 vertical taps x 64 ci (X views one tile row apart: LBO = Wt*128 B), M = 2 horizontal taps
//                  x 64 co (dY views ONE PIXEL apart: LBO = 128 B inside a (Wt+2)-wide box; the 128-byte swizzle of
//                  TMA and UMMA is a function of the absolute shared-memory address, so pixel-granular starts are
//                  valid).  Two MMAs per K step: {kx=2, kx=1} and {kx=0, unused}; X and dY are fetched once per
//                  tile instead of once per horizontal tap.
// ------------------------------------------------------------------------------------------------
struct Wgrad2Params {
    int F, T_out, T_in, H, W, Cin, Cout;
    int KT, KH, KW;
    int pad, pad_t, st;
    int Ht, Wt, tiles_x, tiles_y, pix_tiles;
    int ci_tiles, co_tiles, ksplit;
    int mode;        // 0 FLIP, 1 TAPS
    int n_ci;        // FLIP: N of one MMA (64 or 128)
    float* dW;
};

constexpr int WG2_X_REGION = 22528;     // (8+2)*16*128 = 20480 used by TMA + one spare tile row read by the unused tap
constexpr int WG2_Y_REGION = 20480;     // TAPS: Ht x (Wt+2) pixels (18432 / 20480 B); FLIP: atom 0 (16384 B)
constexpr int WG2_STAGE_FLIP = 2 * WG_X_ATOM + 2 * WG_Y_ATOM;      // 73728
constexpr int WG2_STAGE_TAPS = WG2_X_REGION + WG2_Y_REGION;        // 43008
constexpr int WG2_SMEM = 1024 + 3 * WG2_STAGE_FLIP + 128;          // 222336 >= 4 * WG2_STAGE_TAPS + ...

__device__ __forceinline__ void red_add_v4(float* dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)),
                 "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
                 : "memory");
}

__global__ void __launch_bounds__(256, 1)
conv_wgrad2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const Wgrad2Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const bool taps = p.mode == 1;
    const int stages = taps ? 4 : 3;
    const int stage_bytes = taps ? WG2_STAGE_TAPS : WG2_STAGE_FLIP;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 3 * WG2_STAGE_FLIP);
    uint64_t* empty = full + 4;
    uint64_t* done = empty + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nacc = taps ? 2 : p.KH;
    const int ncol = taps ? 192 : p.n_ci;
    const uint32_t tmem_cols = (nacc * ncol <= 128) ? 128 : ((nacc * ncol <= 256) ? 256 : 512);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmY); }
    if (warp == 2) { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    int item = blockIdx.x;
    const int ks = item % p.ksplit; item /= p.ksplit;
    int cot = 0, cit = 0, kx = 0, kt;
    if (taps) {
        kt = item;
    } else {
        cot = item % p.co_tiles; item /= p.co_tiles;
        cit = item % p.ci_tiles; item /= p.ci_tiles;
        kx = item % p.KW;
        kt = item / p.KW;
    }
    const int ci_atoms = taps ? 1 : (p.n_ci >> 6);
    const int ci0 = cit * p.n_ci, co0 = cot * 128;
    const int t_begin = static_cast<int>(static_cast<long long>(p.pix_tiles) * ks / p.ksplit);
    const int t_end = static_cast<int>(static_cast<long long>(p.pix_tiles) * (ks + 1) / p.ksplit);
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const uint32_t x_bytes = static_cast<uint32_t>((p.Ht + p.KH - 1) * p.Wt * 128);
    const uint32_t y_bytes = taps ? static_cast<uint32_t>(p.Ht * (p.Wt + 2) * 128) : 2u * WG_Y_ATOM;
    const uint32_t stage_tx = ci_atoms * x_bytes + y_bytes;
    const uint32_t y_off = taps ? WG2_X_REGION : 2 * WG_X_ATOM;

    if (warp == 0) {
        if (elect_one()) {
            uint32_t s = 0, ph = 0;
            for (int t = t_begin; t < t_end; ++t) {
                const int f = t / tiles_per_img, r = t - f * tiles_per_img;
                const int smp = f / p.T_out, t_out = f - smp * p.T_out;
                const int t_in = p.st * t_out + kt - p.pad_t;      // out-of-range frames are zero-filled by TMA
                const int y0 = (r / p.tiles_x) * p.Ht, x0 = (r % p.tiles_x) * p.Wt;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], stage_tx);
                uint8_t* st = smem + s * stage_bytes;
                if (taps) {
                    tma_load_5d(st, &tmX, &full[s], 0, x0, y0 - p.pad, t_in, smp);
                    tma_load_5d(st + y_off, &tmY, &full[s], 0, x0 - 1, y0, t_out, smp);
                } else {
                    for (int a = 0; a < ci_atoms; ++a)
                        tma_load_5d(st + a * WG_X_ATOM, &tmX, &full[s], ci0 + a * 64, x0 + kx - p.pad, y0 - p.pad, t_in, smp);
                    for (int a = 0; a < 2; ++a)
                        tma_load_5d(st + y_off + a * WG_Y_ATOM, &tmY, &full[s], co0 + a * 64, x0, y0, t_out, smp);
                }
                if (++s == static_cast<uint32_t>(stages)) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, ncol, 1, 1);
            const uint32_t row_shift = static_cast<uint32_t>(p.Wt * 128);
            // TAPS: a K = 16 step is one 16-pixel row (Wt = 16) or two 8-pixel rows (Wt = 8) of the (Wt+2)-wide dY box
            const uint32_t y_pitch = static_cast<uint32_t>((p.Wt + 2) * 128);
            const uint32_t y_sbo = taps ? (p.Wt == 16 ? 1024u : y_pitch) : 1024u;
            const uint32_t y_adv = taps ? (p.Wt == 16 ? y_pitch : 2 * y_pitch) : 2048u;
            uint32_t s = 0, ph = 0;
            for (int t = t_begin; t < t_end; ++t) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t xb = smem_u32(smem + s * stage_bytes);
                const uint32_t yb = xb + y_off;
                const uint32_t acc = (t > t_begin) ? 1u : 0u;
                if (taps) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t bd = make_sw128_desc(xb + k * 2048, row_shift, 1024);            // ky = 0, 1, 2
                        umma_bf16(tmem_base, make_sw128_desc(yb + k * y_adv, 128, y_sbo), bd, idesc, acc | (k > 0 ? 1u : 0u));
                        umma_bf16(tmem_base + 192, make_sw128_desc(yb + k * y_adv + 256, 128, y_sbo), bd, idesc, acc | (k > 0 ? 1u : 0u));
                    }
                } else {
                    for (int g = 0; g < p.KH; ++g) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma_bf16(tmem_base + g * ncol, make_sw128_desc(yb + k * 2048, WG_Y_ATOM, 1024),
                                      make_sw128_desc(xb + g * row_shift + k * 2048, WG_X_ATOM, 1024), idesc, acc | (k > 0 ? 1u : 0u));
                    }
                }
                umma_commit(&empty[s]);
                if (++s == static_cast<uint32_t>(stages)) { s = 0; ph ^= 1; }
            }
            umma_commit(done);
        }
    } else if (warp >= 4 && t_end > t_begin) {
        const int ew = warp - 4;
        const int row = ew * 32 + lane;
        mbar_wait(done, 0);
        tc_fence_after();
        for (int g = 0; g < nacc; ++g) {
            const uint32_t t_addr = tmem_base + g * ncol + (static_cast<uint32_t>(ew * 32) << 16);
            int kxx = kx, co = co0 + row;
            bool ok = true;
            if (taps) {
                const int a = row >> 6;
                co = row & 63;
                kxx = (g == 0) ? 2 - a : 0;
                ok = (g == 0) || a == 0;
            }
#pragma unroll 1
            for (int c = 0; c < ncol; c += 16) {
                uint32_t v[16];
                tmem_ld16(t_addr + c, v);
                tmem_ld_wait();
                const int ky = taps ? (c >> 6) : g;
                const int ci = taps ? (c & 63) : ci0 + c;
                const int tap = (kt * p.KH + ky) * p.KW + kxx;
                float* dst = p.dW + (static_cast<size_t>(tap) * p.Cout + co) * p.Cin + ci;
                if (ok) {
                    red_add_v4(dst, v[0], v[1], v[2], v[3]);
                    red_add_v4(dst + 4, v[4], v[5], v[6], v[7]);
                    red_add_v4(dst + 8, v[8], v[9], v[10], v[11]);
                    red_add_v4(dst + 12, v[12], v[13], v[14], v[15]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

static std::atomic<int> g_wgrad_impl{0};   // 0 auto (second generation when eligible), 1 first generation only

}  // namespace p2i

using namespace p2i;

static int run_wgrad(const void* x, const void* dy, float* dW, const P2iConvDesc& d, void* stream) {
    P2I_CHECK_ARG(x && dy && dW, "conv_wgrad: null pointer");
    P2I_CHECK_ARG(d.ksize >= 1 && d.ksize <= 3 && (d.kt == 1 || d.kt == 3), "conv_wgrad: taps unsupported");
    const int Cin = d.Cin, Cout = d.Cout;
    P2I_CHECK_ARG(Cin % 64 == 0 && Cout % 64 == 0 && Cin > 0 && Cout > 0, "conv_wgrad: channels must be multiples of 64");
    P2I_CHECK_ARG(Cin == 64 || Cin % 128 == 0, "conv_wgrad: Cin=%d must be 64 or a multiple of 128", Cin);
    P2I_CHECK_ARG(Cin != 64 || d.ksize == 3 || d.ksize == 1, "conv_wgrad: Cin=64 supports k in {1,3}");
    {
        const bool taps_ok = Cin == 64 && Cout == 64 && d.ksize == 3 && d.pad == 1;
        // Measured (tools/bench_wgrad.py, profiles/r1_wgrad_ab.txt): once the K split keeps every launch to ONE wave,
        // the first-generation kernel wins or ties everywhere.  FLIP is ~10 % slower (lanes = co: every 16-byte RED of
        // a warp lands in a different 32-byte sector, 32 sector operations per instruction against 4 for the first
        // generation's 128 contiguous bytes); TAPS needs 2.2x more reduction traffic at a 148-way K split (9 taps x
        // 64 x 64 per CTA) and ends 8 % behind the stacked first-generation layout (35.5 vs 32.9 us at B = 16).
        // Both stay selectable (impl 2) for A/B runs; automatic selection uses the first generation.
        const int impl = g_wgrad_impl.load(std::memory_order_relaxed);
        const bool use2 = impl == 2;
        const bool flip_ok = use2 && !taps_ok && Cout % 128 == 0 && (Cin == 64 || Cin % 128 == 0);
        if (use2 && (taps_ok || flip_ok)) {
            Wgrad2Params q;
            q.F = d.samples * d.T_out; q.T_out = d.T_out; q.T_in = d.T_in;
            q.H = d.H; q.W = d.W; q.Cin = Cin; q.Cout = Cout;
            q.KT = d.kt; q.KH = d.ksize; q.KW = d.ksize; q.pad = d.pad; q.pad_t = d.pad_t; q.st = d.stride_t;
            q.Wt = (d.W >= 16) ? 16 : 8;
            q.Ht = 128 / q.Wt;
            q.tiles_x = cdiv(d.W, q.Wt);
            q.tiles_y = cdiv(d.H, q.Ht);
            q.pix_tiles = q.F * q.tiles_x * q.tiles_y;
            q.mode = taps_ok ? 1 : 0;
            q.n_ci = (Cin == 64) ? 64 : 128;
            q.ci_tiles = taps_ok ? 1 : Cin / q.n_ci;
            q.co_tiles = taps_ok ? 1 : Cout / 128;
            const int items = taps_ok ? q.KT : q.KT * q.KW * q.ci_tiles * q.co_tiles;
            int ks = sm_count() / items;          // floor: one wave of at most sm_count() CTAs (1 CTA per SM)
            if (ks > q.pix_tiles) ks = q.pix_tiles;
            if (ks < 1) ks = 1;
            q.ksplit = ks;
            q.dW = dW;
            CUtensorMap tmX, tmY;
            {
                const uint64_t C = Cin, W = d.W, H = d.H, T = d.T_in;
                const uint64_t dims[5] = {C, W, H, T, uint64_t(d.samples)};
                const uint64_t strides[5] = {0, C * 2, W * C * 2, H * W * C * 2, T * H * W * C * 2};
                const uint32_t box[5] = {64, uint32_t(q.Wt), uint32_t(q.Ht + d.ksize - 1), 1, 1};
                int rc = encode_tmap_bf16(&tmX, x, 5, dims, strides, box, nullptr, true);
                if (rc) return rc;
            }
            {
                const uint64_t C = Cout, W = d.W, H = d.H, T = d.T_out;
                const uint64_t dims[5] = {C, W, H, T, uint64_t(d.samples)};
                const uint64_t strides[5] = {0, C * 2, W * C * 2, H * W * C * 2, T * H * W * C * 2};
                const uint32_t box[5] = {64, uint32_t(taps_ok ? q.Wt + 2 : q.Wt), uint32_t(q.Ht), 1, 1};
                int rc = encode_tmap_bf16(&tmY, dy, 5, dims, strides, box, nullptr, true);
                if (rc) return rc;
            }
            static bool configured2 = false;
            if (!configured2) {
                cudaError_t e = cudaFuncSetAttribute(conv_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG2_SMEM);
                if (e != cudaSuccess) return fail(P2I_ERR_CUDA, "conv_wgrad2 smem attribute: %s", cudaGetErrorString(e));
                configured2 = true;
            }
            set_last_variant(4000000 + q.mode);
            conv_wgrad2_kernel<<<items * ks, 256, WG2_SMEM, as_stream(stream)>>>(tmX, tmY, q);
            P2I_CHECK_LAUNCH("conv_wgrad2_kernel");
            return P2I_OK;
        }
    }
    WgradParams p;
    p.F = d.samples * d.T_out; p.T_out = d.T_out; p.T_in = d.T_in;
    p.H = d.H; p.W = d.W; p.Cin = Cin; p.Cout = Cout;
    p.KT = d.kt; p.KH = d.ksize; p.KW = d.ksize; p.pad = d.pad; p.pad_t = d.pad_t; p.st = d.stride_t;
    p.Wt = (d.W >= 16) ? 16 : 8;
    p.Ht = 128 / p.Wt;
    p.tiles_x = cdiv(d.W, p.Wt);
    p.tiles_y = cdiv(d.H, p.Ht);
    p.pix_tiles = p.F * p.tiles_x * p.tiles_y;
    p.stacked = (Cin == 64) ? 1 : 0;
    p.nt = (Cout % 128 == 0) ? 128 : 64;
    p.ci_tiles = p.stacked ? 1 : Cin / 128;
    p.co_tiles = Cout / p.nt;
    const int items = p.KT * p.KW * p.ci_tiles * p.co_tiles;
    // floor, not ceil: the kernel holds one CTA per SM, so items * ks > sm_count() would run a second, nearly empty wave
    int ks = sm_count() / items;
    if (ks > p.pix_tiles) ks = p.pix_tiles;
    if (ks < 1) ks = 1;
    p.ksplit = ks;
    p.dW = dW;

    CUtensorMap tmX, tmY;
    {
        const uint64_t C = Cin, W = d.W, H = d.H, T = d.T_in;
        const uint64_t dims[5] = {C, W, H, T, uint64_t(d.samples)};
        const uint64_t strides[5] = {0, C * 2, W * C * 2, H * W * C * 2, T * H * W * C * 2};
        const uint32_t box[5] = {64, uint32_t(p.Wt), uint32_t(p.Ht + d.ksize - 1), 1, 1};
        int rc = encode_tmap_bf16(&tmX, x, 5, dims, strides, box, nullptr, true);
        if (rc) return rc;
    }
    {
        const uint64_t C = Cout, W = d.W, H = d.H, T = d.T_out;
        const uint64_t dims[5] = {C, W, H, T, uint64_t(d.samples)};
        const uint64_t strides[5] = {0, C * 2, W * C * 2, H * W * C * 2, T * H * W * C * 2};
        const uint32_t box[5] = {64, uint32_t(p.Wt), uint32_t(p.Ht), 1, 1};
        int rc = encode_tmap_bf16(&tmY, dy, 5, dims, strides, box, nullptr, true);
        if (rc) return rc;
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
        if (e != cudaSuccess) return fail(P2I_ERR_CUDA, "conv_wgrad smem attribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    set_last_variant(3000000 + p.nt * 100 + (p.stacked ? 10 : 0));
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(items * ks);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = WG_SMEM;
        cfg.stream = as_stream(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_enabled() ? 1 : 0;
        cudaError_t e = cudaLaunchKernelEx(&cfg, conv_wgrad_kernel, tmX, tmY, p);
        if (e != cudaSuccess) return fail(P2I_ERR_CUDA, "conv_wgrad launch: %s", cudaGetErrorString(e));
    }
    P2I_CHECK_LAUNCH("conv_wgrad_kernel");
    return P2I_OK;
}

extern "C" int p2i_set_wgrad_impl(int impl) {
    P2I_CHECK_ARG(impl >= 0 && impl <= 2, "set_wgrad_impl: 0 auto | 1 first generation | 2 experimental all-taps / flipped kernels");
    g_wgrad_impl.store(impl);
    return P2I_OK;
}

extern "C" int p2i_conv_wgrad(const void* x, const void* dy, float* dW, const P2iConvDesc* desc, void* stream) {
    P2I_CHECK_ARG(desc, "conv_wgrad: null descriptor");
    return run_wgrad(x, dy, dW, *desc, stream);
}

extern "C" int p2i_conv2d_wgrad(const void* x, const void* dy, float* dW, int B, int H, int W, int Cin, int Cout,
                                int ksize, void* stream) {
    P2I_CHECK_ARG(ksize == 1 || ksize == 3, "conv2d_wgrad: ksize %d unsupported", ksize);
    P2iConvDesc d;
    d.samples = B; d.T_in = 1; d.T_out = 1; d.H = H; d.W = W; d.Cin = Cin; d.Cout = Cout;
    d.kt = 1; d.ksize = ksize; d.pad = ksize / 2; d.pad_t = 0; d.stride_t = 1; d.t_transposed = 0;
    d.act = 0; d.mask_mode = 0; d.out_mode = 0;
    return run_wgrad(x, dy, dW, d, stream);
}
