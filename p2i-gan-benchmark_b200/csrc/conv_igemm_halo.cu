// Convolution as a tcgen05 implicit GEMM, second generation ("halo" kernel) for sm_100a.
//
// Same contract as conv_igemm.cu (reference call sites: F.conv2d in DOConv2d._conv_forward,
// p2igan_bench/modules/deconv_pytorch.py:104-109; the 1x1 projection of UPPos, layer.py:390,398; the spectral-norm
// Conv2d / Conv3d stacks of P2IDiscriminator, models/p2igan.py:120-142) with a different data flow:
//
//   A (activations): ONE TMA box per (temporal tap, 64-channel block) and CTA tile: (16+k-1) x (8+k-1) pixels x 64 ch
//       for a 16 x 8 pixel output tile.  All k*k spatial taps are descriptor views of that box: tap (ky,kx) starts
//       (ky*(8+k-1) + kx) * 128 B into it and its 8-pixel row groups are (8+k-1)*128 B apart (SBO).  The 128-byte
//       swizzle of TMA and UMMA is a function of the absolute shared-memory address, so neither the 128-B start
//       offset nor the 1280-B group pitch needs 1024-B alignment (tools/exp_umma_shift.cu measures exactly this).
//       Activations are fetched from L2 once per tile with a 1.4x halo instead of k times.
//   CG = 2 (CTA pairs, tcgen05 cta_group::2): one MMA covers M = 256 pixels (one 16x8 tile per CTA) x N = NT output
//       channels; each CTA holds its own activation box and HALF of the weight rows.  A tcgen05.mma with both
//       operands in shared memory is bound by the operand fetch (~64 B/clk/SM measured: 33 / 50 / 67 % of the
//       tensor peak at N = 64 / 128 / 256 for a single CTA); sharing B across the pair lifts that to 40 / 67 / 100 %.
//   B (weights, bf16 [tap][Cout][Cin]):
//       RES  : when all taps x channel blocks of this CTA's weight rows fit in shared memory they are loaded ONCE per
//              CTA and stay resident (64/128-channel generator levels, first discriminator layers, 1x1 projections);
//       else : streamed in (NT/CG) x 64 blocks through a ring (CG = 1: each block optionally shared by MB = 2 pixel
//              tiles with two accumulators).
//   D : fp32 in TMEM, NACC x MB x NT columns (double buffered when it fits in 512 columns).
//   Epilogue: TMEM -> registers -> bias / residual / activation / mask -> bf16 -> 128B-swizzled staging tile in
//       shared memory -> TMA store (clips ragged tiles).  The residual / mask tile is TMA-loaded into the SAME staging
//       buffer ahead of time and overwritten in place.  Space-to-depth pack / unpack are just different store maps.
// Warp roles (384 threads, persistent): warp0 TMA producer (A, B), warp1 MMA issuer (leader CTA of a pair only),
// warp2 TMEM allocator + aux-tile producer, warps 4-11 epilogue (two warps per TMEM lane quarter, 32 columns each).
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace p2i {

struct HaloParams {
    int F, T_out, T_in, H, W, Cin, Cout;
    int KT, KS;
    int pad, pad_t, st, tmode;
    int tiles_x, tiles_y, m_tiles, n_tiles, units;
    int MB, SA, SB, NACC, tmem_cols;
    int act;        // 0 none, 1 ReLU, 2 LeakyReLU(0.2)
    int aux_mode;   // 0 none, 1 residual (added before the activation), 2 zero where aux <= 0, 3 x0.2 where aux <= 0
    int out_mode;   // 0 natural, 1 space-to-depth pack, 2 unpack
    int cq;         // out_mode 2: channels of the unpacked tensor (Cout / 4)
    int dbg;        // -DHALO_PROF builds only (env P2I_HALO_DBG): 1 = the epilogue neither stages nor stores its tile, 2 = no TMEM reads either
    uint32_t offB, offS, offBias, offBar;
    const float* bias;
};

// -DHALO_PROF: block 0 prints where each role spends its cycles (waits vs work).  Diagnostics only.
#ifdef HALO_PROF
#define PROF_DECL(...) long long __VA_ARGS__
#define PROF_T0() const long long prof_t0 = clock64()
#define PROF_ADD(var) var += clock64() - prof_t0
#define PROF_WAIT(var, bar, par) do { const long long t__ = clock64(); mbar_wait(bar, par); var += clock64() - t__; } while (0)
#else
#define PROF_DECL(...)
#define PROF_T0()
#define PROF_ADD(var)
#define PROF_WAIT(var, bar, par) mbar_wait(bar, par)
#endif

constexpr int HALO_A_STRIDE = 23552;     // (16+2)*(8+2)*128 = 23040 rounded up to 1024
constexpr int HALO_STG = 16384;          // 128 pixels x 64 channels bf16
constexpr int HALO_NST = 2;
constexpr int HALO_SMEM_MAX = 232448;
constexpr int HALO_THREADS = 384;      // warps 0-3: producer / MMA / alloc+aux / idle; warps 4-11: epilogue
constexpr int HALO_EPI_THREADS = 256;

__device__ __forceinline__ int halo_src_frame(const HaloParams& p, int t_out, int kt, bool& skip) {
    skip = false;
    if (p.tmode == 0) return p.st * t_out + kt - p.pad_t;
    const int s = t_out + p.pad_t - kt;
    if (s % p.st != 0) { skip = true; return 0; }
    return s / p.st;
}

struct TileCoord { int f, smp, t_out, y0, x0; };
__device__ __forceinline__ TileCoord halo_tile(const HaloParams& p, int mt) {
    TileCoord c;
    const int tpi = p.tiles_x * p.tiles_y;
    c.f = mt / tpi;
    const int r = mt - c.f * tpi;
    c.smp = c.f / p.T_out;
    c.t_out = c.f - c.smp * p.T_out;
    c.y0 = (r / p.tiles_x) * 16;
    c.x0 = (r % p.tiles_x) * 8;
    return c;
}

// Unit u of a CTA covers the M tiles  (u*MB + i, i < MB)  for CG == 1  and tile  2u + rank  for CG == 2 (MB == 1).
template <int NT, bool RES, int CG>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmX, const HaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + p.offB;
    uint8_t* sS = smem + p.offS;
    float* sBias = reinterpret_cast<float*>(smem + p.offBias);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.offBar);
    uint64_t* fullA = bars;            // [8]   CG == 2: the leader's copy collects both CTAs' bytes
    uint64_t* emptyA = bars + 8;       // [8]   per CTA (the MMA commit is multicast to both)
    uint64_t* fullB = bars + 16;       // [16]
    uint64_t* emptyB = bars + 32;      // [16]
    uint64_t* tfull = bars + 48;       // [2]   per CTA (multicast commit)
    uint64_t* tempty = bars + 50;      // [2]   leader's copy collects the epilogue warps of both CTAs
    uint64_t* sfull = bars + 52;       // [4]   staging buffers: per CTA
    uint64_t* sempty = bars + 56;      // [4]
    uint64_t* bres = bars + 60;        // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 61);

    constexpr int NB = NT / CG;              // weight rows held by this CTA
    constexpr int BBLK = NB * 128;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < 16; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], (HALO_EPI_THREADS / 32) * CG); }
        for (int i = 0; i < 4; ++i) { mbar_init(&sfull[i], 1); mbar_init(&sempty[i], 1); }
        mbar_init(bres, 1);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
        if (p.aux_mode) tma_prefetch_desc(&tmX);
    }
    if (warp == 2) {
        if (CG == 2) { tmem_alloc_cg2(tmem_slot, p.tmem_cols); tmem_relinquish_cg2(); }
        else { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();     // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int cblocks = p.Cin >> 6;
    const int taps = p.KS * p.KS;
    const int bw = 8 + p.KS - 1;
    const uint32_t a_bytes = static_cast<uint32_t>((16 + p.KS - 1) * bw * 128);
    const int cta = (CG == 2) ? (blockIdx.x >> 1) : blockIdx.x;
    const int ncta = (CG == 2) ? (gridDim.x >> 1) : gridDim.x;
    const int nt = cta % p.n_tiles;
    const int u0 = cta / p.n_tiles, ustep = ncta / p.n_tiles;
    // ONE: exactly one M tile per unit, known at compile time (pairs; resident-weight kernels never block over M).  Keeps the
    // per-unit tile arrays of the role loops in registers: with a run-time count they live in local memory, and the LDL latency
    // on the single MMA-issuing thread cost the 64-channel level 105 clk per instruction against a 64-clk floor (HALO_PROF).
    constexpr bool ONE = (CG == 2) || RES;
    const int MBc = ONE ? 1 : p.MB;
    // M tile i of unit u for this CTA, clamped: a pair's odd last unit computes a duplicate tile that is not stored
    auto tile_of = [&](int u, int i) -> int {
        const int mt = (CG == 2) ? (2 * u + static_cast<int>(rank)) : (u * p.MB + i);
        return mt < p.m_tiles ? mt : p.m_tiles - 1;
    };

    if (warp == 0) {
        if (elect_one()) {
            // ------------------------------------------------------------ TMA producer: resident weights + activation boxes
            // CG == 2: completion bytes of BOTH CTAs land on the leader's barriers; only the leader posts expect_tx
            const int b_row0 = nt * NT + static_cast<int>(rank) * NB;
            if (RES) {
                // weights were final two or more kernels ago (host contract): fetched BEFORE the dependency wait, under the
                // predecessor's tail
                const int nblk = p.KT * taps * cblocks;
                if (rank == 0) mbar_expect_tx(bres, static_cast<uint32_t>(nblk) * BBLK * CG);
                const uint32_t bar = (CG == 2) ? mapa_u32(smem_u32(bres), 0) : 0u;
                for (int t = 0; t < p.KT * taps; ++t)
                    for (int cb = 0; cb < cblocks; ++cb) {
                        uint8_t* dst = sB + static_cast<size_t>(t * cblocks + cb) * BBLK;
                        if (CG == 2) tma_load_3d_cg2(dst, &tmB, bar, cb * 64, b_row0, t);
                        else tma_load_3d(dst, &tmB, bres, cb * 64, b_row0, t);
                    }
            }
            griddep_wait();                 // the activations (and everything else) come from the predecessor
            griddep_launch_dependents();
            uint32_t sa = 0, pa = 0;
            PROF_DECL(w_ea = 0, t_all = 0);
            PROF_T0();
            for (int u = u0; u < p.units; u += ustep) {
                const int nsub = ONE ? 1 : min(p.MB, p.m_tiles - u * p.MB);
                TileCoord tc[2];
                for (int i = 0; i < nsub; ++i) tc[i] = halo_tile(p, tile_of(u, i));
                for (int kt = 0; kt < p.KT; ++kt) {
                    bool skip;
                    (void)halo_src_frame(p, tc[0].t_out, kt, skip);   // tmode == 1: all tiles of a unit share t_out
                    if (skip) continue;
                    for (int cb = 0; cb < cblocks; ++cb) {
                        for (int i = 0; i < nsub; ++i) {
                            bool sk;
                            const int t_in = halo_src_frame(p, tc[i].t_out, kt, sk);
                            PROF_WAIT(w_ea, &emptyA[sa], pa ^ 1);
                            if (rank == 0) mbar_expect_tx(&fullA[sa], a_bytes * CG);
                            if (CG == 2)
                                tma_load_5d_cg2(sA + sa * HALO_A_STRIDE, &tmA, mapa_u32(smem_u32(&fullA[sa]), 0), cb * 64,
                                                tc[i].x0 - p.pad, tc[i].y0 - p.pad, t_in, tc[i].smp);
                            else
                                tma_load_5d(sA + sa * HALO_A_STRIDE, &tmA, &fullA[sa], cb * 64, tc[i].x0 - p.pad,
                                            tc[i].y0 - p.pad, t_in, tc[i].smp);
                            if (++sa == static_cast<uint32_t>(p.SA)) { sa = 0; pa ^= 1; }
                        }
                    }
                }
            }
#ifdef HALO_PROF
            PROF_ADD(t_all);
            if (blockIdx.x == 0) printf("halo prof A producer: total %lld  wait emptyA %lld\n", t_all, w_ea);
#endif
        }
    } else if (warp == 3) {
        if (!RES && elect_one()) {
            // ------------------------------------------------------------ TMA producer: streamed weight blocks
            // Same (unit, kt, cb, tap) order as the MMA issuer consumes them.  No dependency wait: the weights do not come from
            // the predecessor, so the first SB blocks are in flight while the predecessor is still draining.
            const int b_row0 = nt * NT + static_cast<int>(rank) * NB;
            uint32_t sb = 0, pb = 0;
            PROF_DECL(w_eb = 0, t_all = 0);
            PROF_T0();
            for (int u = u0; u < p.units; u += ustep) {
                const int t_out0 = p.tmode ? halo_tile(p, tile_of(u, 0)).t_out : 0;
                for (int kt = 0; kt < p.KT; ++kt) {
                    bool skip;
                    (void)halo_src_frame(p, t_out0, kt, skip);
                    if (skip) continue;
                    for (int cb = 0; cb < cblocks; ++cb) {
                        for (int t = 0; t < taps; ++t) {
                            PROF_WAIT(w_eb, &emptyB[sb], pb ^ 1);
                            if (rank == 0) mbar_expect_tx(&fullB[sb], BBLK * CG);
                            if (CG == 2)
                                tma_load_3d_cg2(sB + sb * BBLK, &tmB, mapa_u32(smem_u32(&fullB[sb]), 0), cb * 64, b_row0,
                                                kt * taps + t);
                            else
                                tma_load_3d(sB + sb * BBLK, &tmB, &fullB[sb], cb * 64, b_row0, kt * taps + t);
                            if (++sb == static_cast<uint32_t>(p.SB)) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            }
#ifdef HALO_PROF
            PROF_ADD(t_all);
            if (blockIdx.x == 0) printf("halo prof B producer: total %lld  wait emptyB %lld\n", t_all, w_eb);
#endif
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            // ------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc = make_idesc_bf16(128 * CG, NT);
            const uint32_t sbo = static_cast<uint32_t>(bw * 128);
            const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
            if (RES) {
                mbar_wait(bres, 0);
                tc_fence_after();
            }
            // descriptor constants: LBO (unused) | SBO | version 1 | SWIZZLE_128B; the 14-bit start address is OR-ed in
            const uint64_t a_hi = (static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32) |
                                  (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
            const uint64_t b_hi = (static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
                                  (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
            const uint32_t SA = static_cast<uint32_t>(p.SA), SB = static_cast<uint32_t>(p.SB);
            const uint32_t acc_cols = static_cast<uint32_t>(MBc * NT);
            const int KS = p.KS, KT = p.KT, tmode = p.tmode;
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0;     // ring positions / parities (no divisions on this thread)
            uint32_t acc = 0, acc_par = 0;
            PROF_DECL(w_fa = 0, w_fb = 0, w_te = 0, t_all = 0);
            PROF_T0();
            for (int u = u0; u < p.units; u += ustep) {
                const int nsub = ONE ? 1 : min(p.MB, p.m_tiles - u * p.MB);
                const int t_out0 = tmode ? halo_tile(p, tile_of(u, 0)).t_out : 0;
                PROF_WAIT(w_te, &tempty[acc], acc_par ^ 1);
                tc_fence_after();
                const uint32_t d_base = tmem_base + acc * acc_cols;
                uint32_t started = 0;
                for (int kt = 0; kt < KT; ++kt) {
                    if (tmode) {
                        bool skip;
                        (void)halo_src_frame(p, t_out0, kt, skip);
                        if (skip) continue;
                    }
                    for (int cb = 0; cb < cblocks; ++cb) {
                        // the nsub activation boxes of this K block
                        uint32_t a_base[2];
                        {
                            uint32_t s = sa, ph = pa;
                            for (int i = 0; i < nsub; ++i) {
                                PROF_WAIT(w_fa, &fullA[s], ph);
                                a_base[i] = sA_u + s * HALO_A_STRIDE;
                                if (++s == SA) { s = 0; ph ^= 1; }
                            }
                        }
                        tc_fence_after();
                        uint32_t b_res = sB_u + static_cast<uint32_t>(kt * taps * cblocks + cb) * BBLK;   // RES: tap stride cblocks*BBLK
                        uint32_t tap_off = 0;
                        for (int ky = 0; ky < KS; ++ky, tap_off += static_cast<uint32_t>(bw - KS) * 128) {
                            for (int kx = 0; kx < KS; ++kx, tap_off += 128) {
                                uint32_t b_addr;
                                if (RES) {
                                    b_addr = b_res;
                                    b_res += static_cast<uint32_t>(cblocks) * BBLK;
                                } else {
                                    PROF_WAIT(w_fb, &fullB[sb], pb);
                                    tc_fence_after();
                                    b_addr = sB_u + sb * BBLK;
                                }
                                const uint64_t bd = b_hi | static_cast<uint64_t>((b_addr >> 4) & 0x3FFF);
                                for (int i = 0; i < nsub; ++i) {
                                    const uint64_t ad = a_hi | static_cast<uint64_t>(((a_base[i] + tap_off) >> 4) & 0x3FFF);
                                    const uint32_t d = d_base + i * NT;
                                    if (CG == 2) {
                                        umma_bf16_cg2(d, ad, bd, idesc, (started >> i) & 1u);
                                        umma_bf16_cg2(d, ad + 2, bd + 2, idesc, 1u);
                                        umma_bf16_cg2(d, ad + 4, bd + 4, idesc, 1u);
                                        umma_bf16_cg2(d, ad + 6, bd + 6, idesc, 1u);
                                    } else {
                                        umma_bf16(d, ad, bd, idesc, (started >> i) & 1u);
                                        umma_bf16(d, ad + 2, bd + 2, idesc, 1u);
                                        umma_bf16(d, ad + 4, bd + 4, idesc, 1u);
                                        umma_bf16(d, ad + 6, bd + 6, idesc, 1u);
                                    }
                                    started |= 1u << i;
                                }
                                if (!RES) {
                                    if (CG == 2) umma_commit_cg2(&emptyB[sb]); else umma_commit(&emptyB[sb]);
                                    if (++sb == SB) { sb = 0; pb ^= 1; }
                                }
                            }
                        }
                        for (int i = 0; i < nsub; ++i) {
                            if (CG == 2) umma_commit_cg2(&emptyA[sa]); else umma_commit(&emptyA[sa]);
                            if (++sa == SA) { sa = 0; pa ^= 1; }
                        }
                    }
                }
                if (CG == 2) umma_commit_cg2(&tfull[acc]); else umma_commit(&tfull[acc]);
                if (++acc == static_cast<uint32_t>(p.NACC)) { acc = 0; acc_par ^= 1; }
            }
#ifdef HALO_PROF
            PROF_ADD(t_all);
            if (blockIdx.x == 0) printf("halo prof mma: total %lld  wait fullA %lld  wait fullB %lld  wait tempty %lld\n", t_all, w_fa, w_fb, w_te);
#endif
        }
    } else if (warp == 2) {
        if (p.aux_mode != 0 && elect_one()) {
            // ------------------------------------------------------------ residual / mask tiles -> staging buffers
            griddep_wait();
            uint32_t cs = 0;
            for (int u = u0; u < p.units; u += ustep) {
                const int nsub = ONE ? 1 : min(p.MB, p.m_tiles - u * p.MB);
                for (int i = 0; i < nsub; ++i) {
                    const TileCoord tc = halo_tile(p, tile_of(u, i));
                    for (int g = 0; g < NT / 64; ++g, ++cs) {
                        const uint32_t b = cs % HALO_NST;
                        mbar_wait(&sempty[b], ((cs / HALO_NST) & 1) ^ 1);
                        mbar_expect_tx(&sfull[b], HALO_STG);
                        tma_load_5d(sS + b * HALO_STG, &tmX, &sfull[b], nt * NT + g * 64, tc.x0, tc.y0, tc.f, 0);
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- epilogue
        // 8 warps: warp w reads TMEM lanes 32*(w%4).. (hardware lane quarter) and the 32-column half (w-4)/4 of each chunk
        const int ew = warp & 3;
        const int hh = (warp - 4) >> 2;
        const int row = ew * 32 + lane;
        const bool has_bias = p.bias != nullptr;
        griddep_wait();          // nothing of this role touches global memory before the predecessor has completed
        if (has_bias)
            for (int i = threadIdx.x - 128; i < NT; i += HALO_EPI_THREADS) sBias[i] = __ldg(p.bias + nt * NT + i);
        named_bar_sync(1, HALO_EPI_THREADS);
        const uint32_t row_off = static_cast<uint32_t>(row) * 128;
        const uint32_t xr = static_cast<uint32_t>(row & 7);
        uint32_t tempty_remote[2] = {0u, 0u};
        if (CG == 2 && rank != 0) {
            tempty_remote[0] = mapa_u32(smem_u32(&tempty[0]), 0);
            tempty_remote[1] = mapa_u32(smem_u32(&tempty[1]), 0);
        }
        uint32_t cs = 0;
        uint32_t acc = 0, acc_par = 0;
        PROF_DECL(w_tf = 0, w_st = 0, t_all = 0, t_store = 0, t_bar = 0);
        PROF_T0();
        for (int u = u0; u < p.units; u += ustep) {
            const int nsub = ONE ? 1 : min(p.MB, p.m_tiles - u * p.MB);
            PROF_WAIT(w_tf, &tfull[acc], acc_par);
            tc_fence_after();
            for (int i = 0; i < nsub; ++i) {
                const TileCoord tc = halo_tile(p, tile_of(u, i));
                const bool store_ok = (CG == 1) || (2 * u + static_cast<int>(rank) < p.m_tiles);
                for (int g = 0; g < NT / 64; ++g, ++cs) {
                    const uint32_t b = cs % HALO_NST;
                    const uint32_t ph = (cs / HALO_NST) & 1;
                    if (p.aux_mode != 0) PROF_WAIT(w_st, &sfull[b], ph);
                    else PROF_WAIT(w_st, &sempty[b], ph ^ 1);
                    uint8_t* stg = sS + b * HALO_STG + row_off;
                    const uint32_t t_addr = tmem_base + acc * (MBc * NT) + i * NT + g * 64 + (static_cast<uint32_t>(ew * 32) << 16);
                    {
                        uint32_t v[32];
#ifdef HALO_PROF
                        if (p.dbg & 2) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) v[e] = 0u;
                        } else
#endif
                        {
                            tmem_ld32(t_addr + hh * 32, v);
                            tmem_ld_wait();
                        }
                        if (i == nsub - 1 && g == NT / 64 - 1) {   // accumulator fully read: release it to the MMA warp
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) {
                                if (CG == 2 && rank != 0) mbar_arrive_cluster(tempty_remote[acc]);
                                else mbar_arrive(&tempty[acc]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float fv[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) fv[e] = __uint_as_float(v[j * 8 + e]);
                            const int col = g * 64 + hh * 32 + j * 8;
                            if (has_bias) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) fv[e] += sBias[col + e];
                            }
                            uint4* sp = reinterpret_cast<uint4*>(stg + (((static_cast<uint32_t>(hh * 4 + j)) ^ xr) << 4));
                            if (p.aux_mode == 1) {
                                const uint4 a = *sp;
                                const uint32_t aa[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 t = unpack_bf16x2(aa[e]);
                                    fv[2 * e] += t.x;
                                    fv[2 * e + 1] += t.y;
                                }
                            }
                            if (p.act == 1) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) fv[e] = fmaxf(fv[e], 0.f);
                            } else if (p.act == 2) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) fv[e] = fv[e] > 0.f ? fv[e] : 0.2f * fv[e];
                            }
                            if (p.aux_mode >= 2) {
                                const uint4 a = *sp;
                                const uint32_t aa[4] = {a.x, a.y, a.z, a.w};
                                const float neg = (p.aux_mode == 3) ? 0.2f : 0.f;
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 t = unpack_bf16x2(aa[e]);
                                    if (!(t.x > 0.f)) fv[2 * e] *= neg;
                                    if (!(t.y > 0.f)) fv[2 * e + 1] *= neg;
                                }
                            }
                            uint4 o;
                            o.x = pack_bf16x2(fv[0], fv[1]);
                            o.y = pack_bf16x2(fv[2], fv[3]);
                            o.z = pack_bf16x2(fv[4], fv[5]);
                            o.w = pack_bf16x2(fv[6], fv[7]);
#ifdef HALO_PROF
                            if (!(p.dbg & 1))
#endif
                            *sp = o;
                        }
                    }
                    fence_proxy_async();
#ifdef HALO_PROF
                    const long long tb0 = clock64();
#endif
                    named_bar_sync(1, HALO_EPI_THREADS);
#ifdef HALO_PROF
                    t_bar += clock64() - tb0;
                    const long long ts0 = clock64();
#endif
                    if (threadIdx.x == 128) {
                        const int n0 = nt * NT + g * 64;
                        const void* src = sS + b * HALO_STG;
#ifdef HALO_PROF
                        if (p.dbg & 1) {
                        } else
#endif
                        if (!store_ok) {
                            // duplicate tile of an odd last pair: nothing to write
                        } else if (p.out_mode == 0) {
                            tma_store_5d(&tmO, src, n0, tc.x0, tc.y0, tc.f, 0);
                        } else if (p.out_mode == 1) {
                            tma_store_5d(&tmO, src, n0, 0, tc.x0 >> 1, 0, tc.f * (p.H >> 1) + (tc.y0 >> 1));
                        } else if (p.cq == 32) {
                            // 64 staged channels = the two horizontally adjacent 32-channel output pixels of row parity n0/64
                            tma_store_5d(&tmO, src, 0, tc.x0, n0 >> 6, tc.f * p.H + tc.y0, 0);
                        } else {
                            const int q = n0 / p.cq, ch0 = n0 - q * p.cq;
                            tma_store_5d(&tmO, src, ch0, q & 1, tc.x0, q >> 1, tc.f * p.H + tc.y0);
                        }
                        bulk_commit();
                        if (cs >= 1) {                       // the previous chunk's store has finished reading its buffer
                            bulk_wait_read<1>();
                            mbar_arrive(&sempty[(cs - 1) % HALO_NST]);
                        }
                    }
#ifdef HALO_PROF
                    t_store += clock64() - ts0;
#endif
                }
            }
            if (++acc == static_cast<uint32_t>(p.NACC)) { acc = 0; acc_par ^= 1; }
        }
#ifdef HALO_PROF
        PROF_ADD(t_all);
        if (blockIdx.x == 0 && threadIdx.x == 128)
            printf("halo prof epilogue: total %lld  wait tfull %lld  wait staging %lld  barrier %lld  store+wait_read %lld  chunks %u\n",
                   t_all, w_tf, w_st, t_bar, t_store, cs);
#endif
        // the staging tiles must outlive the TMA stores' READS only; kernel completion makes the writes visible
        if (threadIdx.x == 128) bulk_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();     // no CTA leaves (or frees TMEM) while its peer may still signal / read it
    if (warp == 2) {
        if (CG == 2) tmem_dealloc_cg2(tmem_base, p.tmem_cols);
        else tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

template <int NT, bool RES, int CG>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmX,
                       const HaloParams& p, int smem_bytes, cudaStream_t st) {
    static bool configured = false;
    static int slots = 0;      // co-resident CTAs (CG == 1) or CTA pairs (CG == 2)
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<NT, RES, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             HALO_SMEM_MAX);
        if (e != cudaSuccess) return fail(P2I_ERR_CUDA, "conv_halo smem attribute: %s", cudaGetErrorString(e));
        slots = sm_count() / CG;
        if (CG == 2) {
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(sm_count());
            q.blockDim = dim3(HALO_THREADS);
            q.dynamicSmemBytes = HALO_SMEM_MAX;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            q.attrs = qa;
            q.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, conv_halo_kernel<NT, RES, CG>, &q) == cudaSuccess && n > 0 && n < slots) slots = n;
            (void)cudaGetLastError();
        }
        configured = true;
    }
    // persistent grid: a multiple of n_tiles so that every CTA (pair) keeps one N tile
    int groups = p.units * p.n_tiles;
    if (groups > slots) groups = (slots / p.n_tiles) * p.n_tiles;
    if (groups < p.n_tiles) return fail(P2I_ERR_INVALID, "conv_halo: %d N tiles do not fit the device", p.n_tiles);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(groups * CG);
    cfg.blockDim = dim3(HALO_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    // programmatic dependent launch: the prologue and the weight fetch overlap the stream predecessor's tail (ptx.cuh)
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_halo_kernel<NT, RES, CG>, tmA, tmB, tmO, tmX, p);
    if (e != cudaSuccess) return fail(P2I_ERR_CUDA, "conv_halo launch: %s", cudaGetErrorString(e));
    P2I_CHECK_LAUNCH("conv_halo_kernel");
    return P2I_OK;
}

static std::atomic<int> g_halo_cg{0};   // 0 auto, 1 single CTA, 2 CTA pairs (measurements)
int set_halo_cg(int cg) { g_halo_cg.store(cg); return P2I_OK; }

// Returns P2I_OK after launching, a negative error, or +1 when the shape is not eligible (caller falls back).
int run_igemm_halo(const void* x, const void* w, const P2iConvDesc& d, const void* residual, const void* mask,
                   const float* bias, void* y, void* stream) {
    if (d.Cin % 64 != 0 || d.Cout % 64 != 0 || d.ksize < 1 || d.ksize > 3 || (d.kt != 1 && d.kt != 3)) return 1;
    if (residual && mask) return 1;
    if (d.out_mode != 0 && (d.H % 16 != 0 || d.W % 8 != 0)) return 1;
    if (d.out_mode == 2 && (d.Cout / 4) % 64 != 0 && d.Cout != 128) return 1;   // unpack: 64-channel multiples or exactly 32
    HaloParams p;
    p.F = d.samples * d.T_out; p.T_out = d.T_out; p.T_in = d.T_in;
    p.H = d.H; p.W = d.W; p.Cin = d.Cin; p.Cout = d.Cout;
    p.KT = d.kt; p.KS = d.ksize;
    p.pad = d.pad; p.pad_t = d.pad_t; p.st = d.stride_t; p.tmode = d.t_transposed;
    p.tiles_x = cdiv(d.W, 8);
    p.tiles_y = cdiv(d.H, 16);
    p.m_tiles = p.F * p.tiles_x * p.tiles_y;
    p.act = d.act;
    p.aux_mode = residual ? 1 : (mask ? (d.mask_mode == 2 ? 3 : 2) : 0);
    p.out_mode = d.out_mode;
    p.cq = d.Cout / 4;
    p.dbg = 0;
#ifdef HALO_PROF
    if (const char* e = getenv("P2I_HALO_DBG")) p.dbg = atoi(e);
#endif
    p.bias = bias;

    const int cblocks = d.Cin / 64;
    const int nblk = d.kt * d.ksize * d.ksize * cblocks;
    const int fixed = 1024 /*alignment slack*/ + HALO_NST * HALO_STG + 1024 /*bias*/ + 512 /*barriers*/;
    const int budget = HALO_SMEM_MAX - fixed;
    const int sms = sm_count();

    // CTA pairs unless a transposed-temporal unit would mix frame parities inside one pair
    int CG = g_halo_cg.load(std::memory_order_relaxed);
    const bool pair_ok = p.m_tiles >= 2 && (d.t_transposed == 0 || (p.tiles_x * p.tiles_y) % 2 == 0);
    if (CG == 0) {
        // measured (profiles/r1_conv_ab.txt): pairs win everywhere except 64-output-channel layers whose weights stay
        // resident in a single CTA (N = 64 leaves nothing to share and the pair's cross-CTA barriers cost ~15 %)
        const bool res1 = d.Cout == 64 && nblk * 8192 + 2 * HALO_A_STRIDE <= budget;
        CG = (pair_ok && !res1) ? 2 : 1;
    }
    if (CG == 2 && !pair_ok) CG = 1;

    int NT;
    bool res;
    int b_bytes;
    if (CG == 2) {
        NT = (d.Cout % 256 == 0) ? 256 : ((d.Cout % 128 == 0) ? 128 : 64);
        if (NT == 256 && cdiv(p.m_tiles, 2) * (d.Cout / 256) < (sms / 2) * 6 / 10) NT = 128;   // too few pairs: narrower tiles
        const int blk = (NT / 2) * 128;
        res = nblk * blk + 2 * HALO_A_STRIDE <= budget;
        p.MB = 1;
        p.SB = res ? 1 : (NT == 256 ? 6 : 8);
        b_bytes = res ? nblk * blk : p.SB * blk;
        p.units = cdiv(p.m_tiles, 2);
    } else {
        res = nblk * 8192 + 2 * HALO_A_STRIDE <= budget && d.Cout == 64;
        if (res) {
            NT = 64;
            p.MB = 1;
            p.SB = 1;
            b_bytes = nblk * 8192;
        } else {
            NT = (d.Cout % 128 == 0) ? 128 : 64;
            p.SB = (NT == 128) ? 4 : 8;
            b_bytes = p.SB * NT * 128;
            // M blocking: halves the weight traffic per FLOP but doubles the work quantum; take it when the SMs stay busy
            const int n_tiles = d.Cout / NT;
            const long long waves1 = cdiv(p.m_tiles * n_tiles, sms);
            const long long waves2 = cdiv(cdiv(p.m_tiles, 2) * n_tiles, sms);
            p.MB = (d.t_transposed == 0 && p.m_tiles >= 2 && waves2 * 2 * 14 <= waves1 * 18) ? 2 : 1;
        }
        p.units = cdiv(p.m_tiles, p.MB);
    }
    p.n_tiles = d.Cout / NT;
    p.SA = (budget - b_bytes) / HALO_A_STRIDE;
    // Ring depth cap (P2I_HALO_SA_MAX, default 8): every stage not taken is 23 KB of shared memory that a concurrent CUDA-core
    // kernel's CTAs (bias column sums, the thin first 3-D layer, Adam ...) can use to co-reside with this kernel's CTA.
    static const int sa_max = [] { const char* e = getenv("P2I_HALO_SA_MAX"); const int v = e ? atoi(e) : 8; return v < 2 ? 2 : (v > 8 ? 8 : v); }();
    if (p.SA > sa_max) p.SA = sa_max;
    if (p.SA < 2 * p.MB) return 1;
    p.NACC = (p.MB * NT * 2 <= 512) ? 2 : 1;
    const int cols = p.NACC * p.MB * NT;
    p.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
    p.offB = static_cast<uint32_t>(p.SA * HALO_A_STRIDE);
    p.offS = p.offB + static_cast<uint32_t>(b_bytes);
    p.offBias = p.offS + HALO_NST * HALO_STG;
    p.offBar = p.offBias + 1024;
    const int smem_bytes = static_cast<int>(p.offBar) + 512 + 1024;

    CUtensorMap tmA, tmB, tmO, tmX;
    {
        const uint64_t C = d.Cin, W = d.W, H = d.H, T = d.T_in;
        const uint64_t dims[5] = {C, W, H, T, uint64_t(d.samples)};
        const uint64_t strides[5] = {0, C * 2, W * C * 2, H * W * C * 2, T * H * W * C * 2};
        const uint32_t box[5] = {64, uint32_t(8 + d.ksize - 1), uint32_t(16 + d.ksize - 1), 1, 1};
        int rc = encode_tmap_bf16(&tmA, x, 5, dims, strides, box, nullptr, true);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {uint64_t(d.Cin), uint64_t(d.Cout), uint64_t(d.kt * d.ksize * d.ksize)};
        const uint64_t strides[3] = {0, uint64_t(d.Cin) * 2, uint64_t(d.Cout) * d.Cin * 2};
        const uint32_t box[3] = {64, uint32_t(NT / CG), 1};
        int rc = encode_tmap_bf16(&tmB, w, 3, dims, strides, box, nullptr, true);
        if (rc) return rc;
    }
    const uint64_t Co = d.Cout, Wd = d.W, Hd = d.H, Fr = uint64_t(p.F);
    {
        // the kernel's own output grid [F, H, W, Cout] (layout of the residual / mask tensors)
        const uint64_t dims[5] = {Co, Wd, Hd, Fr, 1};
        const uint64_t strides[5] = {0, Co * 2, Wd * Co * 2, Hd * Wd * Co * 2, Fr * Hd * Wd * Co * 2};
        const uint32_t box[5] = {64, 8, 16, 1, 1};
        const void* aux = residual ? residual : (mask ? mask : y);
        int rc = encode_tmap_bf16(&tmX, aux, 5, dims, strides, box, nullptr, true);
        if (rc) return rc;
        if (d.out_mode == 0) {
            rc = encode_tmap_bf16(&tmO, y, 5, dims, strides, box, nullptr, true);
            if (rc) return rc;
        }
    }
    if (d.out_mode == 1) {
        // y [F, H/2, W/2, 4, Cout]: pixel (y,x) -> [y/2][x/2][(y&1)*2 + (x&1)]
        const uint64_t dims[5] = {Co, 2, Wd / 2, 2, Fr * (Hd / 2)};
        const uint64_t strides[5] = {0, Co * 2, 4 * Co * 2, 2 * Co * 2, (Wd / 2) * 4 * Co * 2};
        const uint32_t box[5] = {64, 2, 4, 2, 8};
        int rc = encode_tmap_bf16(&tmO, y, 5, dims, strides, box, nullptr, true);
        if (rc) return rc;
    } else if (d.out_mode == 2 && d.Cout == 128) {
        // y [F, 2H, 2W, 32]: channels n = q*32 + ch, q = 2*dy + dx.  The 64 channels (dx = 0, 1) of one dy are the 128
        // contiguous bytes of output pixels (2y+dy, 2x), (2y+dy, 2x+1):  dims {64, W, 2 (dy), F*H}
        const uint64_t dims[5] = {64, Wd, 2, Fr * Hd, 1};
        const uint64_t strides[5] = {0, 128, Wd * 128, 2 * Wd * 128, 2 * Wd * 128 * Fr * Hd};
        const uint32_t box[5] = {64, 8, 1, 16, 1};
        int rc = encode_tmap_bf16(&tmO, y, 5, dims, strides, box, nullptr, true);
        if (rc) return rc;
    } else if (d.out_mode == 2) {
        // y [F, 2H, 2W, C], C = Cout/4: channel n = q*C + ch -> pixel (2y + q/2, 2x + q%2)
        const uint64_t C = Co / 4;
        const uint64_t dims[5] = {C, 2, Wd, 2, Fr * Hd};
        const uint64_t strides[5] = {0, C * 2, 2 * C * 2, 2 * Wd * C * 2, 4 * Wd * C * 2};
        const uint32_t box[5] = {64, 1, 8, 1, 16};
        int rc = encode_tmap_bf16(&tmO, y, 5, dims, strides, box, nullptr, true);
        if (rc) return rc;
    }
    cudaStream_t st = as_stream(stream);
#define HALO_LAUNCH(NT_, RES_, CG_)                                                          \
    do {                                                                                     \
        set_last_variant(2000000 + p.MB * 100000 + NT_ * 100 + (RES_ ? 10 : 0) + CG_);       \
        return launch_halo<NT_, RES_, CG_>(tmA, tmB, tmO, tmX, p, smem_bytes, st);           \
    } while (0)
    if (CG == 2) {
        if (NT == 256) { if (res) HALO_LAUNCH(256, true, 2); else HALO_LAUNCH(256, false, 2); }
        if (NT == 128) { if (res) HALO_LAUNCH(128, true, 2); else HALO_LAUNCH(128, false, 2); }
        if (res) HALO_LAUNCH(64, true, 2); else HALO_LAUNCH(64, false, 2);
    }
    if (res) HALO_LAUNCH(64, true, 1);
    if (NT == 128) HALO_LAUNCH(128, false, 1);
    HALO_LAUNCH(64, false, 1);
#undef HALO_LAUNCH
}

}  // namespace p2i
