// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16 operands, both in shared memory) on every SM at once, as a
// function of the tile shape:  M = 128 (one CTA) or 256 (CTA pair, cta_group::2),  N = 64 / 128 / 256,  with the A
// operand's 8-row groups 1024 B apart (canonical) or 1280 B apart (the conv kernel's halo-box tap views), operands
// rotating over 3 shared-memory K blocks.  Prints cycles per K=16 instruction and the MAC rate per SM, i.e. how far the
// shared-memory operand fetch (A: 128*32 B, B: N/CG*32 B per instruction and CTA) limits each shape.  
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I p2i-gan-benchmark_b200/csrc tools/umma_rate.cu -o tools/build/umma_rate
#include <algorithm>
#include <vector>

#include "ptx.cuh"

using namespace p2i;

constexpr int A_BLK = 24576, B_BLK = 32768, ROT = 3;

// bg_warps: extra warps that stream 16-byte shared-memory loads + stores (a separate 16 KB region) while the MMAs run -- the
// epilogue / TMA staging traffic of the real kernel -- and report how many bytes they moved.  taps: the A descriptor start
// cycles over the nine (ky*10 + kx)*128-byte tap offsets of the conv kernel's halo box instead of staying 1024-B aligned.
template <int CG>
__global__ void __launch_bounds__(384, 1) rate_kernel(int N, int iters, int sbo_a, int taps, int bg_warps, long long* cycles,
                                                      long long* bg_bytes) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + ROT * A_BLK;
    uint8_t* sBG = sB + ROT * B_BLK;                       // 16 KB of background traffic
    uint64_t* done = reinterpret_cast<uint64_t*>(sBG + 16384);
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
    volatile int* stop = reinterpret_cast<volatile int*>(slot + 1);
    if (threadIdx.x == 0) *stop = 0;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    // small pseudo-random bf16 values (realistic toggling, no NaN / Inf)
    for (int i = threadIdx.x; i < (ROT * (A_BLK + B_BLK)) / 2; i += blockDim.x) {
        const uint32_t h = (i * 2654435761u + blockIdx.x * 40503u) >> 20;
        reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16((static_cast<int>(h & 255) - 128) * (1.0f / 256.0f));
    }
    fence_proxy_async();
    if (threadIdx.x == 0) { mbar_init(done, 1); fence_mbar_init(); }
    if (warp == 0) {
        if (CG == 2) { tmem_alloc_cg2(slot, 512); tmem_relinquish_cg2(); } else { tmem_alloc(slot, 512); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t tm = *slot;
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t idesc = make_idesc_bf16(128 * CG, N);
        const uint64_t a_hi = (static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>((sbo_a >> 4) & 0x3FFF) << 32) |
                              (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
        const uint64_t b_hi = (static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
                              (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        const long long t0 = clock64();
        int r = 0;
        int tap = 0;
        for (int it = 0; it < iters; ++it) {
            const uint32_t toff = taps ? static_cast<uint32_t>(((tap / 3) * 10 + tap % 3) * 128) : 0u;
            if (++tap == 9) tap = 0;
            const uint64_t ad = a_hi | static_cast<uint64_t>(((a0 + r * A_BLK + toff) >> 4) & 0x3FFF);
            const uint64_t bd = b_hi | static_cast<uint64_t>(((b0 + r * B_BLK) >> 4) & 0x3FFF);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (CG == 2) umma_bf16_cg2(tm, ad + 2 * k, bd + 2 * k, idesc, (it | k) ? 1u : 0u);
                else umma_bf16(tm, ad + 2 * k, bd + 2 * k, idesc, (it | k) ? 1u : 0u);
            }
            if (++r == ROT) r = 0;
        }
        if (CG == 2) umma_commit_cg2(done); else umma_commit(done);
        mbar_wait(done, 0);
        cycles[blockIdx.x] = clock64() - t0;
        *stop = 1;
    } else if (warp >= 4 && warp < 4 + bg_warps) {
        // background shared-memory traffic until the MMAs are done (the non-leader CTA of a pair stops on its own barrier copy)
        uint4* q = reinterpret_cast<uint4*>(sBG) + (threadIdx.x & 31) + ((warp - 4) & 7) * 32 * 4;
        long long n = 0;
        uint4 v = make_uint4(1, 2, 3, 4);
        while (!*stop && n < (1ll << 22)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { q[k * 32] = v; }
#pragma unroll
            for (int k = 0; k < 4; ++k) { const uint4 t = q[k * 32]; v.x ^= t.y; v.y += t.x; }
            n += 8;
        }
        if ((threadIdx.x & 31) == 0) atomicAdd(reinterpret_cast<unsigned long long*>(bg_bytes + blockIdx.x), static_cast<unsigned long long>(n) * 512ull + (v.x == 0x7fffffffu));
    } else if (threadIdx.x == 0) {
        mbar_wait(done, 0);          // the peer's copy of the multicast commit
        cycles[blockIdx.x] = 0;
        *stop = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) {
        if (CG == 2) tmem_dealloc_cg2(tm, 512); else tmem_dealloc(tm, 512);
    }
}

template <int CG>
static void run(int N, int sbo, int taps, int bg_warps, int sms, long long* d_cyc) {
    const int iters = 2048;
    const int smem = 1024 + ROT * (A_BLK + B_BLK) + 16384 + 64;
    long long* d_bg = d_cyc + 256;
    cudaFuncSetAttribute(rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms / CG * CG);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    std::vector<long long> h(sms);
    double best = 0, bgb = 0;
    std::vector<long long> hb(sms);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(d_cyc, 0, 512 * sizeof(long long));
        cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<CG>, N, iters, sbo, taps, bg_warps, d_cyc, d_bg);
        if (e != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) {
            printf("CG=%d N=%d sbo=%d: %s\n", CG, N, sbo, cudaGetErrorString(e));
            return;
        }
        cudaMemcpy(h.data(), d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaMemcpy(hb.data(), d_bg, sms * sizeof(long long), cudaMemcpyDeviceToHost);
        std::vector<double> v;
        for (int i = 0; i < sms; ++i) if (h[i] > 0) v.push_back(double(h[i]) / (iters * 4));
        std::sort(v.begin(), v.end());
        best = v[v.size() / 2];
        double tot = 0, cyc = 0;
        for (int i = 0; i < sms; ++i) if (h[i] > 0) { tot += double(hb[i]); cyc += double(h[i]); }
        bgb = cyc > 0 ? tot / cyc : 0;      // leaders' CTAs only
    }
    const double mac_per_clk_sm = 128.0 * N * 16 / best;            // per SM (a pair's instruction covers 2 SMs)
    const double bytes = 128 * 32 + (N / CG) * 32;                   // operand bytes fetched per instruction and CTA
    printf("CG=%d M=%3d N=%3d sboA=%4d taps=%d bg_warps=%d : %7.2f clk per K=16 MMA  %6.0f MAC/clk/SM = %5.1f %% of 4096   operand fetch "
           "%5.1f B/clk/SM   background ld+st %5.1f B/clk/SM\n", CG, 128 * CG, N, sbo, taps, bg_warps, best, mac_per_clk_sm,
           100.0 * mac_per_clk_sm / 4096.0, bytes / best, bgb);
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* d_cyc;
    cudaMalloc(&d_cyc, 512 * sizeof(long long));
    for (int N : {64, 128, 256}) {
        run<1>(N, 1024, 0, 0, sms, d_cyc);
        run<2>(N, 1024, 0, 0, sms, d_cyc);
    }
    // the conv kernel's A views: tap-shifted starts inside a halo box, 8-row groups 1280 B apart
    for (int N : {64, 128}) {
        run<1>(N, 1280, 1, 0, sms, d_cyc);
        run<2>(N, 1280, 1, 0, sms, d_cyc);
    }
    // with concurrent shared-memory traffic of other warps (epilogue staging, TMA tiles)
    for (int bg : {1, 2, 4, 8}) {
        run<1>(64, 1280, 1, bg, sms, d_cyc);
        run<2>(64, 1280, 1, bg, sms, d_cyc);
        run<2>(128, 1280, 1, bg, sms, d_cyc);
    }
    return 0;
}
