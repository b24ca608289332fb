// Experiment: can a 128B-swizzled K-major UMMA operand start at a row (128 B) offset inside the 1024-B swizzle atom,
// and can 8-row groups be 1280 B apart?  (Needed for a single halo box per conv tile instead of one box per kx tap.)
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I p2i-gan-benchmark_b200/csrc \
//          tools/exp_umma_shift.cu p2i-gan-benchmark_b200/csrc/common.cu -o gpurun_out/exp_umma_shift -lcuda
#include <math.h>
#include <stdlib.h>
#include <vector>

#include "common.h"
#include "ptx.cuh"

using namespace p2i;

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>(1) << 16;                             // LBO (unused for SW128 K-major)
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(base_off & 7) << 49;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// A: P pixels x 64 ch in smem (pixel p at p*128, TMA SW128).  B: 64 x 64.  D[128][64].
__global__ void __launch_bounds__(128, 1) exp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                     float* out, int a_pixels, int start_rows, int sbo, int base_off) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + 65536;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 8192);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, a_pixels * 128 + 8192);
        tma_load_3d(sA, &tmA, bar, 0, 0, 0);
        tma_load_3d(sB, &tmB, bar, 0, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_bf16(128, 64);
        const uint32_t a0 = smem_u32(sA) + start_rows * 128, b0 = smem_u32(sB);
        for (int k = 0; k < 4; ++k)
            umma_bf16(tm, make_desc(a0 + k * 32, sbo, base_off), make_desc(b0 + k * 32, 1024, 0), idesc, k > 0);
        umma_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    const uint32_t ta = tm + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c = 0; c < 64; c += 16) {
        uint32_t v[16];
        tmem_ld16(ta + c, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + c + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 64);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
    const int P = 18 * 16;   // up to 288 pixels resident
    std::vector<float> A(P * 64), B(64 * 64);
    srand(1);
    for (auto& v : A) v = bf((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : B) v = bf((rand() % 2001 - 1000) / 1000.f);
    std::vector<__nv_bfloat16> Ah(P * 64), Bh(64 * 64);
    for (int i = 0; i < P * 64; ++i) Ah[i] = __float2bfloat16(A[i]);
    for (int i = 0; i < 64 * 64; ++i) Bh[i] = __float2bfloat16(B[i]);
    __nv_bfloat16 *dA, *dB;
    float* dO;
    cudaMalloc(&dA, Ah.size() * 2); cudaMalloc(&dB, Bh.size() * 2); cudaMalloc(&dO, 128 * 64 * 4);
    cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice);
    // full product: G[p][n] = sum_k A[p][k] B[n][k]
    std::vector<float> G(P * 64);
    for (int p = 0; p < P; ++p)
        for (int n = 0; n < 64; ++n) {
            double s = 0;
            for (int k = 0; k < 64; ++k) s += double(A[p * 64 + k]) * B[n * 64 + k];
            G[p * 64 + n] = float(s);
        }
    CUtensorMap tmA, tmB;
    {
        // A as [P/16 rows][16 px][64 ch]; box = whole thing, 256 px max per dim -> 3-D {64,16,18}
        const uint64_t dims[3] = {64, 16, uint64_t(P / 16)};
        const uint64_t strides[3] = {0, 128, 16 * 128};
        const uint32_t box[3] = {64, 16, uint32_t(P / 16)};
        if (encode_tmap_bf16(&tmA, dA, 3, dims, strides, box, nullptr, true)) { printf("tmA failed\n"); return 1; }
        const uint64_t dimsb[3] = {64, 64, 1};
        const uint64_t stridesb[3] = {0, 128, 64 * 128};
        const uint32_t boxb[3] = {64, 64, 1};
        if (encode_tmap_bf16(&tmB, dB, 3, dimsb, stridesb, boxb, nullptr, true)) { printf("tmB failed\n"); return 1; }
    }
    cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80000);
    std::vector<float> O(128 * 64);
    const int sbos[3] = {1024, 1280, 2048};
    for (int si = 0; si < 3; ++si)
        for (int shift = 0; shift < 11; ++shift)
            for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
                const int sbo = sbos[si];
                const int bo = bo_mode ? (shift & 7) : 0;
                if (bo_mode == 1 && bo == 0) continue;
                cudaMemset(dO, 0, 128 * 64 * 4);
                exp_kernel<<<1, 128, 80000>>>(tmA, tmB, dO, P, shift, sbo, bo);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("sbo %d shift %d bo %d: CUDA error %s\n", sbo, shift, bo, cudaGetErrorString(e)); return 2; }
                cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
                // hypothesis: row m reads pixel shift + (m/8)*(sbo/128) + m%8
                int bad = 0, firstbad = -1;
                for (int m = 0; m < 128; ++m) {
                    const int p = shift + (m / 8) * (sbo / 128) + (m % 8);
                    for (int n = 0; n < 64; ++n)
                        if (fabsf(O[m * 64 + n] - G[p * 64 + n]) > 2e-2f) { ++bad; if (firstbad < 0) firstbad = m; }
                }
                // if wrong, which pixel does each row match (first 16 rows)?
                printf("sbo %4d shift %2d base_off %d : %s (bad %d, first bad row %d)", sbo, shift, bo, bad ? "MISMATCH" : "ok", bad, firstbad);
                if (bad) {
                    printf("  row->pixel:");
                    for (int m = 0; m < 20; ++m) {
                        int found = -1;
                        for (int p = 0; p < P && found < 0; ++p) {
                            bool okp = true;
                            for (int n = 0; n < 64 && okp; ++n) okp = fabsf(O[m * 64 + n] - G[p * 64 + n]) <= 2e-2f;
                            if (okp) found = p;
                        }
                        printf(" %d", found);
                    }
                }
                printf("\n");
            }
    return 0;
}
