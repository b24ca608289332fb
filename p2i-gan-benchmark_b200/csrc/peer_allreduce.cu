// Data-parallel gradient exchange over NVLink peer memory: a two-shot all-reduce(sum) of one flat fp32
// buffer, written as ONE kernel per rank that loads / stores the peers' buffers directly (CUDA IPC mappings), so that
// the whole training step -- including its two gradient exchanges -- is a single CUDA graph with no NCCL call inside.
// (The reference has no multi-GPU path; torch DDP's bucketed NCCL all-reduce is what this replaces.)
//
//   shard r = elements [r*S, (r+1)*S) of the buffer, owned by rank r; block b of every rank works on sub-range b of
//   every shard, so all cross-GPU dependencies are between equally numbered blocks (block-level flag barriers, no
//   grid-wide synchronisation):
//     barrier A   block b of every peer has STARTED this kernel (=> the peer's backward kernels are complete)
//     phase 1     reduce-scatter: own shard, sub-range b  =  sum over ranks 0..N-1 in that fixed order (bitwise the same
//                 value whoever computes it); loads bypass L1 (ld.global.cg), result written in place
//     barrier B   block b of every peer has finished phase 1
//     phase 2     all-gather: copy sub-range b of every other shard from its owner
//     barrier C   block b of every peer has finished reading my shard (the next step may overwrite the buffer)
//   Flags are monotonically increasing epochs (no reset, graph-replay safe); st.release.sys / ld.acquire.sys; every spin
//   is bounded (~10 s) and reports through an error word instead of hanging the GPU.
#include "common.h"
#include "ptx.cuh"

namespace p2i {

constexpr int PEER_MAX_RANKS = 8;
constexpr int PEER_MAX_BLOCKS = 320;
constexpr int PEER_THREADS = 256;

struct PeerTable {
    float* buf[PEER_MAX_RANKS];      // buf[r]: rank r's flat buffer as mapped in THIS process (buf[rank] is local)
    int* flags[PEER_MAX_RANKS];      // flags[r]: rank r's flag array int[3][PEER_MAX_BLOCKS][PEER_MAX_RANKS]
    int* epoch;                      // local: incremented by peer_tick_kernel before every all-reduce
    int* err;                        // local: set to 1 when a spin timed out
    long long n;                     // elements of the exchanged range (multiple of 4)
    long long off4;                  // first float4 of the range inside the flat buffers (bucketed exchange)
    int rank, world;
};

__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void peer_tick_kernel(int* epoch) { *epoch += 1; }

// all threads call; threads t < world signal peer t and wait for peer t
__device__ __forceinline__ void peer_barrier(const PeerTable& a, int phase, int epoch) {
    __syncthreads();
    const int t = threadIdx.x;
    if (t < a.world) {
        __threadfence_system();
        st_release_sys(a.flags[t] + (phase * PEER_MAX_BLOCKS + blockIdx.x) * PEER_MAX_RANKS + a.rank, epoch);
        const int* mine = a.flags[a.rank] + (phase * PEER_MAX_BLOCKS + blockIdx.x) * PEER_MAX_RANKS + t;
        const long long t0 = clock64();
        while (ld_acquire_sys(mine) < epoch) {
            if (clock64() - t0 > 20000000000ll) { *a.err = 1; break; }     // ~10 s at 1.9 GHz
        }
    }
    __syncthreads();
}

// WT = world size known at compile time (2, 4, 8: the peer loop is fully unrolled so that the loads from ALL peers are in
// flight together -- with a run-time trip count each peer's load waited for the previous peer's add) or 0 = generic.
// Register budget: a bucket exchange runs UNDER the backward pass, next to a tensor-core CTA that holds up to ~40 k of an
// SM's 64 k registers (384 threads x 104) -- 256 threads x <= 64 registers fit beside it, 512 x 122 (round 1's WT = 8 form)
// would keep the conv CTA off the SM for as long as the exchange spins.  Hence 256 threads, minBlocks = 4 (<= 64 registers) and fewer
// independent loads per peer at larger world sizes (what is in flight per thread stays 8 x 16 bytes).
template <int WT>
__global__ void __launch_bounds__(PEER_THREADS, 4) peer_allreduce_kernel(const PeerTable a) {
    const int epoch = *a.epoch;
    const int W = WT ? WT : a.world, G = gridDim.x;
    const long long n4 = a.n >> 2;
    const long long shard4 = (n4 + W - 1) / W;                 // float4 per shard
    const long long per4 = (shard4 + G - 1) / G;               // float4 per (shard, block)
    const long long lo = static_cast<long long>(blockIdx.x) * per4;
    const long long hi = (lo + per4 < shard4) ? lo + per4 : shard4;
    constexpr int U = WT == 8 ? 1 : (WT == 4 ? 2 : 4);          // independent 16-byte loads per peer and thread

    peer_barrier(a, 0, epoch);
    {
        const long long s0 = static_cast<long long>(a.rank) * shard4;
        float4* dst = reinterpret_cast<float4*>(a.buf[a.rank]) + a.off4;
        for (long long i = lo + threadIdx.x; i < hi; i += U * PEER_THREADS) {
            bool ok[U];
            long long j[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                j[u] = s0 + i + u * PEER_THREADS;
                ok[u] = i + u * PEER_THREADS < hi && j[u] < n4;
            }
            float4 acc[U];
            if (WT) {
                float4 v[WT ? WT : 1][U];
#pragma unroll
                for (int r = 0; r < WT; ++r)
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        v[r][u] = ok[u] ? __ldcg(reinterpret_cast<const float4*>(a.buf[r]) + a.off4 + j[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    acc[u] = v[0][u];
#pragma unroll
                    for (int r = 1; r < WT; ++r) { acc[u].x += v[r][u].x; acc[u].y += v[r][u].y; acc[u].z += v[r][u].z; acc[u].w += v[r][u].w; }
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int r = 0; r < W; ++r) {
                    const float4* src = reinterpret_cast<const float4*>(a.buf[r]) + a.off4;
                    float4 v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) v[u] = ok[u] ? __ldcg(src + j[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < U; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u]) dst[j[u]] = acc[u];
        }
    }
    peer_barrier(a, 1, epoch);
    {
        float4* dst = reinterpret_cast<float4*>(a.buf[a.rank]) + a.off4;
        if (WT) {
            // all W-1 owners' sub-ranges in flight together
            for (long long i = lo + threadIdx.x; i < hi; i += U * PEER_THREADS) {
                float4 v[WT ? WT : 1][U];
#pragma unroll
                for (int k = 1; k < WT; ++k) {
                    const int s = (a.rank + k) % (WT ? WT : 1);
                    const float4* src = reinterpret_cast<const float4*>(a.buf[s]) + a.off4;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const long long j = static_cast<long long>(s) * shard4 + i + u * PEER_THREADS;
                        v[k][u] = (i + u * PEER_THREADS < hi && j < n4) ? __ldcg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int k = 1; k < WT; ++k) {
                    const int s = (a.rank + k) % (WT ? WT : 1);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const long long j = static_cast<long long>(s) * shard4 + i + u * PEER_THREADS;
                        if (i + u * PEER_THREADS < hi && j < n4) dst[j] = v[k][u];
                    }
                }
            }
        } else {
            for (int k = 1; k < W; ++k) {
                const int s = (a.rank + k) % W;                    // start with different owners on different ranks
                const long long s0 = static_cast<long long>(s) * shard4;
                const float4* src = reinterpret_cast<const float4*>(a.buf[s]) + a.off4;
                for (long long i = lo + threadIdx.x; i < hi; i += U * PEER_THREADS) {
                    float4 v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const long long j = s0 + i + u * PEER_THREADS;
                        v[u] = (i + u * PEER_THREADS < hi && j < n4) ? __ldcg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const long long j = s0 + i + u * PEER_THREADS;
                        if (i + u * PEER_THREADS < hi && j < n4) dst[j] = v[u];
                    }
                }
            }
        }
    }
    peer_barrier(a, 2, epoch);
}

}  // namespace p2i

using namespace p2i;

extern "C" int p2i_peer_alloc(void** out, long long bytes) {
    P2I_CHECK_ARG(out && bytes > 0, "peer_alloc: bad arguments");
    cudaError_t e = cudaMalloc(out, static_cast<size_t>(bytes));
    if (e != cudaSuccess) return fail(P2I_ERR_CUDA, "peer_alloc: %s", cudaGetErrorString(e));
    e = cudaMemset(*out, 0, static_cast<size_t>(bytes));
    if (e != cudaSuccess) return fail(P2I_ERR_CUDA, "peer_alloc memset: %s", cudaGetErrorString(e));
    return P2I_OK;
}

extern "C" int p2i_peer_free(void* p) {
    cudaError_t e = cudaFree(p);
    if (e != cudaSuccess) return fail(P2I_ERR_CUDA, "peer_free: %s", cudaGetErrorString(e));
    return P2I_OK;
}

extern "C" int p2i_peer_export(const void* p, void* handle64) {
    P2I_CHECK_ARG(p && handle64, "peer_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(p));
    if (e != cudaSuccess) return fail(P2I_ERR_CUDA, "peer_export: %s", cudaGetErrorString(e));
    return P2I_OK;
}

extern "C" int p2i_peer_import(const void* handle64, void** out) {
    P2I_CHECK_ARG(handle64 && out, "peer_import: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(P2I_ERR_CUDA, "peer_import: %s", cudaGetErrorString(e));
    return P2I_OK;
}

extern "C" int p2i_peer_close(void* p) {
    cudaError_t e = cudaIpcCloseMemHandle(p);
    if (e != cudaSuccess) return fail(P2I_ERR_CUDA, "peer_close: %s", cudaGetErrorString(e));
    return P2I_OK;
}

extern "C" int p2i_peer_flags_bytes(void) { return 3 * PEER_MAX_BLOCKS * PEER_MAX_RANKS * static_cast<int>(sizeof(int)); }

static int peer_allreduce_launch(void* const* bufs, void* const* flags, int rank, int world, long long offset, long long n,
                                 int blocks, int* epoch_dev, int* err_dev, void* stream) {
    P2I_CHECK_ARG(bufs && flags && epoch_dev && err_dev, "peer_allreduce: null pointer");
    P2I_CHECK_ARG(world >= 1 && world <= PEER_MAX_RANKS && rank >= 0 && rank < world, "peer_allreduce: bad rank/world %d/%d", rank, world);
    P2I_CHECK_ARG(n > 0 && n % 4 == 0 && offset >= 0 && offset % 4 == 0, "peer_allreduce: offset and n must be multiples of 4 (n > 0)");
    PeerTable a;
    for (int r = 0; r < PEER_MAX_RANKS; ++r) {
        a.buf[r] = r < world ? static_cast<float*>(bufs[r]) : nullptr;
        a.flags[r] = r < world ? static_cast<int*>(flags[r]) : nullptr;
        P2I_CHECK_ARG(r >= world || (a.buf[r] && a.flags[r]), "peer_allreduce: null peer pointer for rank %d", r);
    }
    a.epoch = epoch_dev; a.err = err_dev; a.n = n; a.off4 = offset >> 2; a.rank = rank; a.world = world;
    int grid = blocks > 0 ? blocks : 2 * sm_count();          // whole-buffer exchange: two 256-thread CTAs per SM
    if (grid > PEER_MAX_BLOCKS) grid = PEER_MAX_BLOCKS;
    peer_tick_kernel<<<1, 1, 0, as_stream(stream)>>>(epoch_dev);
    P2I_CHECK_LAUNCH("peer_tick_kernel");
    if (world == 2) peer_allreduce_kernel<2><<<grid, PEER_THREADS, 0, as_stream(stream)>>>(a);
    else if (world == 4) peer_allreduce_kernel<4><<<grid, PEER_THREADS, 0, as_stream(stream)>>>(a);
    else if (world == 8) peer_allreduce_kernel<8><<<grid, PEER_THREADS, 0, as_stream(stream)>>>(a);
    else peer_allreduce_kernel<0><<<grid, PEER_THREADS, 0, as_stream(stream)>>>(a);
    P2I_CHECK_LAUNCH("peer_allreduce_kernel");
    return P2I_OK;
}

extern "C" int p2i_peer_allreduce(void* const* bufs, void* const* flags, int rank, int world, long long n, int* epoch_dev,
                                  int* err_dev, void* stream) {
    return peer_allreduce_launch(bufs, flags, rank, world, 0, n, 0, epoch_dev, err_dev, stream);
}

extern "C" int p2i_peer_allreduce_range(void* const* bufs, void* const* flags, int rank, int world, long long offset,
                                        long long n, int blocks, int* epoch_dev, int* err_dev, void* stream) {
    return peer_allreduce_launch(bufs, flags, rank, world, offset, n, blocks, epoch_dev, err_dev, stream);
}
