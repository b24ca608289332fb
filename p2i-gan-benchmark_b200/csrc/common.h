// Host-side helpers shared by every translation unit of libp2i_sm100a.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/p2i_b200.h"

namespace p2i {

extern std::atomic<long long> g_launches;

// Thread-local last-error string, exposed through p2i_last_error().
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define P2I_CHECK_ARG(cond, ...)                                   \
    do {                                                           \
        if (!(cond)) return ::p2i::fail(P2I_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define P2I_CHECK_LAUNCH(name)                                                            \
    do {                                                                                  \
        ::p2i::g_launches.fetch_add(1, std::memory_order_relaxed);                        \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) return ::p2i::fail(P2I_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// cuTensorMapEncodeTiled resolved at run time (no link-time dependency on libcuda).
// dims/strides innermost first; strides in BYTES for dims 1..rank-1.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* elem_strides, bool swizzle128);

int sm_count();

// Debug record of the kernel variant the calling thread launched last (p2i_conv_last_variant; tests assert that the
// instantiation under test is the one the benchmark runs).
void set_last_variant(int code);

// Programmatic dependent launch of the tensor-core kernels (p2i_set_pdl; default off, P2I_PDL=1 in the environment turns it on).
bool pdl_enabled();

// d3d_first_mma.cu: tensor-core (mma.sync) forward / weight gradient of the one-input-channel Conv3d d3d.0
bool d3d_first_mma_ok(int T, int H, int W);
int d3d_first_fwd_mma(const float* x, const float* w, const float* sigma, const float* bias, void* y, int B, int T, int H, int W,
                      cudaStream_t stream);
int d3d_first_bwd_w_mma(const void* dpre, const float* x, float* dW, float* db, int B, int T, int H, int W, cudaStream_t stream);

}  // namespace p2i
