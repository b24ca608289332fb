// Error reporting, launch counter and the TMA descriptor encoder shared by all kernels.
#include <atomic>
#include <mutex>
#include <stdlib.h>
#include <string.h>

#include "common.h"

namespace p2i {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
static thread_local int g_last_variant = 0;
void set_last_variant(int code) { g_last_variant = code; }
int last_variant() { return g_last_variant; }
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
    int v = g_pdl.load(std::memory_order_relaxed);
    if (v < 0) {
        // default OFF: measured neutral on the training step (6.44 vs 6.40 ms, profiles/r2_pdl_ab.txt) -- every kernel keeps
        // one CTA per SM, so a dependent CTA only starts where its predecessor's CTA has already exited -- and an early-started
        // kernel's waiting CTAs (220 KB of shared memory each) keep other streams' kernels off those SMs
        const char* e = getenv("P2I_PDL");
        v = (e && e[0] == '1') ? 1 : 0;
        g_pdl.store(v);
    }
    return v != 0;
}
void set_pdl(int on) { g_pdl.store(on ? 1 : 0); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* elem_strides, bool swizzle128) {
    EncodeTiledFn fn = resolve_encode();
    if (!fn) return fail(P2I_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gd[5];
    cuuint64_t gs[4];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gd[i] = dims[i];
        bx[i] = box[i];
        es[i] = elem_strides ? elem_strides[i] : 1;
        if (i > 0) gs[i - 1] = strides_bytes[i];
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gd,
                    gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(P2I_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
    return P2I_OK;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

}  // namespace p2i

extern "C" {
int p2i_abi_version(void) { return 1; }
const char* p2i_last_error(void) { return p2i::g_err; }
long long p2i_launch_count(void) { return p2i::g_launches.load(); }
int p2i_conv_last_variant(void) { return p2i::last_variant(); }
int p2i_set_pdl(int on) { p2i::set_pdl(on); return P2I_OK; }
}
