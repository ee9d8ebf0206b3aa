// Shared helpers for the sparsebev_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/sparsebev_b200.h"

namespace sbev {

void set_error(const char* fmt, ...);
enum { OPT_GEMM_IMPL = 0, OPT_MIX_IMPL = 1, OPT_SASA_IMPL = 2, OPT_GATHER_VARIANT = 3, OPT_DENSE_IMPL = 4, OPT_DENSE_CLUSTER = 5, OPT_PDL = 6, OPT_DENSE_NSPLIT = 7, OPT_DENSE_VEC4 = 8, OPT_DENSE_FUSE_POINTS = 9, OPT_MIX_ORDER = 10, OPT_LEGACY_ROTATION = 11, OPT_DENSE_PACK = 12, OPT_SASA_KQ = 13, OPT_DENSE_WS = 14, OPT_DENSE_WS_GROUPS = 15, OPT_GEMM_L2_HINTS = 16, OPT_GATHER_L2_HINT = 17, OPT_COUNT = 18 };
// 2-D bf16 row-major [rows, cols] tensor map, box = [box_rows, 64 cols], 128 B swizzle, zero OOB fill (gemm_tcgen05.cu)
int make_bf16_map(CUtensorMap* out, const void* ptr, long long rows, long long cols, int box_rows);
int make_bf16_map_ex(CUtensorMap* out, const void* ptr, long long rows, long long cols, int box_rows, int box_cols, int swizzle_bytes);
int get_option(int id);
int set_option(const char* name, int value);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return SBEV_ERR_CUDA;
    }
    return SBEV_OK;
}

// Programmatic dependent launch (PDL).  Every hot-path kernel runs its global-memory-free prologue (barrier init, TMEM
// allocation, descriptor prefetch, index math), then pdl_wait() -- which returns once the preceding kernel of the stream
// has completed and flushed -- and immediately pdl_trigger(), so the NEXT kernel of the stream may be scheduled as soon as
// every CTA of this one is running: its launch latency and prologue hide under this kernel's tail.  No kernel touches
// global memory before its pdl_wait(), so the usual stream-order semantics are preserved (RAW, WAR and WAW alike).
// Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Launch `kernel` with the programmatic-stream-serialization attribute (option "pdl", default on).  ONLY for kernels
// that call pdl_wait() before their first global access.
template <typename... P, typename... A>
inline void launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = get_option(OPT_PDL) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);      // errors are picked up by check_launch()
}

// One-time per-DEVICE setup (cudaFuncSetAttribute opt-ins are per device, not per process: a process that drives
// cuda:0 and then cuda:1 must set them on both).  `first_use()` is true once for every device ordinal; a race between
// two host threads only repeats an idempotent call.
struct DeviceOnce {
    unsigned long long mask[4] = {0, 0, 0, 0};          // 256 device ordinals
    bool first_use() {
        int dev = 0;
        cudaGetDevice(&dev);
        const unsigned long long bit = 1ull << (dev & 63);
        unsigned long long& m = mask[(dev >> 6) & 3];
        if (m & bit) return false;
        m |= bit;
        return true;
    }
};
#define SBEV_PER_DEVICE_ONCE(...)                      \
    do {                                               \
        static ::sbev::DeviceOnce once_per_device_;    \
        if (once_per_device_.first_use()) { __VA_ARGS__; } \
    } while (0)

// SM count of the CURRENT device (cached per device ordinal).
inline int device_num_sms() {
    static int cache[256] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& n = cache[dev & 255];
    if (n == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        n = v > 0 ? v : 148;
    }
    return n;
}

#define SBEV_REQUIRE(cond, code, ...)            \
    do {                                         \
        if (!(cond)) {                           \
            ::sbev::set_error(__VA_ARGS__);      \
            return (code);                       \
        }                                        \
    } while (0)

__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace sbev
