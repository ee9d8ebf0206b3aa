// Shared helpers for the sparsebev_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/sparsebev_b200.h"

namespace sbev {

void set_error(const char* fmt, ...);
enum { OPT_GEMM_IMPL = 0, OPT_MIX_IMPL = 1, OPT_SASA_IMPL = 2, OPT_GATHER_VARIANT = 3, OPT_DENSE_IMPL = 4, OPT_DENSE_CLUSTER = 5, OPT_COUNT = 6 };
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
