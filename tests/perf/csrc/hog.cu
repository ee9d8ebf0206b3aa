// Development tool (tests/perf/coresidency.py): a kernel that only OCCUPIES shared memory / threads on every SM for a given time,
// to measure how many CTAs of another kernel (the fused gather) become resident next to a persistent 1-CTA-per-SM kernel
// (the parameter GEMM) as a function of the shared memory it leaves free.
#include <cuda_runtime.h>
#include <stdint.h>

extern "C" __global__ void __launch_bounds__(192, 1) hog_kernel(long long cycles, int* sink) {
    extern __shared__ uint8_t hog_smem[];
    const long long t0 = clock64();
    if (threadIdx.x == 0) hog_smem[0] = 1;
    while (clock64() - t0 < cycles) { }
    if (sink != nullptr && threadIdx.x == 0 && hog_smem[0] == 77) sink[0] = 1;
}

extern "C" int hog_launch(int ctas, int smem_bytes, long long cycles, void* stream) {
    static int configured = -1;
    if (configured != smem_bytes) {
        if (cudaFuncSetAttribute(hog_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return (int)cudaGetLastError();
        configured = smem_bytes;
    }
    hog_kernel<<<ctas, 192, smem_bytes, (cudaStream_t)stream>>>(cycles, nullptr);
    return (int)cudaGetLastError();
}
