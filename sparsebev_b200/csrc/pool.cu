// Non-convolution pieces of the VoVNet-99-eSE image branch (BASELINE config 5), NHWC bf16 on the device:
//   * 3x3 stride-2 max pool with explicit padding / ceil_mode   (/root/reference/models/backbones/vovnet.py:228: the OSA stages
//     start with nn.MaxPool2d(kernel_size=3, stride=2, ceil_mode=True); mmdet's ResNet stem uses padding 1, floor)
//   * effective squeeze-excitation (vovnet.py:166-178): global average pool -> 1x1 conv (a [C,C] mat-vec per image) ->
//     hard sigmoid relu6(x + 3) / 6 -> channel-wise scale of the block output (+ the OSA identity, vovnet.py:217-218)
// All memory-bound elementwise / reduction work: coalesced 16-byte accesses, deterministic two-stage reduction (no float atomics).
#include "common.cuh"
#include <cuda_bf16.h>

namespace sbev {

__global__ void __launch_bounds__(256)
maxpool3x3s2_ex_kernel(const __nv_bfloat16* __restrict__ x, int Nimg, int H, int W, int C, int pad, int Ho, int Wo, __nv_bfloat16* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int c8 = C / 8;
    const long long total = (long long)Nimg * Ho * Wo * c8;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int cg = (int)(i % c8);
        const long long p = i / c8;
        const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
        float m[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = -3.0e38f;
        for (int dh = 0; dh < 3; ++dh)
            for (int dw = 0; dw < 3; ++dw) {
                const int hh = 2 * ho - pad + dh, ww = 2 * wo - pad + dw;
                if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;          // padding / ceil_mode overhang: ignored (-inf)
                const uint4 r = __ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + hh) * W + ww) * C + cg * 8));
                const uint32_t r4[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r4[e]));
                    m[2 * e] = fmaxf(m[2 * e], f.x); m[2 * e + 1] = fmaxf(m[2 * e + 1], f.y);
                }
            }
        uint32_t w4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(m[2 * e], m[2 * e + 1]);
            w4[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(out + p * C + cg * 8) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
}

constexpr int GAP_PIX = 128;          // pixels per partial sum

// partial[n][chunk][c] = sum over the chunk's pixels of x[n][pixel][c]      (grid: chunks x Nimg)
__global__ void __launch_bounds__(256)
gap_partial_kernel(const __nv_bfloat16* __restrict__ x, int HW, int C, int chunks, float* __restrict__ partial) {
    pdl_wait();
    pdl_trigger();
    const int chunk = blockIdx.x, n = blockIdx.y;
    const int p0 = chunk * GAP_PIX, p1 = min(HW, p0 + GAP_PIX);
    for (int c2 = threadIdx.x; c2 < C / 2; c2 += 256) {
        float a = 0.f, b = 0.f;
        const __nv_bfloat162* src = reinterpret_cast<const __nv_bfloat162*>(x + ((long long)n * HW + p0) * C) + c2;
        for (int p = p0; p < p1; ++p, src += C / 2) {
            const float2 f = __bfloat1622float2(__ldg(src));
            a += f.x; b += f.y;
        }
        *reinterpret_cast<float2*>(partial + ((long long)n * chunks + chunk) * C + 2 * c2) = make_float2(a, b);
    }
}

// gate[n][o] = relu6(bias[o] + sum_c W[o][c] * mean[n][c] + 3) / 6      (grid: C/32 x Nimg; warp w of 8 -> outputs 4w..4w+3 of the block's 32)
__global__ void __launch_bounds__(256)
ese_gate_kernel(const float* __restrict__ partial, int chunks, int HW, int C, const float* __restrict__ Wfc, const float* __restrict__ bfc,
                float* __restrict__ gate) {
    extern __shared__ float mean_s[];          // [C]
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += 256) {
        float s = 0.f;
        for (int k = 0; k < chunks; ++k) s += partial[((long long)n * chunks + k) * C + c];          // fixed order: deterministic
        mean_s[c] = s / (float)HW;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = 0; j < 4; ++j) {
        const int o = blockIdx.x * 32 + warp * 4 + j;
        if (o >= C) break;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(__ldg(Wfc + (long long)o * C + c), mean_s[c], s);
        s = warp_sum(s);
        if (lane == 0) gate[(long long)n * C + o] = fminf(fmaxf(s + __ldg(bfc + o) + 3.f, 0.f), 6.f) / 6.f;
    }
}

// out = x * gate[n][c] (+ identity), NHWC bf16, 8 channels per thread
__global__ void __launch_bounds__(256)
ese_scale_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gate, const __nv_bfloat16* __restrict__ identity,
                 long long total8, int HW, int C, __nv_bfloat16* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int c8 = C / 8;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total8; i += (long long)gridDim.x * 256) {
        const int cg = (int)(i % c8);
        const long long pix = i / c8;
        const int n = (int)(pix / HW);
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(x) + i);
        uint4 idv = make_uint4(0u, 0u, 0u, 0u);
        if (identity != nullptr) idv = __ldg(reinterpret_cast<const uint4*>(identity) + i);
        const float4 g0 = ldg4(gate + (long long)n * C + cg * 8), g1 = ldg4(gate + (long long)n * C + cg * 8 + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const uint32_t r4[4] = {r.x, r.y, r.z, r.w}, i4[4] = {idv.x, idv.y, idv.z, idv.w};
        uint32_t w4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r4[e]));
            const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&i4[e]));
            const __nv_bfloat162 h = __floats2bfloat162_rn(f.x * g[2 * e] + d.x, f.y * g[2 * e + 1] + d.y);
            w4[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        reinterpret_cast<uint4*>(out)[i] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
}

}  // namespace sbev

using namespace sbev;

extern "C" int sbev_maxpool3x3s2_ex_nhwc_fwd(const uint16_t* x, int Nimg, int H, int W, int C, int pad, int ceil_mode, uint16_t* out, void* stream) {
    SBEV_REQUIRE(x && out, SBEV_ERR_INVALID, "sbev_maxpool3x3s2_ex_nhwc_fwd: null pointer");
    SBEV_REQUIRE(C % 8 == 0 && C > 0 && H > 0 && W > 0 && Nimg >= 0 && (pad == 0 || pad == 1), SBEV_ERR_UNSUPPORTED,
                 "sbev_maxpool3x3s2_ex_nhwc_fwd: C must be a multiple of 8, pad 0 or 1");
    SBEV_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, SBEV_ERR_INVALID, "sbev_maxpool3x3s2_ex_nhwc_fwd: operands must be 16-byte aligned");
    if (Nimg == 0) return SBEV_OK;
    // torch.nn.MaxPool2d output size: floor|ceil((size + 2*pad - 3) / 2) + 1; in ceil mode the last window must start inside the input or its left padding
    auto osz = [&](int s) {
        const int num = s + 2 * pad - 3;
        if (num < 0) return 0;
        int o = (ceil_mode ? (num + 1) / 2 : num / 2) + 1;
        if (ceil_mode && (o - 1) * 2 >= s + pad) --o;
        return o;
    };
    const int Ho = osz(H), Wo = osz(W);
    SBEV_REQUIRE(Ho > 0 && Wo > 0, SBEV_ERR_INVALID, "sbev_maxpool3x3s2_ex_nhwc_fwd: input smaller than the window");
    const long long total = (long long)Nimg * Ho * Wo * (C / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_pdl(maxpool3x3s2_ex_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(x), Nimg, H, W, C, pad, Ho, Wo,
               reinterpret_cast<__nv_bfloat16*>(out));
    return check_launch("sbev_maxpool3x3s2_ex_nhwc_fwd");
}

extern "C" long long sbev_ese_workspace_floats(int Nimg, int H, int W, int C) {
    if (Nimg < 0 || H <= 0 || W <= 0 || C <= 0) return -1;
    const long long chunks = ((long long)H * W + GAP_PIX - 1) / GAP_PIX;
    return (long long)Nimg * chunks * C + (long long)Nimg * C;
}

extern "C" int sbev_ese_nhwc_fwd(const uint16_t* x, int Nimg, int H, int W, int C, const float* fc_weight, const float* fc_bias,
                                 const uint16_t* identity, float* workspace, uint16_t* out, void* stream) {
    SBEV_REQUIRE(x && fc_weight && fc_bias && workspace && out, SBEV_ERR_INVALID, "sbev_ese_nhwc_fwd: null pointer");
    SBEV_REQUIRE(C % 8 == 0 && C > 0 && C <= 4096 && H > 0 && W > 0 && Nimg >= 0, SBEV_ERR_UNSUPPORTED, "sbev_ese_nhwc_fwd: C must be a multiple of 8, <= 4096");
    SBEV_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(identity) |
                   reinterpret_cast<uintptr_t>(workspace)) & 15) == 0, SBEV_ERR_INVALID, "sbev_ese_nhwc_fwd: operands must be 16-byte aligned");
    if (Nimg == 0) return SBEV_OK;
    const int HW = H * W, chunks = (HW + GAP_PIX - 1) / GAP_PIX;
    float* partial = workspace;
    float* gate = workspace + (long long)Nimg * chunks * C;
    cudaStream_t st = (cudaStream_t)stream;
    launch_pdl(gap_partial_kernel, dim3(chunks, Nimg), dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(x), HW, C, chunks, partial);
    launch_pdl(ese_gate_kernel, dim3((C + 31) / 32, Nimg), dim3(256), (size_t)C * sizeof(float), st, (const float*)partial, chunks, HW, C, fc_weight, fc_bias, gate);
    const long long total8 = (long long)Nimg * HW * (C / 8);
    long long blocks = (total8 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_pdl(ese_scale_kernel, dim3((unsigned)blocks), dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(x), (const float*)gate,
               reinterpret_cast<const __nv_bfloat16*>(identity), total8, HW, C, reinterpret_cast<__nv_bfloat16*>(out));
    return check_launch("sbev_ese_nhwc_fwd");
}
