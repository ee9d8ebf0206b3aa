// AdaptiveMixing, stage 2 of 3: the per-(query, group) dynamic mixing
//     h = relu(LN_{Pin x C}(x @ M));   y = relu(LN_{Pout x C}(S @ h))
// between the two tensor-core GEMMs (parameter generation, output projection; see gemm_tcgen05.cu).
//
// Behavioural reference: /root/reference/models/sparsebev_transformer.py:358-375.
//   params row layout per (query, group): [ M: C x C row-major | S: Pout x Pin row-major ]
//   x [BQ, G, Pin, C];  y flattened as [BQ, G*Pout*C] = the out_proj GEMM's A operand.
//
// One CTA (256 threads) per (query, group): M, x and S^T are staged in shared memory, both small
// matmuls are register-tiled fp32 FFMA (exact fp32 like the reference), both two-dimensional
// LayerNorms are two-pass block reductions on register-resident tiles, and the result leaves as the
// bf16 (hi, lo) pair the bf16x3 out_proj GEMM consumes -- the same bytes as one fp32 copy.
#include "common.cuh"
#include <cuda_bf16.h>

namespace sbev {

constexpr int MIX_C = 64;
constexpr int MIX_POUT = 128;
constexpr int MIX_ST_LD = MIX_POUT + 4;     // S^T row stride (keeps 16 B alignment, spreads banks)

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();                 // red[] free to overwrite
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = red[lane & 7];
    t += __shfl_xor_sync(0xffffffffu, t, 4);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    return t;
}

// RPT = rows of h per thread in stage 1 (Pin <= 16*RPT)
template <int RPT>
__global__ void __launch_bounds__(256)
mix_kernel(const float* __restrict__ params, const float* __restrict__ x, int G, int Pin,
           __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo, float* __restrict__ y_f32) {
    extern __shared__ __align__(16) float smem[];
    float* Ms = smem;                              // [64][64]
    float* xs = Ms + MIX_C * MIX_C;                // [Pin][64]
    float* St = xs + Pin * MIX_C;                  // [Pin][132]  (S transposed)
    float* hs = St + Pin * MIX_ST_LD;              // [Pin][64]
    __shared__ float red[8];

    const int tid = threadIdx.x;
    const long long qg = blockIdx.x;               // bq*G + g
    const long long per_group = (long long)MIX_C * MIX_C + (long long)MIX_POUT * Pin;
    const float* pm = params + qg * per_group;
    const float* ps = pm + MIX_C * MIX_C;
    const float* px = x + qg * Pin * MIX_C;

    for (int i = tid * 4; i < MIX_C * MIX_C; i += 1024)
        *reinterpret_cast<float4*>(Ms + i) = ldg4(pm + i);
    for (int i = tid * 4; i < Pin * MIX_C; i += 1024)
        *reinterpret_cast<float4*>(xs + i) = ldg4(px + i);
    for (int i = tid; i < MIX_POUT * Pin; i += 256) {
        const int o = i / Pin, p = i - o * Pin;
        St[p * MIX_ST_LD + o] = __ldg(ps + i);
    }
    __syncthreads();

    const int tx = tid & 15, ty = tid >> 4;

    // ---- stage 1: h[p][c'] = sum_c x[p][c] M[c][c'],  rows p = ty + 16*i, cols 4tx..4tx+3
    float4 h[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) h[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int c = 0; c < MIX_C; c += 4) {
        const float4 m0 = *reinterpret_cast<const float4*>(Ms + (c + 0) * MIX_C + 4 * tx);
        const float4 m1 = *reinterpret_cast<const float4*>(Ms + (c + 1) * MIX_C + 4 * tx);
        const float4 m2 = *reinterpret_cast<const float4*>(Ms + (c + 2) * MIX_C + 4 * tx);
        const float4 m3 = *reinterpret_cast<const float4*>(Ms + (c + 3) * MIX_C + 4 * tx);
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int p = ty + 16 * i;
            if (p < Pin) {
                const float4 xv = *reinterpret_cast<const float4*>(xs + p * MIX_C + c);
#define SBEV_MIX_FMA(acc, s, m) acc.x = fmaf(s, m.x, acc.x); acc.y = fmaf(s, m.y, acc.y); acc.z = fmaf(s, m.z, acc.z); acc.w = fmaf(s, m.w, acc.w);
                SBEV_MIX_FMA(h[i], xv.x, m0) SBEV_MIX_FMA(h[i], xv.y, m1) SBEV_MIX_FMA(h[i], xv.z, m2) SBEV_MIX_FMA(h[i], xv.w, m3)
            }
        }
    }
    // LayerNorm over the whole Pin x 64 tile (no affine, eps 1e-5), then ReLU
    {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < RPT; ++i) if (ty + 16 * i < Pin) s += (h[i].x + h[i].y) + (h[i].z + h[i].w);
        const float n = (float)(Pin * MIX_C);
        const float mean = block_sum_256(s, red) / n;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < RPT; ++i) if (ty + 16 * i < Pin) {
            const float a = h[i].x - mean, b = h[i].y - mean, c = h[i].z - mean, d = h[i].w - mean;
            ss += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd = rsqrtf(block_sum_256(ss, red) / n + 1e-5f);
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int p = ty + 16 * i;
            if (p < Pin) {
                float4 o;
                o.x = fmaxf((h[i].x - mean) * rstd, 0.f); o.y = fmaxf((h[i].y - mean) * rstd, 0.f);
                o.z = fmaxf((h[i].z - mean) * rstd, 0.f); o.w = fmaxf((h[i].w - mean) * rstd, 0.f);
                *reinterpret_cast<float4*>(hs + p * MIX_C + 4 * tx) = o;
            }
        }
    }
    __syncthreads();

    // ---- stage 2: y[o][c'] = sum_p S[o][p] h[p][c'],  rows o = 8ty..8ty+7, cols 4tx..4tx+3
    float4 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int p = 0; p < Pin; ++p) {
        const float4 hv = *reinterpret_cast<const float4*>(hs + p * MIX_C + 4 * tx);
        const float4 s0 = *reinterpret_cast<const float4*>(St + p * MIX_ST_LD + 8 * ty);
        const float4 s1 = *reinterpret_cast<const float4*>(St + p * MIX_ST_LD + 8 * ty + 4);
        SBEV_MIX_FMA(acc[0], s0.x, hv) SBEV_MIX_FMA(acc[1], s0.y, hv) SBEV_MIX_FMA(acc[2], s0.z, hv) SBEV_MIX_FMA(acc[3], s0.w, hv)
        SBEV_MIX_FMA(acc[4], s1.x, hv) SBEV_MIX_FMA(acc[5], s1.y, hv) SBEV_MIX_FMA(acc[6], s1.z, hv) SBEV_MIX_FMA(acc[7], s1.w, hv)
#undef SBEV_MIX_FMA
    }
    {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += (acc[i].x + acc[i].y) + (acc[i].z + acc[i].w);
        const float n = (float)(MIX_POUT * MIX_C);
        const float mean = block_sum_256(s, red) / n;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float a = acc[i].x - mean, b = acc[i].y - mean, c = acc[i].z - mean, d = acc[i].w - mean;
            ss += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd = rsqrtf(block_sum_256(ss, red) / n + 1e-5f);
        const long long obase = qg * (MIX_POUT * MIX_C);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int o = 8 * ty + i;
            float v[4] = {fmaxf((acc[i].x - mean) * rstd, 0.f), fmaxf((acc[i].y - mean) * rstd, 0.f),
                          fmaxf((acc[i].z - mean) * rstd, 0.f), fmaxf((acc[i].w - mean) * rstd, 0.f)};
            const long long off = obase + o * MIX_C + 4 * tx;
            if (y_f32) *reinterpret_cast<float4*>(y_f32 + off) = make_float4(v[0], v[1], v[2], v[3]);
            if (y_hi) {
                __nv_bfloat16 hh[4], ll[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { hh[k] = __float2bfloat16_rn(v[k]); ll[k] = __float2bfloat16_rn(v[k] - __bfloat162float(hh[k])); }
                *reinterpret_cast<uint2*>(y_hi + off) = *reinterpret_cast<const uint2*>(hh);
                if (y_lo) *reinterpret_cast<uint2*>(y_lo + off) = *reinterpret_cast<const uint2*>(ll);
            }
        }
    }
}

}  // namespace sbev

using namespace sbev;

extern "C" int sbev_mix_fwd(const float* params, const float* x, int BQ, int G, int Pin, int Pout, int C,
                            uint16_t* y_hi, uint16_t* y_lo, float* y_f32, void* stream) {
    SBEV_REQUIRE(params && x && (y_hi || y_f32), SBEV_ERR_INVALID, "sbev_mix_fwd: null pointer");
    SBEV_REQUIRE(C == MIX_C && Pout == MIX_POUT, SBEV_ERR_UNSUPPORTED, "sbev_mix_fwd: needs C=64, out_points=128 (got %d, %d)", C, Pout);
    SBEV_REQUIRE(Pin >= 1 && Pin <= 128, SBEV_ERR_UNSUPPORTED, "sbev_mix_fwd: in_points must be in [1,128] (got %d)", Pin);
    SBEV_REQUIRE(BQ >= 0 && G > 0, SBEV_ERR_INVALID, "sbev_mix_fwd: bad sizes");
    SBEV_REQUIRE((((long long)MIX_C * MIX_C + (long long)Pout * Pin) & 3) == 0, SBEV_ERR_UNSUPPORTED, "sbev_mix_fwd: in_points must be a multiple of 4... (C*C + Pout*Pin) % 4 != 0");
    if (BQ == 0) return SBEV_OK;
    const size_t smem = sizeof(float) * ((size_t)MIX_C * MIX_C + (size_t)Pin * MIX_C * 2 + (size_t)Pin * MIX_ST_LD);
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(y_hi);
    __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(y_lo);
    const int grid = BQ * G;
#define SBEV_LAUNCH_MIX(R)                                                                                         \
    do {                                                                                                           \
        if (smem > 48 * 1024) cudaFuncSetAttribute(mix_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        mix_kernel<R><<<grid, 256, smem, st>>>(params, x, G, Pin, hi, lo, y_f32);                                  \
    } while (0)
    if (Pin <= 32) SBEV_LAUNCH_MIX(2);
    else if (Pin <= 64) SBEV_LAUNCH_MIX(4);
    else SBEV_LAUNCH_MIX(8);
#undef SBEV_LAUNCH_MIX
    return check_launch("sbev_mix_fwd");
}
