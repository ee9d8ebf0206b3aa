// Small dense blocks of the decoder layer: Linear (+bias) (+residual) (+LayerNorm) (+ReLU),
// the sampling-head epilogue (box decode -> sample points, softmax over levels), the fp32 ->
// bf16 (hi, lo) split that feeds the tcgen05 GEMMs, and the split-K reduce + residual + LayerNorm
// that closes AdaptiveMixing.
//
// Behavioural reference (paths under /root/reference):
//   models/sparsebev_transformer.py:113-144 (position_encoder, cls/reg branches), :125 FFN,
//   :166-176 (layer forward), :262-263,279-283,298-299 (sampling head),
//   models/sparsebev_sampling.py:8-24, models/bbox/utils.py:63-77, models/utils.py:49-84.
//
// The 900 x 256 x 256 GEMMs here are latency-bound, not FLOP-bound (118 MFLOP each), so they stay
// on the fp32 FFMA pipe: one CTA owns 8 full rows (so LayerNorm never leaves the CTA), weights are
// pre-transposed [K, ldw] so every lane streams 16 B coalesced from L2, 113 CTAs cover Q=900.
#include "common.cuh"
#include <cuda_bf16.h>

namespace sbev {

constexpr int DENSE_ROWS = 8;

// y[M,N] = epilogue(x[M,K] @ Wt[K,ldw])   (Wt = W^T, zero-padded to ldw = multiple of 4)
__global__ void __launch_bounds__(256)
dense_rows8_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ Wt, int ldw,
                   const float* __restrict__ bias, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                   const float* __restrict__ residual, int M, int K, int N, int flags, float* __restrict__ y) {
    extern __shared__ float smem[];
    const int Kp = (K + 3) & ~3;
    float* xs = smem;                         // [8][Kp]
    float* ys = smem + DENSE_ROWS * Kp;       // [8][ldw]
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * DENSE_ROWS;
    for (int i = tid; i < DENSE_ROWS * Kp; i += 256) {
        const int r = i / Kp, k = i - r * Kp;
        xs[i] = (row0 + r < M && k < K) ? __ldg(x + (long long)(row0 + r) * ldx + k) : 0.f;
    }
    __syncthreads();
    const int ncg = ldw >> 2;                 // column groups of 4
    const int rg = tid >> 6;                  // 0..3 -> rows 2rg, 2rg+1
    const float* x0 = xs + (2 * rg) * Kp;
    const float* x1 = x0 + Kp;
    for (int cg = tid & 63; cg < ncg; cg += 64) {
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        const float* wp = Wt + 4 * cg;
        int k = 0;
        for (; k + 4 <= K; k += 4) {
            const float4 w0 = ldg4(wp + (long long)(k + 0) * ldw);
            const float4 w1 = ldg4(wp + (long long)(k + 1) * ldw);
            const float4 w2 = ldg4(wp + (long long)(k + 2) * ldw);
            const float4 w3 = ldg4(wp + (long long)(k + 3) * ldw);
            const float4 xa = *reinterpret_cast<const float4*>(x0 + k);
            const float4 xb = *reinterpret_cast<const float4*>(x1 + k);
#define SBEV_FMA4(acc, xv, wv) acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y); acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
            SBEV_FMA4(a0, xa.x, w0) SBEV_FMA4(a0, xa.y, w1) SBEV_FMA4(a0, xa.z, w2) SBEV_FMA4(a0, xa.w, w3)
            SBEV_FMA4(a1, xb.x, w0) SBEV_FMA4(a1, xb.y, w1) SBEV_FMA4(a1, xb.z, w2) SBEV_FMA4(a1, xb.w, w3)
        }
        for (; k < K; ++k) {
            const float4 w0 = ldg4(wp + (long long)k * ldw);
            SBEV_FMA4(a0, x0[k], w0) SBEV_FMA4(a1, x1[k], w0)
#undef SBEV_FMA4
        }
        *reinterpret_cast<float4*>(ys + (2 * rg) * ldw + 4 * cg) = a0;
        *reinterpret_cast<float4*>(ys + (2 * rg + 1) * ldw + 4 * cg) = a1;
    }
    __syncthreads();
    // epilogue: one warp per row
    const int warp = tid >> 5, lane = tid & 31;
    const int row = row0 + warp;
    if (row >= M) return;
    float* yr = ys + warp * ldw;
    const bool pre_res = (flags & SBEV_DENSE_RES_PRE_LN) && residual != nullptr;
    for (int n = lane; n < N; n += 32) {
        float v = yr[n];
        if (bias) v += __ldg(bias + n);
        if (pre_res) v += __ldg(residual + (long long)row * N + n);
        yr[n] = v;
    }
    float mean = 0.f, rstd = 1.f;
    if (ln_w != nullptr) {
        float s = 0.f;
        for (int n = lane; n < N; n += 32) s += yr[n];
        mean = warp_sum(s) / (float)N;
        float ss = 0.f;
        for (int n = lane; n < N; n += 32) { const float d = yr[n] - mean; ss += d * d; }
        rstd = rsqrtf(warp_sum(ss) / (float)N + 1e-5f);
    }
    for (int n = lane; n < N; n += 32) {
        float v = yr[n];
        if (ln_w != nullptr) v = (v - mean) * rstd * __ldg(ln_w + n) + __ldg(ln_b + n);
        if (flags & SBEV_DENSE_RELU) v = fmaxf(v, 0.f);
        if (!pre_res && residual != nullptr) v += __ldg(residual + (long long)row * N + n);
        y[(long long)row * N + n] = v;
    }
}

// Box decode + offsets -> lidar-frame sample points; softmax over levels.  Thread per point.
__global__ void __launch_bounds__(128)
sample_points_kernel(const float* __restrict__ query_bbox, const float* __restrict__ offset,
                     const float* __restrict__ logits, float r0, float r1, float r2, float r3, float r4, float r5,
                     int BQ, int GP, int L, float* __restrict__ points, float* __restrict__ scale_w) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)BQ * GP) return;
    const long long bq = idx / GP;
    const float* bb = query_bbox + bq * 10;
    // decode_bbox: xyz*(hi-lo)+lo, exp(log-size), atan2(sin, cos)   (separate mul and add as in torch)
    const float cx = __fadd_rn(__fmul_rn(__ldg(bb + 0), __fsub_rn(r3, r0)), r0);
    const float cy = __fadd_rn(__fmul_rn(__ldg(bb + 1), __fsub_rn(r4, r1)), r1);
    const float cz = __fadd_rn(__fmul_rn(__ldg(bb + 2), __fsub_rn(r5, r2)), r2);
    const float sw = expf(__ldg(bb + 3)), sl = expf(__ldg(bb + 4)), sh = expf(__ldg(bb + 5));
    const float yaw = atan2f(__ldg(bb + 6), __ldg(bb + 7));
    const float s = sinf(yaw), c = cosf(yaw);
    const float* op = offset + idx * 3;
    const float dx = __fmul_rn(sw, __ldg(op)), dy = __fmul_rn(sl, __ldg(op + 1)), dz = __fmul_rn(sh, __ldg(op + 2));
    // rotate counter-clockwise by yaw about z: x' = x*c + y*(-s), y' = x*s + y*c
    const float rx = __fadd_rn(__fmul_rn(dx, c), __fmul_rn(dy, -s));
    const float ry = __fadd_rn(__fmul_rn(dx, s), __fmul_rn(dy, c));
    points[idx * 3 + 0] = __fadd_rn(cx, rx);
    points[idx * 3 + 1] = __fadd_rn(cy, ry);
    points[idx * 3 + 2] = __fadd_rn(cz, dz);
    // softmax over L
    const float* lp = logits + idx * L;
    float mx = -INFINITY;
    for (int l = 0; l < L; ++l) mx = fmaxf(mx, __ldg(lp + l));
    float e[SBEV_MAX_LEVELS], sum = 0.f;
    for (int l = 0; l < L; ++l) { e[l] = expf(__ldg(lp + l) - mx); sum += e[l]; }
    for (int l = 0; l < L; ++l) scale_w[idx * L + l] = __fdiv_rn(e[l], sum);
}

__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ x, long long n, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 + 4 <= n) {
        const float4 v = *reinterpret_cast<const float4*>(x + i4);
        const float f[4] = {v.x, v.y, v.z, v.w};
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { h[k] = __float2bfloat16_rn(f[k]); l[k] = __float2bfloat16_rn(f[k] - __bfloat162float(h[k])); }
        *reinterpret_cast<uint2*>(hi + i4) = *reinterpret_cast<const uint2*>(h);
        if (lo) *reinterpret_cast<uint2*>(lo + i4) = *reinterpret_cast<const uint2*>(l);
    } else {
        for (long long i = i4; i < n; ++i) {
            const __nv_bfloat16 h = __float2bfloat16_rn(x[i]);
            hi[i] = h;
            if (lo) lo[i] = __float2bfloat16_rn(x[i] - __bfloat162float(h));
        }
    }
}

// out[row] = LN(sum_z partial[z][row] + bias + residual[row]); one warp per row, N <= 1024.
__global__ void __launch_bounds__(256)
reduce_ln_kernel(const float* __restrict__ partial, int nsplit, const float* __restrict__ bias,
                 const float* __restrict__ residual, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                 int M, int N, float* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= M) return;
    float v[32];
    const int per = (N + 31) / 32;
    float s = 0.f;
    for (int i = 0; i < per; ++i) {
        const int n = lane + 32 * i;
        float a = 0.f;
        if (n < N) {
            for (int z = 0; z < nsplit; ++z) a += __ldg(partial + ((long long)z * M + row) * N + n);
            if (bias) a += __ldg(bias + n);
            if (residual) a += __ldg(residual + (long long)row * N + n);
        }
        v[i] = a;
        s += (n < N) ? a : 0.f;
    }
    float mean = 0.f, rstd = 1.f;
    if (ln_w != nullptr) {
        mean = warp_sum(s) / (float)N;
        float ss = 0.f;
        for (int i = 0; i < per; ++i) { const int n = lane + 32 * i; if (n < N) { const float d = v[i] - mean; ss += d * d; } }
        rstd = rsqrtf(warp_sum(ss) / (float)N + 1e-5f);
    }
    for (int i = 0; i < per; ++i) {
        const int n = lane + 32 * i;
        if (n < N) {
            float a = v[i];
            if (ln_w != nullptr) a = (a - mean) * rstd * __ldg(ln_w + n) + __ldg(ln_b + n);
            out[(long long)row * N + n] = a;
        }
    }
}

// bbox refinement + velocity rescale (sparsebev_transformer.py:155-160,179-183):
// xyz = sigmoid(delta_xyz + inverse_sigmoid(proposal_xyz)); dims 3..9 taken raw; vel /= time_diff[b,1].
__global__ void __launch_bounds__(256)
refine_bbox_kernel(const float* __restrict__ proposal, const float* __restrict__ delta, const float* __restrict__ time_diff,
                   int B, int Q, int T, int code, float* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * Q * code) return;
    const int c = (int)(idx % code);
    const long long bq = idx / code;
    float v = __ldg(delta + idx);
    if (c < 3) {
        float x = fminf(fmaxf(__ldg(proposal + bq * code + c), 0.f), 1.f);
        const float x1 = fmaxf(x, 1e-5f), x2 = fmaxf(1.f - x, 1e-5f);
        v = v + logf(__fdiv_rn(x1, x2));
        v = __fdiv_rn(1.f, 1.f + expf(-v));
    } else if (c >= 8 && T > 1) {
        float td = __ldg(time_diff + (bq / Q) * T + 1);
        if (td < 1e-5f) td = 1.0f;
        v = __fdiv_rn(v, td);
    }
    out[idx] = v;
}

}  // namespace sbev

using namespace sbev;

extern "C" int sbev_dense_fwd(const float* x, int ldx, const float* Wt, int ldw, const float* bias,
                              const float* ln_w, const float* ln_b, const float* residual,
                              int M, int K, int N, int flags, float* y, void* stream) {
    SBEV_REQUIRE(x && Wt && y, SBEV_ERR_INVALID, "sbev_dense_fwd: null pointer");
    SBEV_REQUIRE(M >= 0 && K > 0 && N > 0 && ldx >= K, SBEV_ERR_INVALID, "sbev_dense_fwd: bad sizes");
    SBEV_REQUIRE(ldw >= N && (ldw & 3) == 0, SBEV_ERR_INVALID, "sbev_dense_fwd: ldw must be a multiple of 4 and >= N");
    SBEV_REQUIRE((ln_w == nullptr) == (ln_b == nullptr), SBEV_ERR_INVALID, "sbev_dense_fwd: ln_w and ln_b go together");
    SBEV_REQUIRE((reinterpret_cast<uintptr_t>(Wt) & 15) == 0, SBEV_ERR_INVALID, "sbev_dense_fwd: Wt not 16-byte aligned");
    if (M == 0) return SBEV_OK;
    const int Kp = (K + 3) & ~3;
    const size_t smem = sizeof(float) * (size_t)DENSE_ROWS * (Kp + ldw);
    SBEV_REQUIRE(smem <= 200 * 1024, SBEV_ERR_UNSUPPORTED, "sbev_dense_fwd: K + N too large (%d + %d)", K, N);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(dense_rows8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dense_rows8_kernel<<<(M + DENSE_ROWS - 1) / DENSE_ROWS, 256, smem, (cudaStream_t)stream>>>(
        x, ldx, Wt, ldw, bias, ln_w, ln_b, residual, M, K, N, flags, y);
    return check_launch("sbev_dense_fwd");
}

extern "C" int sbev_sample_points_fwd(const float* query_bbox, const float* offset, const float* scale_logits,
                                      const float* pc_range, int BQ, int GP, int L,
                                      float* points, float* scale_w, void* stream) {
    SBEV_REQUIRE(query_bbox && offset && scale_logits && pc_range && points && scale_w, SBEV_ERR_INVALID,
                 "sbev_sample_points_fwd: null pointer");
    SBEV_REQUIRE(BQ >= 0 && GP > 0 && L >= 1 && L <= SBEV_MAX_LEVELS, SBEV_ERR_INVALID, "sbev_sample_points_fwd: bad sizes");
    const long long total = (long long)BQ * GP;
    if (total == 0) return SBEV_OK;
    sample_points_kernel<<<(int)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        query_bbox, offset, scale_logits, pc_range[0], pc_range[1], pc_range[2], pc_range[3], pc_range[4], pc_range[5],
        BQ, GP, L, points, scale_w);
    return check_launch("sbev_sample_points_fwd");
}

extern "C" int sbev_split_bf16(const float* x, int64_t n, uint16_t* hi, uint16_t* lo, void* stream) {
    SBEV_REQUIRE(x && hi, SBEV_ERR_INVALID, "sbev_split_bf16: null pointer");
    SBEV_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(hi) & 7) == 0 &&
                 (reinterpret_cast<uintptr_t>(lo) & 7) == 0, SBEV_ERR_INVALID, "sbev_split_bf16: misaligned");
    if (n <= 0) return SBEV_OK;
    const long long threads = (n + 3) / 4;
    split_bf16_kernel<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x, n, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo));
    return check_launch("sbev_split_bf16");
}

extern "C" int sbev_reduce_ln_fwd(const float* partial, int nsplit, const float* bias, const float* residual,
                                  const float* ln_w, const float* ln_b, int M, int N, float* out, void* stream) {
    SBEV_REQUIRE(partial && out && nsplit >= 1, SBEV_ERR_INVALID, "sbev_reduce_ln_fwd: bad arguments");
    SBEV_REQUIRE(N > 0 && N <= 1024, SBEV_ERR_UNSUPPORTED, "sbev_reduce_ln_fwd: N must be in (0,1024]");
    SBEV_REQUIRE((ln_w == nullptr) == (ln_b == nullptr), SBEV_ERR_INVALID, "sbev_reduce_ln_fwd: ln_w and ln_b go together");
    if (M <= 0) return SBEV_OK;
    reduce_ln_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(partial, nsplit, bias, residual, ln_w, ln_b, M, N, out);
    return check_launch("sbev_reduce_ln_fwd");
}

extern "C" int sbev_refine_bbox_fwd(const float* proposal, const float* delta, const float* time_diff,
                                    int B, int Q, int T, int code_size, float* out, void* stream) {
    SBEV_REQUIRE(proposal && delta && time_diff && out, SBEV_ERR_INVALID, "sbev_refine_bbox_fwd: null pointer");
    SBEV_REQUIRE(B >= 0 && Q >= 0 && T >= 1 && code_size >= 10, SBEV_ERR_INVALID, "sbev_refine_bbox_fwd: bad sizes");
    const long long total = (long long)B * Q * code_size;
    if (total == 0) return SBEV_OK;
    refine_bbox_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(proposal, delta, time_diff, B, Q, T, code_size, out);
    return check_launch("sbev_refine_bbox_fwd");
}
