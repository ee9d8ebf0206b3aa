// Multi-scale multi-view bilinear gather ("msmv_sampling") for sm_100a, and the fused
// adaptive spatio-temporal sampling front-end.
//
// Behavioural reference (paths under /root/reference, nothing copied):
//   forward kernel semantics   models/csrc/msmv_sampling/msmv_sampling_forward.cu:27-267
//   backward kernel semantics  models/csrc/msmv_sampling/msmv_sampling_backward.cu:29-361
//   sampling_4d                models/sparsebev_sampling.py:27-130
//
// Design (B200): the gather is HBM/L2-bandwidth bound -- 4 corners x L levels x 256 B rows
// per sample point, 0.25 flop/byte.  One HALF-WARP owns one sample point: 16 lanes x float4
// = one 256 B channel row per LDG.128, so every load instruction of a warp moves two full
// rows, and each lane keeps 4*L independent 16 B loads in flight (16 for L=4) before the
// first use.  Coordinates / weights are fetched once per point and broadcast by shuffle
// instead of being re-read by all 64 channel threads as in the reference.  Work is ordered
// slice-major ((b,t,g) outermost) so the CTAs resident at any moment share one ~23 MB slice
// of the feature pyramid in L2.  Outputs are written as whole 256 B (fused layout) or
// 64*P*4 B (op layout, transposed through shared memory) contiguous segments.
#include "common.cuh"
#include <stdlib.h>

namespace sbev {

struct LevelSet {
    const float* ptr[SBEV_MAX_LEVELS];
    int H[SBEV_MAX_LEVELS];
    int W[SBEV_MAX_LEVELS];
    // fused-path addressing (floats): see sbev_sampling4d_fwd in the public header
    long long s_bt[SBEV_MAX_LEVELS], s_g[SBEV_MAX_LEVELS], s_v[SBEV_MAX_LEVELS], s_px[SBEV_MAX_LEVELS];
};

struct Tap {
    int y0, x0;
    float w1, w2, w3, w4;
    bool inside, ok1, ok2, ok3, ok4;
};

// Same fp32 operation order as the reference device code (forward.cu:33-46,123-126):
// single multiply loc*(size-1), floor, differences, products of the two 1-D weights.
__device__ __forceinline__ Tap make_tap(float u, float v, int H, int W) {
    Tap t;
    const float y = v * (float)(H - 1);
    const float x = u * (float)(W - 1);
    t.inside = (y > -1.f) && (x > -1.f) && (y < (float)H) && (x < (float)W);
    const float yf = floorf(y), xf = floorf(x);
    t.y0 = t.inside ? (int)yf : 0;
    t.x0 = t.inside ? (int)xf : 0;
    const float ly = y - yf, lx = x - xf;
    const float hy = 1.f - ly, hx = 1.f - lx;
    t.w1 = hy * hx; t.w2 = hy * lx; t.w3 = ly * hx; t.w4 = ly * lx;
    const bool yl = t.y0 >= 0, yh = t.y0 + 1 <= H - 1, xl = t.x0 >= 0, xh = t.x0 + 1 <= W - 1;
    t.ok1 = t.inside && yl && xl;
    t.ok2 = t.inside && yl && xh;
    t.ok3 = t.inside && yh && xl;
    t.ok4 = t.inside && yh && xh;
    return t;
}

__device__ __forceinline__ int view_from_coord(float z, int N) {
    return (int)roundf(z * (float)(N - 1));     // forward.cu:110 (round-half-away-from-zero)
}

// Gather all L levels for 4*VEC consecutive channels.  base[l] already points at
// (slice, view, channel lane); pxs[l] = floats between neighbouring pixels (pixel offsets fit 32 bits).
// LB = levels whose 4*LB*VEC loads are in flight together (LB == L: maximum memory-level parallelism;
// smaller LB: fewer live registers -> more resident warps).
// PF: while the loads of one level block are in flight, the lines the NEXT block will read are requested into L2
// (prefetch.global.L2, no destination registers): the kernel is DRAM-latency-bound at 2 CTAs/SM (three dependent round
// trips per point: geometry, levels 0-1, levels 2-3), and this overlaps the third with the second.
// SKIP: a level block none of the warp's points falls inside (dead points: no camera sees them) is skipped as a whole --
// warp-uniform branch, so the ~30 % of taps that load nothing also issue nothing.
template <int L, int LB, int VEC = 1, int VSTRIDE = 4, bool PF = false, bool SKIP = false>      // VSTRIDE: floats between the VEC float4 of one lane
__device__ __forceinline__ void gather_levels_v(const float* const (&base)[L], const int (&H)[L],
                                                const int (&W)[L], const int (&pxs)[L],
                                                float u, float v, const float (&wt)[L], bool live, float4 (&acc)[VEC]) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = zero;
#pragma unroll
    for (int l0 = 0; l0 < L; l0 += LB) {
        Tap tp[LB];
        float4 c1[LB][VEC], c2[LB][VEC], c3[LB][VEC], c4[LB][VEC];
        if (SKIP) {
            bool any_in = false;
#pragma unroll
            for (int i = 0; i < LB; ++i)
                if (l0 + i < L) { tp[i] = make_tap(u, v, H[l0 + i], W[l0 + i]); any_in |= live && tp[i].inside; }
            if (!__any_sync(0xffffffffu, any_in)) continue;
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            const int l = l0 + i;
            if (l < L) {
                if (!SKIP) tp[i] = make_tap(u, v, H[l], W[l]);
                const int row = W[l] * pxs[l];
                const float* p = base[l] + (tp[i].y0 * row + tp[i].x0 * pxs[l]);
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    c1[i][e] = (live && tp[i].ok1) ? ldg4(p + VSTRIDE * e) : zero;
                    c2[i][e] = (live && tp[i].ok2) ? ldg4(p + pxs[l] + VSTRIDE * e) : zero;
                    c3[i][e] = (live && tp[i].ok3) ? ldg4(p + row + VSTRIDE * e) : zero;
                    c4[i][e] = (live && tp[i].ok4) ? ldg4(p + row + pxs[l] + VSTRIDE * e) : zero;
                }
            }
        }
        if (PF && l0 + LB < L) {
#pragma unroll
            for (int i = 0; i < LB; ++i) {
                const int l = l0 + LB + i;
                if (l < L) {
                    const Tap t = make_tap(u, v, H[l], W[l]);
                    const int row = W[l] * pxs[l];
                    const float* p = base[l] + (t.y0 * row + t.x0 * pxs[l]);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        if (live && t.ok1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + VSTRIDE * e));
                        if (live && t.ok2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + pxs[l] + VSTRIDE * e));
                        if (live && t.ok3) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + row + VSTRIDE * e));
                        if (live && t.ok4) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + row + pxs[l] + VSTRIDE * e));
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            const int l = l0 + i;
            if (l < L && tp[i].inside) {
                const Tap& t = tp[i];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    acc[e].x += (t.w1 * c1[i][e].x + t.w2 * c2[i][e].x + t.w3 * c3[i][e].x + t.w4 * c4[i][e].x) * wt[l];
                    acc[e].y += (t.w1 * c1[i][e].y + t.w2 * c2[i][e].y + t.w3 * c3[i][e].y + t.w4 * c4[i][e].y) * wt[l];
                    acc[e].z += (t.w1 * c1[i][e].z + t.w2 * c2[i][e].z + t.w3 * c3[i][e].z + t.w4 * c4[i][e].z) * wt[l];
                    acc[e].w += (t.w1 * c1[i][e].w + t.w2 * c2[i][e].w + t.w3 * c3[i][e].w + t.w4 * c4[i][e].w) * wt[l];
                }
            }
        }
    }
}

template <int L, int LB>
__device__ __forceinline__ float4 gather_levels(const float* const (&base)[L], const int (&H)[L],
                                                const int (&W)[L], const int (&pxs)[L],
                                                float u, float v, const float (&wt)[L], bool live) {
    float4 acc[1];
    gather_levels_v<L, LB, 1, 4>(base, H, W, pxs, u, v, wt, live, acc);
    return acc[0];
}

// ------------------------------------------------------------------------------------------
// Op-boundary forward, C == 64.  One warp per (b', q): lanes < P prefetch that query's
// coordinates and weights, the two half-warps then walk the P points two at a time, results go
// to a per-warp shared tile [P][64] and leave as one contiguous 64*P float segment in the
// reference's [B',Q,C,P] order.
template <int L>
__global__ void __launch_bounds__(256)
msmv_fwd_c64_kernel(LevelSet lv, const float* __restrict__ loc, const float* __restrict__ wgt,
                    int Bp, int N, int Q, int P, float* __restrict__ out) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = lane >> 4, j = lane & 15;
    float* tile = smem + warp * (64 * P);
    const int items = Bp * Q;
    for (int item = blockIdx.x * 8 + warp; item < items; item += gridDim.x * 8) {
        const int b = item / Q;
        float lu = 0.f, lvv = 0.f, lz = 0.f, lw[L];
#pragma unroll
        for (int l = 0; l < L; ++l) lw[l] = 0.f;
        if (lane < P) {
            const float* lp = loc + ((long long)item * P + lane) * 3;
            lu = __ldg(lp); lvv = __ldg(lp + 1); lz = __ldg(lp + 2);
            const float* wp = wgt + ((long long)item * P + lane) * L;
#pragma unroll
            for (int l = 0; l < L; ++l) lw[l] = __ldg(wp + l);
        }
        for (int p0 = 0; p0 < P; p0 += 2) {
            const int p = p0 + half;
            const bool live = p < P;
            const int src = live ? p : P - 1;
            const float u = __shfl_sync(0xffffffffu, lu, src);
            const float v = __shfl_sync(0xffffffffu, lvv, src);
            const float z = __shfl_sync(0xffffffffu, lz, src);
            float wt[L];
#pragma unroll
            for (int l = 0; l < L; ++l) wt[l] = __shfl_sync(0xffffffffu, lw[l], src);
            const int view = view_from_coord(z, N);
            const bool view_ok = (view >= 0) && (view < N);
            const float* base[L]; int H[L], W[L]; int pxs[L];
#pragma unroll
            for (int l = 0; l < L; ++l) {
                H[l] = lv.H[l]; W[l] = lv.W[l]; pxs[l] = 64;
                base[l] = lv.ptr[l] + ((long long)b * N + (view_ok ? view : 0)) * H[l] * W[l] * 64 + 4 * j;
            }
            const float4 acc = gather_levels<L, L>(base, H, W, pxs, u, v, wt, live && view_ok);
            if (live) *reinterpret_cast<float4*>(tile + p * 64 + 4 * j) = acc;
        }
        __syncwarp();
        float* dst = out + (long long)item * 64 * P;
        for (int i = lane * 4; i < 64 * P; i += 128) {
            float4 o;
            o.x = tile[((i + 0) % P) * 64 + (i + 0) / P];
            o.y = tile[((i + 1) % P) * 64 + (i + 1) / P];
            o.z = tile[((i + 2) % P) * 64 + (i + 2) / P];
            o.w = tile[((i + 3) % P) * 64 + (i + 3) / P];
            *reinterpret_cast<float4*>(dst + i) = o;
        }
        __syncwarp();
    }
}

// Generic-C fallback (any C, reference thread mapping but coordinates hoisted): thread per (b',q,c).
__global__ void __launch_bounds__(256)
msmv_fwd_generic_kernel(LevelSet lv, int L, const float* __restrict__ loc, const float* __restrict__ wgt,
                        int Bp, int N, int C, int Q, int P, float* __restrict__ out) {
    const long long total = (long long)Bp * Q * C;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        const long long item = idx / C;
        const int b = (int)(item / Q);
        for (int p = 0; p < P; ++p) {
            const float* lp = loc + (item * P + p) * 3;
            const float u = __ldg(lp), v = __ldg(lp + 1);
            const int view = view_from_coord(__ldg(lp + 2), N);
            float acc = 0.f;
            if (view >= 0 && view < N) {
                for (int l = 0; l < L; ++l) {
                    const int H = lv.H[l], W = lv.W[l];
                    const Tap t = make_tap(u, v, H, W);
                    if (!t.inside) continue;
                    const float* p0 = lv.ptr[l] + (((long long)b * N + view) * H * W + (long long)t.y0 * W + t.x0) * C + c;
                    const float v1 = t.ok1 ? __ldg(p0) : 0.f;
                    const float v2 = t.ok2 ? __ldg(p0 + C) : 0.f;
                    const float v3 = t.ok3 ? __ldg(p0 + (long long)W * C) : 0.f;
                    const float v4 = t.ok4 ? __ldg(p0 + (long long)W * C + C) : 0.f;
                    acc += (t.w1 * v1 + t.w2 * v2 + t.w3 * v3 + t.w4 * v4) * __ldg(wgt + (item * P + p) * L + l);
                }
            }
            out[idx * P + p] = acc;
        }
    }
}

__global__ void msmv_indices_kernel(LevelSet lv, int L, const float* __restrict__ loc, long long npts, int N,
                                    int32_t* view, int32_t* y0, int32_t* x0, int32_t* inside) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npts;
         i += (long long)gridDim.x * blockDim.x) {
        const float u = loc[i * 3], v = loc[i * 3 + 1];
        view[i] = view_from_coord(loc[i * 3 + 2], N);
        for (int l = 0; l < L; ++l) {
            const float y = v * (float)(lv.H[l] - 1), x = u * (float)(lv.W[l] - 1);
            y0[i * L + l] = (int)floorf(y);
            x0[i * L + l] = (int)floorf(x);
            inside[i * L + l] = (y > -1.f) && (x > -1.f) && (y < (float)lv.H[l]) && (x < (float)lv.W[l]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Backward, C == 64: half-warp per (b',q,p).  grad wrt weights and locations are owned by exactly
// one half-warp -> shuffle reduction + plain store (deterministic, no atomics, unlike the
// reference); only the scatter into grad_feats uses (128-bit vector) atomics.
__device__ __forceinline__ void red_add4(float* addr, float4 v) {
#if __CUDA_ARCH__ >= 900
    atomicAdd(reinterpret_cast<float4*>(addr), v);
#else
    atomicAdd(addr, v.x); atomicAdd(addr + 1, v.y); atomicAdd(addr + 2, v.z); atomicAdd(addr + 3, v.w);
#endif
}

struct GradLevelSet { float* ptr[SBEV_MAX_LEVELS]; };

// SCATTER = false: only grad_loc / grad_w (the deterministic path accumulates grad_feats per pixel, below).
template <int L, bool SCATTER = true>
__global__ void __launch_bounds__(256)
msmv_bwd_c64_kernel(LevelSet lv, GradLevelSet glv, const float* __restrict__ grad_out,
                    const float* __restrict__ loc, const float* __restrict__ wgt,
                    int Bp, int N, int Q, int P,
                    float* __restrict__ grad_loc, float* __restrict__ grad_w) {
    const int lane = threadIdx.x & 31, j = lane & 15;
    const long long npts = (long long)Bp * Q * P;
    const long long hw_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const long long hw_stride = ((long long)gridDim.x * blockDim.x) >> 4;
    const long long iters = (npts + hw_stride - 1) / hw_stride;        // uniform trip count (shuffles inside)
    for (long long it = 0; it < iters; ++it) {
        const long long pi_raw = hw_id + it * hw_stride;
        const bool live = pi_raw < npts;
        const long long pi = live ? pi_raw : npts - 1;
        const int p = (int)(pi % P);
        const long long item = pi / P;
        const int b = (int)(item / Q);
        const float u = __ldg(loc + pi * 3), v = __ldg(loc + pi * 3 + 1);
        const int view = view_from_coord(__ldg(loc + pi * 3 + 2), N);
        const bool view_ok = view >= 0 && view < N;
        float4 g;
        {
            const float* gp = grad_out + (item * 64 + 4 * j) * P + p;
            g.x = __ldg(gp); g.y = __ldg(gp + P); g.z = __ldg(gp + 2 * P); g.w = __ldg(gp + 3 * P);
        }
        float gx_acc = 0.f, gy_acc = 0.f;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const int H = lv.H[l], W = lv.W[l];
            const Tap t = make_tap(u, v, H, W);
            const float aw = __ldg(wgt + pi * L + l);
            float gw_part = 0.f, gx_part = 0.f, gy_part = 0.f;
            if (t.inside && view_ok) {
                const long long off = (((long long)b * N + view) * H * W + (long long)t.y0 * W + t.x0) * 64 + 4 * j;
                const float* p0 = lv.ptr[l] + off;
                float* q0 = glv.ptr[l] + off;
                const long long row = (long long)W * 64;
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 v1 = t.ok1 ? ldg4(p0) : zero;
                const float4 v2 = t.ok2 ? ldg4(p0 + 64) : zero;
                const float4 v3 = t.ok3 ? ldg4(p0 + row) : zero;
                const float4 v4 = t.ok4 ? ldg4(p0 + row + 64) : zero;
                const float hy = t.w1 + t.w2, ly = t.w3 + t.w4;      // hy*(hx+lx) = hy up to 1 ulp; recompute exactly below
                (void)hy; (void)ly;
                const float y = v * (float)(H - 1), x = u * (float)(W - 1);
                const float fly = y - floorf(y), flx = x - floorf(x);
                const float fhy = 1.f - fly, fhx = 1.f - flx;
                const float4 gv = make_float4(g.x * aw, g.y * aw, g.z * aw, g.w * aw);   // top_grad * attn_weight
                if (SCATTER && live) {
                    if (t.ok1) red_add4(q0, make_float4(t.w1 * gv.x, t.w1 * gv.y, t.w1 * gv.z, t.w1 * gv.w));
                    if (t.ok2) red_add4(q0 + 64, make_float4(t.w2 * gv.x, t.w2 * gv.y, t.w2 * gv.z, t.w2 * gv.w));
                    if (t.ok3) red_add4(q0 + row, make_float4(t.w3 * gv.x, t.w3 * gv.y, t.w3 * gv.z, t.w3 * gv.w));
                    if (t.ok4) red_add4(q0 + row + 64, make_float4(t.w4 * gv.x, t.w4 * gv.y, t.w4 * gv.z, t.w4 * gv.w));
                }
#define SBEV_BWD_CH(comp)                                                                              \
                {                                                                                      \
                    const float a1 = v1.comp, a2 = v2.comp, a3 = v3.comp, a4 = v4.comp;                \
                    const float val = t.w1 * a1 + t.w2 * a2 + t.w3 * a3 + t.w4 * a4;                   \
                    const float gyw = -fhx * a1 - flx * a2 + fhx * a3 + flx * a4;                      \
                    const float gxw = -fhy * a1 + fhy * a2 - fly * a3 + fly * a4;                      \
                    gw_part += g.comp * val;                                                           \
                    gx_part += gxw * gv.comp;                                                          \
                    gy_part += gyw * gv.comp;                                                          \
                }
                SBEV_BWD_CH(x) SBEV_BWD_CH(y) SBEV_BWD_CH(z) SBEV_BWD_CH(w)
#undef SBEV_BWD_CH
            }
            gw_part = half_warp_sum(gw_part);
            gx_acc += (float)(W - 1) * half_warp_sum(gx_part);
            gy_acc += (float)(H - 1) * half_warp_sum(gy_part);
            if (live && j == 0) grad_w[pi * L + l] = gw_part;
        }
        if (live && j == 0) {
            grad_loc[pi * 3 + 0] = gx_acc;
            grad_loc[pi * 3 + 1] = gy_acc;
            grad_loc[pi * 3 + 2] = 0.f;       // reference never writes the view-coordinate gradient
        }
    }
}

// Generic-C backward: thread per (b',q,c,p), scalar atomics everywhere (reference mapping).
__global__ void __launch_bounds__(256)
msmv_bwd_generic_kernel(LevelSet lv, GradLevelSet glv, int L, const float* __restrict__ grad_out,
                        const float* __restrict__ loc, const float* __restrict__ wgt,
                        int Bp, int N, int C, int Q, int P, float* grad_loc, float* grad_w) {
    const long long total = (long long)Bp * Q * C * P;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(idx % P);
        const int c = (int)((idx / P) % C);
        const long long item = idx / ((long long)P * C);
        const int b = (int)(item / Q);
        const long long pi = item * P + p;
        const float g = grad_out[idx];
        const float u = loc[pi * 3], v = loc[pi * 3 + 1];
        const int view = view_from_coord(loc[pi * 3 + 2], N);
        if (view < 0 || view >= N) continue;
        for (int l = 0; l < L; ++l) {
            const int H = lv.H[l], W = lv.W[l];
            const Tap t = make_tap(u, v, H, W);
            if (!t.inside) continue;
            const float aw = wgt[pi * L + l], gv = g * aw;
            const long long off = (((long long)b * N + view) * H * W + (long long)t.y0 * W + t.x0) * C + c;
            const float* p0 = lv.ptr[l] + off;
            float* q0 = glv.ptr[l] + off;
            const long long row = (long long)W * C;
            const float y = v * (float)(H - 1), x = u * (float)(W - 1);
            const float ly = y - floorf(y), lx = x - floorf(x), hy = 1.f - ly, hx = 1.f - lx;
            float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
            if (t.ok1) { v1 = p0[0]; atomicAdd(q0, t.w1 * gv); }
            if (t.ok2) { v2 = p0[C]; atomicAdd(q0 + C, t.w2 * gv); }
            if (t.ok3) { v3 = p0[row]; atomicAdd(q0 + row, t.w3 * gv); }
            if (t.ok4) { v4 = p0[row + C]; atomicAdd(q0 + row + C, t.w4 * gv); }
            const float val = t.w1 * v1 + t.w2 * v2 + t.w3 * v3 + t.w4 * v4;
            atomicAdd(grad_w + pi * L + l, g * val);
            atomicAdd(grad_loc + pi * 3, (float)(W - 1) * (-hy * v1 + hy * v2 - ly * v3 + ly * v4) * gv);
            atomicAdd(grad_loc + pi * 3 + 1, (float)(H - 1) * (-hx * v1 - lx * v2 + hx * v3 + lx * v4) * gv);
        }
    }
}


// ------------------------------------------------------------------------------------------
// Deterministic grad_feats ("col2im without floating-point atomics").  The scatter of the reference
// (msmv_sampling_backward.cu:29-224: one atomicAdd per corner and channel) is inverted into a per-pixel gather:
//   1. bin:    every (point, level, corner) contribution counts into its pixel            (integer atomics: exact)
//   2. scan:   exclusive prefix sum of the counts -> segment offsets
//   3. fill:   contribution ids (point*4 + corner) dropped into their pixel's segment      (order within a segment arbitrary)
//   4. reduce: one half-warp per pixel walks its segment in ASCENDING id order (selection by repeated half-warp min)
//              and sums  tap * scale_weight * grad_out  in registers; every pixel row -- touched or not -- is written
//              exactly once by plain 16-byte stores, so grad_feats needs no zero fill either.
// The summation order is a function of the inputs only: two runs give bit-identical gradients.
struct PixelSpace { long long base[SBEV_MAX_LEVELS + 1]; };      // first pixel id of every level; base[L] = pixel count

template <bool FILL>
__global__ void __launch_bounds__(256)
msmv_bwd_bin_kernel(LevelSet lv, PixelSpace ps, int L, const float* __restrict__ loc, long long npts, int N, int QP,
                    int* __restrict__ cnt, const int* __restrict__ off, int* __restrict__ ids) {
    const long long total = npts * L;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pi = i / L;
        const int l = (int)(i - pi * L);
        const int view = view_from_coord(__ldg(loc + pi * 3 + 2), N);
        if (view < 0 || view >= N) continue;
        const int H = lv.H[l], W = lv.W[l];
        const Tap t = make_tap(__ldg(loc + pi * 3), __ldg(loc + pi * 3 + 1), H, W);
        if (!t.inside) continue;
        const long long b = pi / QP;
        const long long pix = ps.base[l] + (b * N + view) * H * W + (long long)t.y0 * W + t.x0;
        const bool ok[4] = {t.ok1, t.ok2, t.ok3, t.ok4};
        const long long px[4] = {pix, pix + 1, pix + W, pix + W + 1};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (!ok[c]) continue;
            if (!FILL) atomicAdd(cnt + px[c], 1);
            else ids[off[px[c]] + atomicSub(cnt + px[c], 1) - 1] = (int)(pi * 4 + c);
        }
    }
}

constexpr int SCAN_ITEMS = 2048;                                 // counts per block (256 threads x 8)

__device__ __forceinline__ int block_exclusive_scan_256(int v, int* total, int* sh /* [9] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (threadIdx.x == 0) { int run = 0; for (int w = 0; w < 8; ++w) { const int t = sh[w]; sh[w] = run; run += t; } sh[8] = run; }
    __syncthreads();
    const int excl = sh[warp] + inc - v;
    *total = sh[8];
    __syncthreads();
    return excl;
}

__global__ void __launch_bounds__(256) scan_block_sums_kernel(const int* __restrict__ cnt, long long n, int* __restrict__ bsum) {
    __shared__ int sh[9];
    const long long base = (long long)blockIdx.x * SCAN_ITEMS + threadIdx.x * 8;
    int s = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) if (base + e < n) s += cnt[base + e];
    int total;
    block_exclusive_scan_256(s, &total, sh);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) scan_bsums_kernel(int* __restrict__ bsum, int nb) {      // ONE block: in-place exclusive scan
    __shared__ int sh[9];
    int carry = 0;
    for (int b0 = 0; b0 < nb; b0 += 256) {
        const int i = b0 + threadIdx.x;
        const int v = i < nb ? bsum[i] : 0;
        int total;
        const int ex = block_exclusive_scan_256(v, &total, sh);
        if (i < nb) bsum[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(256) scan_apply_kernel(const int* __restrict__ cnt, long long n, const int* __restrict__ bsum,
                                                         int* __restrict__ off) {
    __shared__ int sh[9];
    const long long base = (long long)blockIdx.x * SCAN_ITEMS + threadIdx.x * 8;
    int v[8], s = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) { v[e] = base + e < n ? cnt[base + e] : 0; s += v[e]; }
    int total;
    int run = bsum[blockIdx.x] + block_exclusive_scan_256(s, &total, sh);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        if (base + e < n) off[base + e] = run;
        run += v[e];
        if (base + e == n - 1) off[n] = run;                   // grand total closes the last segment
    }
}

template <int L>
__global__ void __launch_bounds__(256)
msmv_bwd_reduce_kernel(LevelSet lv, GradLevelSet glv, PixelSpace ps, const int* __restrict__ off, const int* __restrict__ ids,
                       const float* __restrict__ grad_out, const float* __restrict__ loc, const float* __restrict__ wgt, int P) {
    const int lane = threadIdx.x & 31, j = lane & 15;
    const unsigned hmask = 0xffffu << (lane & 16);              // the two half-warps of a warp walk different segments
    const long long total = ps.base[L];
    const long long hw_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const long long hw_stride = ((long long)gridDim.x * blockDim.x) >> 4;
    for (long long pix = hw_id; pix < total; pix += hw_stride) {
        int l = 0;
#pragma unroll
        for (int k = 1; k < L; ++k) if (pix >= ps.base[k]) l = k;
        const int H = lv.H[l], W = lv.W[l];
        const long long local = pix - ps.base[l];
        const int beg = __ldg(off + pix), n = __ldg(off + pix + 1) - beg;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int mine = (n <= 16 && j < n) ? __ldg(ids + beg + j) : 0x7fffffff;     // short segment: one id per lane
        int last = -1;
        for (int k = 0; k < n; ++k) {
            int m = 0x7fffffff;
            if (n <= 16) { if (mine > last) m = mine; }
            else for (int i = j; i < n; i += 16) { const int v = __ldg(ids + beg + i); if (v > last && v < m) m = v; }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(hmask, m, o));
            last = m;                                                                 // ids are unique: strictly ascending walk
            const long long pi = m >> 2;
            const int corner = m & 3;
            const Tap t = make_tap(__ldg(loc + pi * 3), __ldg(loc + pi * 3 + 1), H, W);
            const float tw = corner == 0 ? t.w1 : corner == 1 ? t.w2 : corner == 2 ? t.w3 : t.w4;
            const float aw = __ldg(wgt + pi * L + l);
            const long long item = pi / P;
            const int p = (int)(pi - item * P);
            const float* gp = grad_out + (item * 64 + 4 * j) * P + p;
            const float4 gv = make_float4(__ldg(gp) * aw, __ldg(gp + P) * aw, __ldg(gp + 2 * P) * aw, __ldg(gp + 3 * P) * aw);
            acc.x += tw * gv.x; acc.y += tw * gv.y; acc.z += tw * gv.z; acc.w += tw * gv.w;
        }
        *reinterpret_cast<float4*>(glv.ptr[l] + local * 64 + 4 * j) = acc;
    }
}

// ------------------------------------------------------------------------------------------
// Fused adaptive spatio-temporal sampling (see sbev_sampling4d_fwd in the header).
// Half-warp per sample (b,t,g,q,p), slice-major order.  Lanes 0..N-1 of the half-warp each
// project the point into one camera; a ballot picks the first valid view.
struct FusedParams {
    const float* points;     // [B,Q,G*P,3]
    const float* velocity;   // (b,q) -> velocity + (b*Q+q)*ld_vel: [B,Q,2] packed (ld_vel 2) or query_bbox + 8 (ld_vel 10)
    const float* time_diff;  // [B,T]
    const float* lidar2img;  // [B,T*N,16]
    const float* scale_w;    // [B,Q,G,P,L]
    float* out[SBEV_MAX_PEERS];   // n_out buffers [B,Q,G,To*P,64] that all receive the window's rows (own + peer GPUs)
    float* loc_out;          // optional [B*Tl*G,Q,P,3]
    int B, T, G, N, Q, P;
    float image_h, image_w, eps;
    int t0, Tl;              // frame window [t0, t0+Tl) held by feats (t0 = 0, Tl = T: all frames)
    int ld_vel;
    int o0, To, n_out;       // the out buffers cover frames [o0, o0+To): the window itself, or all T frames (scatter form)
    int q_per_rank;          // > 0 (owner form, B == 1): query q's rows go ONLY to out[q / q_per_rank], a [q_per_rank,G,T*P,64] buffer
};

// LPP = lanes per sample point: 16 (4 channels per lane) or 8 (8 channels per lane; halves the per-point geometry that
// every lane of a point computes redundantly -- the kernel is issue-bound, not bandwidth-bound, on the realistic rig).
template <int L, int LB, int MINB, int LPP, bool PF = false, bool SKIP = false>
__global__ void __launch_bounds__(256, MINB)
sampling4d_c64_kernel(LevelSet lv, FusedParams prm) {
    constexpr int VEC = 16 / LPP;                     // float4 per lane
    constexpr int PPB = 256 / LPP;                    // points per block
    const int lane = threadIdx.x & 31, half = lane / LPP, j = lane % LPP;
    const int T = prm.T, G = prm.G, N = prm.N, Q = prm.Q, P = prm.P;
    // slice (b,t,g) = blockIdx.y (uniform); sample (q,p) inside the slice from blockIdx.x: all 32-bit, no 64-bit div/mod
    pdl_wait();
    pdl_trigger();
    const int s = blockIdx.y, Tl = prm.Tl;
    const int g = s % G, btl = s / G, tl = btl % Tl, b = btl / Tl;   // btl indexes the LOCAL feature slices
    const int t = prm.t0 + tl, bt = b * T + t;                       // bt indexes time_diff / lidar2img (all T frames)
    const int idx_raw = blockIdx.x * PPB + (threadIdx.x / LPP);
    const bool live = idx_raw < Q * P;
    const int idx = live ? idx_raw : Q * P - 1;
    const int q = (P == 4) ? (idx >> 2) : (idx / P);
    const int p = idx - q * P;
    const long long bq = (long long)b * Q + q;

    // motion warp (sparsebev_transformer.py:286-295): xy -= vel * time_diff[t]; no FMA contraction
    const float* pp = prm.points + (bq * G * P + g * P + p) * 3;
    const float td = __ldg(prm.time_diff + bt);
    const float px = __fsub_rn(__ldg(pp), __fmul_rn(__ldg(prm.velocity + bq * prm.ld_vel), td));
    const float py = __fsub_rn(__ldg(pp + 1), __fmul_rn(__ldg(prm.velocity + bq * prm.ld_vel + 1), td));
    const float pz = __ldg(pp + 2);

    // scale weights: the reference pairs loc slice (b,t,g) with weight slice (b,g',t'),
    // (g',t') = divmod(t*G+g, T)  (sparsebev_sampling.py:112-119); weights do not depend on t'.
    const int gw = (t * G + g) / T;
    const float* wp = prm.scale_w + ((bq * G + gw) * P + p) * L;
    float wt[L];
#pragma unroll
    for (int l = 0; l < L; ++l) wt[l] = __ldg(wp + l);

    // projection to view j (sparsebev_sampling.py:50-79), fixed order ((x*m0 + y*m1) + z*m2) + m3
    float un = 0.f, vn = 0.f;
    bool valid = false;
    if (j < N) {
        const float4* m = reinterpret_cast<const float4*>(prm.lidar2img + ((long long)bt * N + j) * 16);
        const float4 r0 = __ldg(m), r1 = __ldg(m + 1), r2 = __ldg(m + 2);
        const float cx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, r0.x), __fmul_rn(py, r0.y)), __fmul_rn(pz, r0.z)), r0.w);
        const float cy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, r1.x), __fmul_rn(py, r1.y)), __fmul_rn(pz, r1.z)), r1.w);
        const float dz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, r2.x), __fmul_rn(py, r2.y)), __fmul_rn(pz, r2.z)), r2.w);
        const float safe = fmaxf(dz, prm.eps);
        un = __fdiv_rn(__fdiv_rn(cx, safe), prm.image_w);
        vn = __fdiv_rn(__fdiv_rn(cy, safe), prm.image_h);
        valid = (dz > prm.eps) && (vn > 0.f) && (vn < 1.f) && (un > 0.f) && (un < 1.f);
    }
    const unsigned ball = (__ballot_sync(0xffffffffu, valid) >> (LPP * half)) & ((1u << LPP) - 1u);
    const int view = ball ? (__ffs(ball) - 1) : 0;          // argmax of 0/1 flags: first valid, else 0
    const float u = __shfl_sync(0xffffffffu, un, LPP * half + view);
    const float v = __shfl_sync(0xffffffffu, vn, LPP * half + view);

    const float* base[L]; int H[L], W[L]; int pxs[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        H[l] = lv.H[l]; W[l] = lv.W[l]; pxs[l] = (int)lv.s_px[l];
        base[l] = lv.ptr[l] + ((long long)btl * lv.s_bt[l] + (long long)g * lv.s_g[l] + (long long)view * lv.s_v[l] + 4 * j);
    }
    float4 acc[VEC];
    gather_levels_v<L, LB, VEC, 4 * LPP, PF, SKIP>(base, H, W, pxs, u, v, wt, live, acc);     // lane j: channels 4j..4j+3 (+ 4*LPP per extra vector)
    if (live) {
        if (prm.q_per_rank > 0) {                      // owner form: the row goes to the one rank that mixes query q
            const int ow = q / prm.q_per_rank;
            float* dst = prm.out[ow] + ((((long long)(q - ow * prm.q_per_rank) * G + g) * (T * P) + (t * P + p)) * 64 + 4 * j);
#pragma unroll
            for (int e = 0; e < VEC; ++e) *reinterpret_cast<float4*>(dst + 4 * LPP * e) = acc[e];
        } else {
        const long long row = ((bq * G + g) * (prm.To * P) + ((t - prm.o0) * P + p)) * 64 + 4 * j;
        for (int w = 0; w < prm.n_out; ++w) {          // n_out > 1: the same 256 B row also goes to the peers over NVLink
            float* dst = prm.out[w] + row;
#pragma unroll
            for (int e = 0; e < VEC; ++e) *reinterpret_cast<float4*>(dst + 4 * LPP * e) = acc[e];
        }
        }
        if (prm.loc_out != nullptr && j == 0) {
            float* lo = prm.loc_out + (((long long)s * Q + q) * P + p) * 3;
            lo[0] = u; lo[1] = v; lo[2] = __fdiv_rn((float)view, (float)(N - 1));
        }
    }
}

// ------------------------------------------------------------------------------------------ host
static int fill_levels(LevelSet& lv, const float* const* feats, const int* hw, int L) {
    SBEV_REQUIRE(L >= 1 && L <= SBEV_MAX_LEVELS, SBEV_ERR_UNSUPPORTED, "num levels %d outside [1,%d]", L, SBEV_MAX_LEVELS);
    SBEV_REQUIRE(hw != nullptr, SBEV_ERR_INVALID, "hw is null");
    for (int l = 0; l < SBEV_MAX_LEVELS; ++l) {
        lv.ptr[l] = nullptr; lv.H[l] = 1; lv.W[l] = 1;
        lv.s_bt[l] = lv.s_g[l] = lv.s_v[l] = lv.s_px[l] = 0;
    }
    for (int l = 0; l < L; ++l) {
        if (feats) {
            SBEV_REQUIRE(feats[l] != nullptr, SBEV_ERR_INVALID, "feats[%d] is null", l);
            SBEV_REQUIRE((reinterpret_cast<uintptr_t>(feats[l]) & 15) == 0, SBEV_ERR_INVALID, "feats[%d] not 16-byte aligned", l);
            lv.ptr[l] = feats[l];
        }
        lv.H[l] = hw[2 * l]; lv.W[l] = hw[2 * l + 1];
        SBEV_REQUIRE(lv.H[l] > 0 && lv.W[l] > 0, SBEV_ERR_INVALID, "level %d has non-positive size", l);
    }
    return SBEV_OK;
}

static int grid_for(long long work_items, int per_block, int max_blocks = 148 * 32) {
    long long g = (work_items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int)g;
}

}  // namespace sbev

using namespace sbev;

extern "C" int sbev_msmv_fwd(const float* const* feats, const int* hw, int L, const float* loc, const float* w,
                             int Bp, int N, int C, int Q, int P, float* out, void* stream) {
    SBEV_REQUIRE(Bp >= 0 && Q >= 0 && N > 0 && C > 0 && P > 0, SBEV_ERR_INVALID, "sbev_msmv_fwd: bad sizes");
    SBEV_REQUIRE(P <= SBEV_MAX_POINTS, SBEV_ERR_INVALID, "num_point exceed limits (%d > %d)", P, SBEV_MAX_POINTS);
    SBEV_REQUIRE(L >= 1 && L <= SBEV_MAX_LEVELS, SBEV_ERR_UNSUPPORTED, "num levels %d outside [1,%d]", L, SBEV_MAX_LEVELS);
    if ((long long)Bp * Q == 0) return SBEV_OK;          // empty problem: nothing to read or write (buffers may be null)
    SBEV_REQUIRE(feats && loc && w && out, SBEV_ERR_INVALID, "sbev_msmv_fwd: null pointer");
    SBEV_REQUIRE((long long)Bp * Q < (1ll << 31), SBEV_ERR_UNSUPPORTED, "sbev_msmv_fwd: Bp*Q too large");
    LevelSet lv;
    int rc = fill_levels(lv, feats, hw, L);
    if (rc) return rc;
    if ((long long)Bp * Q == 0) return SBEV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 64) {
        const size_t smem = (size_t)8 * 64 * P * sizeof(float);
        const int grid = grid_for((long long)Bp * Q, 8);
#define SBEV_LAUNCH_FWD(LL)                                                                                  \
        case LL: {                                                                                           \
            if (smem > 48 * 1024)                                                                            \
                cudaFuncSetAttribute(msmv_fwd_c64_kernel<LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            msmv_fwd_c64_kernel<LL><<<grid, 256, smem, st>>>(lv, loc, w, Bp, N, Q, P, out);                  \
        } break;
        switch (L) { SBEV_LAUNCH_FWD(1) SBEV_LAUNCH_FWD(2) SBEV_LAUNCH_FWD(3) SBEV_LAUNCH_FWD(4) SBEV_LAUNCH_FWD(5) }
#undef SBEV_LAUNCH_FWD
    } else {
        msmv_fwd_generic_kernel<<<grid_for((long long)Bp * Q * C, 256), 256, 0, st>>>(lv, L, loc, w, Bp, N, C, Q, P, out);
    }
    return check_launch("sbev_msmv_fwd");
}

extern "C" int sbev_msmv_indices(const int* hw, int L, const float* loc, int Bp, int N, int Q, int P,
                                 int32_t* view, int32_t* y0, int32_t* x0, int32_t* inside, void* stream) {
    SBEV_REQUIRE(loc && view && y0 && x0 && inside, SBEV_ERR_INVALID, "sbev_msmv_indices: null pointer");
    LevelSet lv;
    int rc = fill_levels(lv, nullptr, hw, L);
    if (rc) return rc;
    const long long npts = (long long)Bp * Q * P;
    if (npts == 0) return SBEV_OK;
    msmv_indices_kernel<<<grid_for(npts, 256), 256, 0, (cudaStream_t)stream>>>(lv, L, loc, npts, N, view, y0, x0, inside);
    return check_launch("sbev_msmv_indices");
}

extern "C" int sbev_msmv_bwd(const float* grad_out, const float* const* feats, const int* hw, int L,
                             const float* loc, const float* w, int Bp, int N, int C, int Q, int P,
                             float* const* grad_feats, float* grad_loc, float* grad_w, void* stream) {
    SBEV_REQUIRE(grad_out && feats && loc && w && grad_feats && grad_loc && grad_w, SBEV_ERR_INVALID, "sbev_msmv_bwd: null pointer");
    SBEV_REQUIRE(Bp >= 0 && Q >= 0 && N > 0 && C > 0 && P > 0, SBEV_ERR_INVALID, "sbev_msmv_bwd: bad sizes");
    SBEV_REQUIRE(P <= SBEV_MAX_POINTS, SBEV_ERR_INVALID, "num_point exceed limits (%d > %d)", P, SBEV_MAX_POINTS);
    LevelSet lv;
    int rc = fill_levels(lv, feats, hw, L);
    if (rc) return rc;
    GradLevelSet glv;
    cudaStream_t st = (cudaStream_t)stream;
    for (int l = 0; l < SBEV_MAX_LEVELS; ++l) glv.ptr[l] = nullptr;
    for (int l = 0; l < L; ++l) {
        SBEV_REQUIRE(grad_feats[l] != nullptr, SBEV_ERR_INVALID, "grad_feats[%d] is null", l);
        glv.ptr[l] = grad_feats[l];
        cudaMemsetAsync(grad_feats[l], 0, sizeof(float) * (size_t)Bp * N * lv.H[l] * lv.W[l] * C, st);
    }
    const long long npts = (long long)Bp * Q * P;
    if (npts == 0) return check_launch("sbev_msmv_bwd(memset)");
    if (C == 64) {
        const int grid = grid_for(npts, 16);
#define SBEV_LAUNCH_BWD(LL) case LL: msmv_bwd_c64_kernel<LL><<<grid, 256, 0, st>>>(lv, glv, grad_out, loc, w, Bp, N, Q, P, grad_loc, grad_w); break;
        switch (L) { SBEV_LAUNCH_BWD(1) SBEV_LAUNCH_BWD(2) SBEV_LAUNCH_BWD(3) SBEV_LAUNCH_BWD(4) SBEV_LAUNCH_BWD(5) }
#undef SBEV_LAUNCH_BWD
    } else {
        cudaMemsetAsync(grad_loc, 0, sizeof(float) * (size_t)npts * 3, st);
        cudaMemsetAsync(grad_w, 0, sizeof(float) * (size_t)npts * L, st);
        msmv_bwd_generic_kernel<<<grid_for(npts * C, 256), 256, 0, st>>>(lv, glv, L, grad_out, loc, w, Bp, N, C, Q, P, grad_loc, grad_w);
    }
    return check_launch("sbev_msmv_bwd");
}

// Workspace layout of the deterministic backward (ints): cnt[npix] | off[npix + 1] | bsum[blocks] | ids[npts * L * 4]
static int det_layout(const int* hw, int L, int Bp, int N, int Q, int P, PixelSpace* ps, long long* n_blocks, long long* n_ids) {
    SBEV_REQUIRE(hw && L >= 1 && L <= SBEV_MAX_LEVELS, SBEV_ERR_INVALID, "sbev_msmv_bwd_det: bad level description");
    SBEV_REQUIRE(Bp >= 0 && Q >= 0 && N > 0 && P > 0, SBEV_ERR_INVALID, "sbev_msmv_bwd_det: bad sizes");
    long long base = 0;
    for (int l = 0; l < L; ++l) {
        SBEV_REQUIRE(hw[2 * l] > 0 && hw[2 * l + 1] > 0, SBEV_ERR_INVALID, "level %d has non-positive size", l);
        ps->base[l] = base;
        base += (long long)Bp * N * hw[2 * l] * hw[2 * l + 1];
    }
    for (int l = L; l <= SBEV_MAX_LEVELS; ++l) ps->base[l] = base;
    *n_blocks = (base + SCAN_ITEMS - 1) / SCAN_ITEMS;
    *n_ids = (long long)Bp * Q * P * L * 4;
    SBEV_REQUIRE((long long)Bp * Q * P * 4 < (1ll << 31) && *n_ids < (1ll << 31), SBEV_ERR_UNSUPPORTED, "sbev_msmv_bwd_det: too many sample points");
    return SBEV_OK;
}

extern "C" long long sbev_msmv_bwd_det_workspace(const int* hw, int L, int Bp, int N, int Q, int P) {
    PixelSpace ps; long long nb, ni;
    if (det_layout(hw, L, Bp, N, Q, P, &ps, &nb, &ni)) return -1;
    return 4 * (2 * ps.base[L] + 1 + nb + ni) + 64;
}

extern "C" int sbev_msmv_bwd_det(const float* grad_out, const float* const* feats, const int* hw, int L,
                                 const float* loc, const float* w, int Bp, int N, int C, int Q, int P,
                                 float* const* grad_feats, float* grad_loc, float* grad_w,
                                 void* workspace, long long workspace_bytes, void* stream) {
    SBEV_REQUIRE(feats && grad_feats, SBEV_ERR_INVALID, "sbev_msmv_bwd_det: null pointer");
    SBEV_REQUIRE((long long)Bp * Q * P == 0 || (grad_out && loc && w && grad_loc && grad_w), SBEV_ERR_INVALID,
                 "sbev_msmv_bwd_det: null pointer");          // no sample points: only the (zero) grad_feats are produced
    SBEV_REQUIRE(C == 64, SBEV_ERR_UNSUPPORTED, "sbev_msmv_bwd_det: needs C = 64 (got %d)", C);
    SBEV_REQUIRE(P <= SBEV_MAX_POINTS, SBEV_ERR_INVALID, "num_point exceed limits (%d > %d)", P, SBEV_MAX_POINTS);
    PixelSpace ps; long long nb, ni;
    int rc = det_layout(hw, L, Bp, N, Q, P, &ps, &nb, &ni);
    if (rc) return rc;
    const long long npix = ps.base[L];
    SBEV_REQUIRE(workspace && workspace_bytes >= 4 * (2 * npix + 1 + nb + ni) + 64 && (reinterpret_cast<uintptr_t>(workspace) & 3) == 0,
                 SBEV_ERR_INVALID, "sbev_msmv_bwd_det: workspace too small (see sbev_msmv_bwd_det_workspace)");
    LevelSet lv;
    rc = fill_levels(lv, feats, hw, L);
    if (rc) return rc;
    GradLevelSet glv;
    for (int l = 0; l < SBEV_MAX_LEVELS; ++l) glv.ptr[l] = nullptr;
    for (int l = 0; l < L; ++l) {
        SBEV_REQUIRE(grad_feats[l] != nullptr && (reinterpret_cast<uintptr_t>(grad_feats[l]) & 15) == 0, SBEV_ERR_INVALID,
                     "grad_feats[%d] is null or not 16-byte aligned", l);
        glv.ptr[l] = grad_feats[l];
    }
    if (npix == 0) return SBEV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int* cnt = reinterpret_cast<int*>(workspace);
    int* off = cnt + npix;
    int* bsum = off + npix + 1;
    int* ids = bsum + nb;
    const long long npts = (long long)Bp * Q * P;
    cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)npix, st);
    if (npts > 0) msmv_bwd_bin_kernel<false><<<grid_for(npts * L, 256), 256, 0, st>>>(lv, ps, L, loc, npts, N, Q * P, cnt, nullptr, nullptr);
    scan_block_sums_kernel<<<(unsigned)nb, 256, 0, st>>>(cnt, npix, bsum);
    scan_bsums_kernel<<<1, 256, 0, st>>>(bsum, (int)nb);
    scan_apply_kernel<<<(unsigned)nb, 256, 0, st>>>(cnt, npix, bsum, off);
    if (npts > 0) msmv_bwd_bin_kernel<true><<<grid_for(npts * L, 256), 256, 0, st>>>(lv, ps, L, loc, npts, N, Q * P, cnt, off, ids);
    const int rgrid = grid_for(npix, 16);
#define SBEV_LAUNCH_DET(LL)                                                                                                   \
    case LL:                                                                                                                  \
        msmv_bwd_reduce_kernel<LL><<<rgrid, 256, 0, st>>>(lv, glv, ps, off, ids, grad_out, loc, w, P);                         \
        if (npts > 0) msmv_bwd_c64_kernel<LL, false><<<grid_for(npts, 16), 256, 0, st>>>(lv, glv, grad_out, loc, w, Bp, N, Q, P, grad_loc, grad_w); \
        break;
    switch (L) { SBEV_LAUNCH_DET(1) SBEV_LAUNCH_DET(2) SBEV_LAUNCH_DET(3) SBEV_LAUNCH_DET(4) SBEV_LAUNCH_DET(5) }
#undef SBEV_LAUNCH_DET
    return check_launch("sbev_msmv_bwd_det");
}

static int launch_sampling4d(const float* const* feats, const int* hw, int L,
                             const int64_t* stride_bt, const int64_t* stride_g,
                             const int64_t* stride_v, const int64_t* stride_px,
                             const float* points, const float* velocity, int ld_vel, const float* time_diff,
                             const float* lidar2img, const float* scale_w,
                             int B, int T, int t0, int Tl, int G, int N, int C, int Q, int P,
                             float image_h, float image_w, float eps,
                             float* const* outs, int n_out, bool full_out, float* loc_out, void* stream, const char* who,
                             int q_per_rank = 0) {
    SBEV_REQUIRE(ld_vel >= 2, SBEV_ERR_INVALID, "%s: ld_vel must be >= 2", who);
    SBEV_REQUIRE(t0 >= 0 && Tl >= 0 && t0 + Tl <= T, SBEV_ERR_INVALID, "%s: frame window [%d,%d) outside [0,%d)", who, t0, t0 + Tl, T);
    SBEV_REQUIRE(feats && stride_bt && stride_g && stride_v && stride_px && points && velocity && time_diff &&
                 lidar2img && scale_w && outs, SBEV_ERR_INVALID, "%s: null pointer", who);
    SBEV_REQUIRE(n_out >= 1 && n_out <= SBEV_MAX_PEERS, SBEV_ERR_INVALID, "%s: between 1 and %d output buffers (got %d)", who, SBEV_MAX_PEERS, n_out);
    SBEV_REQUIRE(B >= 0 && T > 0 && G > 0 && N > 0 && Q >= 0 && P > 0, SBEV_ERR_INVALID, "%s: bad sizes", who);
    SBEV_REQUIRE(C == 64, SBEV_ERR_UNSUPPORTED, "%s: channels per group must be 64 (got %d)", who, C);
    SBEV_REQUIRE(N <= 16, SBEV_ERR_UNSUPPORTED, "%s: at most 16 views (got %d)", who, N);
    SBEV_REQUIRE((reinterpret_cast<uintptr_t>(lidar2img) & 15) == 0, SBEV_ERR_INVALID, "lidar2img not 16-byte aligned");
    LevelSet lv;
    int rc = fill_levels(lv, feats, hw, L);
    if (rc) return rc;
    for (int l = 0; l < L; ++l) {
        lv.s_bt[l] = stride_bt[l]; lv.s_g[l] = stride_g[l]; lv.s_v[l] = stride_v[l]; lv.s_px[l] = stride_px[l];
        SBEV_REQUIRE(((stride_bt[l] | stride_g[l] | stride_v[l] | stride_px[l]) & 3) == 0, SBEV_ERR_INVALID,
                     "level %d strides must be multiples of 4 floats", l);
    }
    const long long total = (long long)B * Tl * G * Q * P;
    if (total == 0) return SBEV_OK;
    SBEV_REQUIRE((long long)B * Tl * G <= 65535 && (long long)Q * P < (1ll << 30), SBEV_ERR_UNSUPPORTED, "%s: too many slices / samples per slice", who);
    for (int l = 0; l < L; ++l)
        SBEV_REQUIRE((long long)lv.H[l] * lv.W[l] * stride_px[l] < (1ll << 31), SBEV_ERR_UNSUPPORTED, "level %d too large for 32-bit pixel offsets", l);
    FusedParams prm{};
    prm.points = points; prm.velocity = velocity; prm.time_diff = time_diff; prm.lidar2img = lidar2img; prm.scale_w = scale_w;
    for (int w = 0; w < n_out; ++w) {
        SBEV_REQUIRE(outs[w] != nullptr && (reinterpret_cast<uintptr_t>(outs[w]) & 15) == 0, SBEV_ERR_INVALID, "%s: out[%d] null or not 16-byte aligned", who, w);
        prm.out[w] = outs[w];
    }
    prm.loc_out = loc_out;
    prm.B = B; prm.T = T; prm.G = G; prm.N = N; prm.Q = Q; prm.P = P;
    prm.image_h = image_h; prm.image_w = image_w; prm.eps = eps;
    prm.t0 = t0; prm.Tl = Tl; prm.n_out = n_out; prm.ld_vel = ld_vel;
    prm.o0 = full_out ? 0 : t0; prm.To = full_out ? T : Tl;
    prm.q_per_rank = q_per_rank;
    if (q_per_rank > 0) {
        SBEV_REQUIRE(B == 1, SBEV_ERR_UNSUPPORTED, "%s: the owner form needs B == 1 (got %d)", who, B);
        SBEV_REQUIRE((long long)q_per_rank * n_out >= Q, SBEV_ERR_INVALID, "%s: %d ranks x %d queries do not cover Q = %d", who, n_out, q_per_rank, Q);
    }
    // 0 = 16 lanes/point, all levels in flight (2 CTAs/SM); 1 = 16 lanes/point, two levels at a time (3 CTAs/SM);
    // 2 = 8 lanes/point (8 channels per lane), two levels at a time (needs N <= 8 views); 3 = 2 + L2 prefetch of the next level pair;
    // 4 = 2 + warp-uniform skip of level blocks without a live tap; 5 = 4 at 3 CTAs/SM; 6 = 4 with one level at a time, 4 CTAs/SM
    int variant = get_option(OPT_GATHER_VARIANT);
    if (variant >= 2 && N > 8) variant = 1;
    const int ppb = variant >= 2 ? 32 : 16;
    const dim3 grid((Q * P + ppb - 1) / ppb, B * Tl * G);
#define SBEV_LAUNCH_FUSED(LL)                                                                                     \
    case LL:                                                                                                      \
        if (variant == 6) launch_pdl(sampling4d_c64_kernel<LL, 1, 4, 8, false, true>, grid, dim3(256), 0, (cudaStream_t)stream, lv, prm);      \
        else if (variant == 5 && LL >= 2) launch_pdl(sampling4d_c64_kernel<LL, 2, 3, 8, false, true>, grid, dim3(256), 0, (cudaStream_t)stream, lv, prm);      \
        else if (variant == 4 && LL >= 2) launch_pdl(sampling4d_c64_kernel<LL, 2, 2, 8, false, true>, grid, dim3(256), 0, (cudaStream_t)stream, lv, prm);      \
        else if (variant == 3 && LL >= 3) launch_pdl(sampling4d_c64_kernel<LL, 2, 2, 8, true>, grid, dim3(256), 0, (cudaStream_t)stream, lv, prm);      \
        else if (variant >= 2 && LL >= 2) launch_pdl(sampling4d_c64_kernel<LL, 2, 2, 8>, grid, dim3(256), 0, (cudaStream_t)stream, lv, prm);      \
        else if (variant >= 2) launch_pdl(sampling4d_c64_kernel<LL, LL, 2, 8>, grid, dim3(256), 0, (cudaStream_t)stream, lv, prm);           \
        else if (variant == 1 && LL >= 3) launch_pdl(sampling4d_c64_kernel<LL, 2, 3, 16>, grid, dim3(256), 0, (cudaStream_t)stream, lv, prm); \
        else launch_pdl(sampling4d_c64_kernel<LL, LL, 1, 16>, grid, dim3(256), 0, (cudaStream_t)stream, lv, prm);               \
        break;
    switch (L) { SBEV_LAUNCH_FUSED(1) SBEV_LAUNCH_FUSED(2) SBEV_LAUNCH_FUSED(3) SBEV_LAUNCH_FUSED(4) SBEV_LAUNCH_FUSED(5) }
#undef SBEV_LAUNCH_FUSED
    return check_launch(who);
}

extern "C" int sbev_sampling4d_fwd(const float* const* feats, const int* hw, int L,
                                   const int64_t* stride_bt, const int64_t* stride_g,
                                   const int64_t* stride_v, const int64_t* stride_px,
                                   const float* points, const float* velocity, const float* time_diff,
                                   const float* lidar2img, const float* scale_w,
                                   int B, int T, int G, int N, int C, int Q, int P,
                                   float image_h, float image_w, float eps,
                                   float* out, float* loc_out, void* stream) {
    return launch_sampling4d(feats, hw, L, stride_bt, stride_g, stride_v, stride_px, points, velocity, 2, time_diff, lidar2img, scale_w,
                             B, T, 0, T, G, N, C, Q, P, image_h, image_w, eps, &out, 1, false, loc_out, stream, "sbev_sampling4d_fwd");
}

extern "C" int sbev_sampling4d_window_fwd(const float* const* feats, const int* hw, int L,
                                          const int64_t* stride_bt, const int64_t* stride_g,
                                          const int64_t* stride_v, const int64_t* stride_px,
                                          const float* points, const float* velocity, int ld_vel, const float* time_diff,
                                          const float* lidar2img, const float* scale_w,
                                          int B, int T, int t0, int Tl, int G, int N, int C, int Q, int P,
                                          float image_h, float image_w, float eps,
                                          float* out, float* loc_out, void* stream) {
    return launch_sampling4d(feats, hw, L, stride_bt, stride_g, stride_v, stride_px, points, velocity, ld_vel, time_diff, lidar2img, scale_w,
                             B, T, t0, Tl, G, N, C, Q, P, image_h, image_w, eps, &out, 1, false, loc_out, stream, "sbev_sampling4d_window_fwd");
}

extern "C" int sbev_sampling4d_scatter_fwd(const float* const* feats, const int* hw, int L,
                                           const int64_t* stride_bt, const int64_t* stride_g,
                                           const int64_t* stride_v, const int64_t* stride_px,
                                           const float* points, const float* velocity, int ld_vel, const float* time_diff,
                                           const float* lidar2img, const float* scale_w,
                                           int B, int T, int t0, int Tl, int G, int N, int C, int Q, int P,
                                           float image_h, float image_w, float eps,
                                           float* const* outs, int n_out, float* loc_out, void* stream) {
    return launch_sampling4d(feats, hw, L, stride_bt, stride_g, stride_v, stride_px, points, velocity, ld_vel, time_diff, lidar2img, scale_w,
                             B, T, t0, Tl, G, N, C, Q, P, image_h, image_w, eps, outs, n_out, true, loc_out, stream, "sbev_sampling4d_scatter_fwd");
}

extern "C" int sbev_sampling4d_owner_fwd(const float* const* feats, const int* hw, int L,
                                         const int64_t* stride_bt, const int64_t* stride_g,
                                         const int64_t* stride_v, const int64_t* stride_px,
                                         const float* points, const float* velocity, int ld_vel, const float* time_diff,
                                         const float* lidar2img, const float* scale_w,
                                         int T, int t0, int Tl, int G, int N, int C, int Q, int P,
                                         float image_h, float image_w, float eps,
                                         float* const* outs, int n_ranks, int q_per_rank, void* stream) {
    SBEV_REQUIRE(q_per_rank > 0, SBEV_ERR_INVALID, "sbev_sampling4d_owner_fwd: q_per_rank must be positive");
    return launch_sampling4d(feats, hw, L, stride_bt, stride_g, stride_v, stride_px, points, velocity, ld_vel, time_diff, lidar2img, scale_w,
                             1, T, t0, Tl, G, N, C, Q, P, image_h, image_w, eps, outs, n_ranks, true, nullptr, stream, "sbev_sampling4d_owner_fwd", q_per_rank);
}
