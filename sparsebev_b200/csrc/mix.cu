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
#include <stdlib.h>
#include <mutex>

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


// ------------------------------------------------------------------------------------------------
// Tensor-core version: both small matmuls on mma.sync.m16n8k16 (bf16 inputs, fp32 accumulate) with the
// bf16x3 split (a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi, ~2^-17 relative) so the result stays fp32-grade.
// Every operand is split ONCE while it is staged into shared memory (hi and lo bf16 copies, rows padded so
// ldmatrix is bank-conflict free); stage 1 = x[PK x 64] . M[64 x 64] with warp w owning output columns
// 8w..8w+7; stage 2 = S[128 x PK] . h[PK x 64] with warp w owning rows 16w..16w+15; both LayerNorms are
// block reductions over the accumulator fragments; the result leaves through shared memory as 128 B rows.
constexpr int MX_LD = 72;            // bf16 row stride of the K=64 operands (x, M^T): 144 B, == 4 words mod 32

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
// (count, mean, M2) running statistics and their pairwise merge (Chan et al.): one block reduction gives an
// accurately centred variance, instead of one reduction for the mean and a second for the squared deviations.
struct Stat { float n, mean, m2; };
__device__ __forceinline__ Stat stat_merge(const Stat& a, const Stat& b) {
    Stat r;
    r.n = a.n + b.n;
    if (r.n == 0.f) { r.mean = 0.f; r.m2 = 0.f; return r; }
    const float d = b.mean - a.mean, fb = b.n / r.n;
    r.mean = a.mean + d * fb;
    r.m2 = a.m2 + b.m2 + d * d * a.n * fb;
    return r;
}
__device__ __forceinline__ Stat block_stat_256(Stat s, float* red /* [8*3] */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Stat t;
        t.n = __shfl_xor_sync(0xffffffffu, s.n, o); t.mean = __shfl_xor_sync(0xffffffffu, s.mean, o); t.m2 = __shfl_xor_sync(0xffffffffu, s.m2, o);
        s = stat_merge(s, t);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) { red[warp * 3] = s.n; red[warp * 3 + 1] = s.mean; red[warp * 3 + 2] = s.m2; }
    __syncthreads();
    Stat t;
    t.n = red[(lane & 7) * 3]; t.mean = red[(lane & 7) * 3 + 1]; t.m2 = red[(lane & 7) * 3 + 2];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        Stat u;
        u.n = __shfl_xor_sync(0xffffffffu, t.n, o); u.mean = __shfl_xor_sync(0xffffffffu, t.mean, o); u.m2 = __shfl_xor_sync(0xffffffffu, t.m2, o);
        t = stat_merge(t, u);
    }
    return t;
}
// Same reduction when every thread contributes the SAME number of elements N0 (mix_tma_kernel: in_points == 32): both sides
// of every pairwise merge then hold n = N0 * 2^step elements, so  mean = a + d/2,  M2 = M2a + M2b + d^2 * n/2  -- no
// division, no count to shuffle, no empty-side branch (the general merge was ~25 % of the kernel's instructions).
template <int N0>
__device__ __forceinline__ Stat block_stat_256_eq(float mean, float m2, float* red /* [8*2] */) {
    float n = (float)N0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float bm = __shfl_xor_sync(0xffffffffu, mean, o), b2 = __shfl_xor_sync(0xffffffffu, m2, o);
        const float d = bm - mean;
        mean += 0.5f * d;
        m2 = (m2 + b2) + d * d * (0.5f * n);
        n *= 2.f;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) { red[warp * 2] = mean; red[warp * 2 + 1] = m2; }
    __syncthreads();
    mean = red[(lane & 7) * 2]; m2 = red[(lane & 7) * 2 + 1];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        const float bm = __shfl_xor_sync(0xffffffffu, mean, o), b2 = __shfl_xor_sync(0xffffffffu, m2, o);
        const float d = bm - mean;
        mean += 0.5f * d;
        m2 = (m2 + b2) + d * d * (0.5f * n);
        n *= 2.f;
    }
    Stat r; r.n = n; r.mean = mean; r.m2 = m2;
    return r;
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                     const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    mma_bf16(d, al, bh);
    mma_bf16(d, ah, bl);
    mma_bf16(d, ah, bh);
}

// NM = number of 16-row tiles of the in_points dimension (PK = 16*NM >= Pin).
// Persistent CTAs: while item i is being multiplied, the raw fp32 operands of item i + gridDim.x (one contiguous
// [M | S] block of the params row + the x tile) are already in flight into a shared staging buffer via
// cp.async.bulk + mbarrier, so the global-load latency of the 40 KB operand set is hidden behind the MMAs.
template <int NM>
__global__ void __launch_bounds__(256)
mix_mma_kernel(const float* __restrict__ params, const float* __restrict__ x, int G, int Pin, int num_items,
               __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo, float* __restrict__ y_f32) {
    constexpr int PK = 16 * NM;
    constexpr int PS = PK + 8;                       // bf16 row stride of the K=PK operands (S, h^T)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int per_group = MIX_C * MIX_C + MIX_POUT * Pin;                  // floats: [ M 64x64 | S 128xPin ]
    float* raw_ms = reinterpret_cast<float*>(smem_raw);                     // staged raw fp32 operands of the current item
    float* raw_x = raw_ms + per_group;                                     // [Pin][64]
    __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(raw_x + Pin * MIX_C);   // [PK][72]
    __nv_bfloat16* xl = xh + PK * MX_LD;
    __nv_bfloat16* mh = xl + PK * MX_LD;                                   // M [64 c][72 (c')]  natural layout, read with ldmatrix.trans
    __nv_bfloat16* ml = mh + MIX_C * MX_LD;
    __nv_bfloat16* sh = ml + MIX_C * MX_LD;                                // S [128][PS]
    __nv_bfloat16* sl = sh + MIX_POUT * PS;
    __nv_bfloat16* hh = sl + MIX_POUT * PS;                                // h [PK p][72 (c')]  natural layout, read with ldmatrix.trans
    __nv_bfloat16* hl = hh + PK * MX_LD;
    __nv_bfloat16* oh = xh;                                                // epilogue staging [128][72] x2 re-uses the operand arrays
    __nv_bfloat16* ol = oh + MIX_POUT * MX_LD;
    __shared__ float red[24];
    __shared__ uint64_t full_bar;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g8 = lane >> 2, t4 = lane & 3;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full_bar);
    const uint32_t bytes_ms = (uint32_t)per_group * 4u, bytes_x = (uint32_t)Pin * MIX_C * 4u;

    auto prefetch = [&](long long item) {             // one thread: arm the barrier, launch the two bulk copies
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes_ms + bytes_x) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(raw_ms)), "l"(params + item * per_group), "r"(bytes_ms), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(raw_x)), "l"(x + item * Pin * MIX_C), "r"(bytes_x), "r"(bar) : "memory");
    };
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    pdl_wait();
    pdl_trigger();
    if (tid == 0 && (long long)blockIdx.x < num_items) prefetch(blockIdx.x);
    __syncthreads();

    // ldmatrix lane addressing: A (16x16): row = r + 8*(id&1), col = 8*(id>>1); B (8 n x 16 k): row = r, col = 8*(id&1)
    const int lm_r = lane & 7, lm_id = lane >> 3;
    const int a_row = lm_r + 8 * (lm_id & 1), a_col = 8 * (lm_id >> 1);

    uint32_t parity = 0;
    for (long long qg = blockIdx.x; qg < num_items; qg += gridDim.x, parity ^= 1) {
        {   // wait for this item's raw operands (bounded spin: a protocol bug traps instead of hanging)
            uint32_t done = 0;
            for (uint32_t spins = 0; !done; ++spins) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                if (!done && spins > (1u << 26)) __trap();
            }
        }
        // ---- split the operands once into bf16 (hi, lo) arrays laid out for ldmatrix
        for (int i = tid; i < PK * 16; i += 256) {                  // x: rows p (zero beyond Pin), 4 channels per thread
            const int p = i >> 4, c = (i & 15) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < Pin) v = *reinterpret_cast<const float4*>(raw_x + p * MIX_C + c);
            uint32_t h0, l0, h1, l1;
            split2(v.x, v.y, h0, l0); split2(v.z, v.w, h1, l1);
            *reinterpret_cast<uint2*>(xh + p * MX_LD + c) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(xl + p * MX_LD + c) = make_uint2(l0, l1);
        }
        for (int i = tid; i < MIX_C * 16; i += 256) {               // M[c][c'] kept as is (k-major rows), 4 columns per thread
            const int c = i >> 4, n = (i & 15) * 4;
            const float4 v = *reinterpret_cast<const float4*>(raw_ms + c * MIX_C + n);
            uint32_t h0, l0, h1, l1;
            split2(v.x, v.y, h0, l0); split2(v.z, v.w, h1, l1);
            *reinterpret_cast<uint2*>(mh + c * MX_LD + n) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(ml + c * MX_LD + n) = make_uint2(l0, l1);
        }
        {
            const float* raw_s = raw_ms + MIX_C * MIX_C;
            const int q4 = Pin >> 2;                                // float4 groups per S row (Pin % 4 == 0)
            for (int i = tid; i < MIX_POUT * (PK / 4); i += 256) {
                const int o = i / (PK / 4), pq = i - o * (PK / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pq < q4) v = *reinterpret_cast<const float4*>(raw_s + o * Pin + pq * 4);
                uint32_t h0, l0, h1, l1;
                split2(v.x, v.y, h0, l0); split2(v.z, v.w, h1, l1);
                *reinterpret_cast<uint2*>(sh + o * PS + pq * 4) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(sl + o * PS + pq * 4) = make_uint2(l0, l1);
            }
        }
        __syncthreads();
        if (tid == 0 && qg + gridDim.x < num_items) {               // staging buffer is free again: fetch the next item now
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            prefetch(qg + gridDim.x);
        }

        // ---- stage 1: h[p][c'] = sum_c x[p][c] M[c][c'], warp w owns columns c' = 8w..8w+7, all NM row tiles
        float acc1[NM][4];
#pragma unroll
        for (int m = 0; m < NM; ++m) { acc1[m][0] = acc1[m][1] = acc1[m][2] = acc1[m][3] = 0.f; }
#pragma unroll
        for (int k0 = 0; k0 < MIX_C; k0 += 16) {
            uint32_t bh[2], bl[2];
            ldsm_x2_trans(bh, mh + (k0 + (lane & 15)) * MX_LD + 8 * warp);
            ldsm_x2_trans(bl, ml + (k0 + (lane & 15)) * MX_LD + 8 * warp);
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                uint32_t ah[4], al[4];
                ldsm_x4(ah, xh + (16 * m + a_row) * MX_LD + k0 + a_col);
                ldsm_x4(al, xl + (16 * m + a_row) * MX_LD + k0 + a_col);
                mma3(acc1[m], ah, al, bh, bl);
            }
        }
        {   // LayerNorm over the valid Pin x 64 block + ReLU, written as h (hi, lo) in natural [p][c'] layout
            Stat st;
            {
                float cnt = 0.f, sum = 0.f;
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    if (16 * m + g8 < Pin) { sum += acc1[m][0] + acc1[m][1]; cnt += 2.f; }
                    if (16 * m + g8 + 8 < Pin) { sum += acc1[m][2] + acc1[m][3]; cnt += 2.f; }
                }
                const float lm = cnt > 0.f ? sum / cnt : 0.f;
                float m2 = 0.f;
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    if (16 * m + g8 < Pin) { const float a = acc1[m][0] - lm, b = acc1[m][1] - lm; m2 += a * a + b * b; }
                    if (16 * m + g8 + 8 < Pin) { const float a = acc1[m][2] - lm, b = acc1[m][3] - lm; m2 += a * a + b * b; }
                }
                st.n = cnt; st.mean = lm; st.m2 = m2;
            }
            st = block_stat_256(st, red);
            const float mean = st.mean, rstd = rsqrtf(st.m2 / (float)(Pin * MIX_C) + 1e-5f);
#pragma unroll
            for (int m = 0; m < NM; ++m)
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const int p = 16 * m + g8 + 8 * hrow, c = 8 * warp + 2 * t4;
                    const bool ok = p < Pin;
                    const float v0 = ok ? fmaxf((acc1[m][2 * hrow] - mean) * rstd, 0.f) : 0.f;
                    const float v1 = ok ? fmaxf((acc1[m][2 * hrow + 1] - mean) * rstd, 0.f) : 0.f;
                    uint32_t h, l;
                    split2(v0, v1, h, l);
                    *reinterpret_cast<uint32_t*>(hh + p * MX_LD + c) = h;
                    *reinterpret_cast<uint32_t*>(hl + p * MX_LD + c) = l;
                }
        }
        __syncthreads();

        // ---- stage 2: y[o][c'] = sum_p S[o][p] h[p][c'], warp w owns rows o = 16w..16w+15, all 8 column tiles
        float acc2[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) { acc2[n][0] = acc2[n][1] = acc2[n][2] = acc2[n][3] = 0.f; }
#pragma unroll
        for (int k0 = 0; k0 < PK; k0 += 16) {
            uint32_t ah[4], al[4];
            ldsm_x4(ah, sh + (16 * warp + a_row) * PS + k0 + a_col);
            ldsm_x4(al, sl + (16 * warp + a_row) * PS + k0 + a_col);
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                uint32_t bh[2], bl[2];
                ldsm_x2_trans(bh, hh + (k0 + (lane & 15)) * MX_LD + 8 * n);
                ldsm_x2_trans(bl, hl + (k0 + (lane & 15)) * MX_LD + 8 * n);
                mma3(acc2[n], ah, al, bh, bl);
            }
        }
        Stat st2;
        {
            float sum = 0.f;
#pragma unroll
            for (int n = 0; n < 8; ++n) sum += (acc2[n][0] + acc2[n][1]) + (acc2[n][2] + acc2[n][3]);
            const float lm = sum * (1.f / 32.f);
            float m2 = 0.f;
#pragma unroll
            for (int n = 0; n < 8; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) { const float d = acc2[n][i] - lm; m2 += d * d; }
            st2.n = 32.f; st2.mean = lm; st2.m2 = m2;
        }
        st2 = block_stat_256(st2, red);                                      // (its barriers also order the smem re-use below)
        const float mean = st2.mean, rstd = rsqrtf(st2.m2 / (float)(MIX_POUT * MIX_C) + 1e-5f);

        // ---- epilogue: ReLU(LN) -> (hi, lo) staged in shared memory (operand arrays are dead now) -> 16 B coalesced stores
        const long long obase = qg * (MIX_POUT * MIX_C);
        __syncthreads();
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int o = 16 * warp + g8 + 8 * hrow, c = 8 * n + 2 * t4;
                const float v0 = fmaxf((acc2[n][2 * hrow] - mean) * rstd, 0.f), v1 = fmaxf((acc2[n][2 * hrow + 1] - mean) * rstd, 0.f);
                uint32_t h, l;
                split2(v0, v1, h, l);
                *reinterpret_cast<uint32_t*>(oh + o * MX_LD + c) = h;
                *reinterpret_cast<uint32_t*>(ol + o * MX_LD + c) = l;
                if (y_f32) *reinterpret_cast<float2*>(y_f32 + obase + o * MIX_C + c) = make_float2(v0, v1);
            }
        __syncthreads();
        if (y_hi) {
            for (int i = tid; i < MIX_POUT * 8; i += 256) {
                const int o = i >> 3, ch = (i & 7) * 8;
                *reinterpret_cast<uint4*>(y_hi + obase + o * MIX_C + ch) = *reinterpret_cast<const uint4*>(oh + o * MX_LD + ch);
                if (y_lo) *reinterpret_cast<uint4*>(y_lo + obase + o * MIX_C + ch) = *reinterpret_cast<const uint4*>(ol + o * MX_LD + ch);
            }
        }
        __syncthreads();                                                    // staging rows read before the next item's split overwrites them
    }
}


// ------------------------------------------------------------------------------------------------
// TMA-fed version (in_points == 32): the parameter GEMM already stored the dynamic parameters as bf16 (hi, lo)
// (sbev_gemm_bf16_tn_split), so M [64x64] and S [128x32] of an item arrive by four 2-D TMA loads -- 128-byte / 64-byte
// swizzled, i.e. directly in the bank-conflict-free layout ldmatrix wants -- into a double-buffered operand set; no
// conversion pass for the 8192 parameters of an item, only the 32x64 x tile is split on the fly.
struct MixMaps { CUtensorMap m_hi, m_lo, s_hi, s_lo, y_hi, y_lo; };   // y_*: [items*128 rows][64 bf16], box 128 rows, 128 B swizzle

__global__ void __launch_bounds__(256, 2)
mix_tma_kernel(const __grid_constant__ MixMaps maps, const float* __restrict__ x, int num_items, int G, int order,
               __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo, float* __restrict__ y_f32) {
    constexpr int Pin = 32, PK = 32;
    constexpr int SET_BYTES = 4 * 8192;                       // M_hi | M_lo | S_hi | S_lo
    constexpr int XRAW_BYTES = Pin * MIX_C * 4;               // 8 KB
    extern __shared__ uint8_t mt_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mt_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sets = base;                                                           // [2][SET_BYTES]
    float* xraw = reinterpret_cast<float*>(base + 2 * SET_BYTES);                   // [2][Pin*64]
    __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(base + 2 * SET_BYTES + 2 * XRAW_BYTES);   // [32][72]
    __nv_bfloat16* xl = xh + PK * MX_LD;
    __nv_bfloat16* hh = xl + PK * MX_LD;                                            // h [32 p][72 (c')]
    __nv_bfloat16* hl = hh + PK * MX_LD;
    __shared__ float red[24];
    __shared__ uint64_t full_bar[2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g8 = lane >> 2, t4 = lane & 3;

    auto prefetch = [&](long long item, int slot) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full_bar[slot]);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sets + slot * SET_BYTES);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(SET_BYTES + XRAW_BYTES) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst), "l"(&maps.m_hi), "r"(bar), "r"(0), "r"((int)(item * 128)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst + 8192), "l"(&maps.m_lo), "r"(bar), "r"(0), "r"((int)(item * 128)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst + 16384), "l"(&maps.s_hi), "r"(bar), "r"(0), "r"((int)(item * 256 + 128)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst + 24576), "l"(&maps.s_lo), "r"(bar), "r"(0), "r"((int)(item * 256 + 128)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(xraw + slot * Pin * MIX_C)), "l"(x + item * Pin * MIX_C), "r"(XRAW_BYTES), "r"(bar) : "memory");
    };
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&full_bar[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    pdl_wait();
    pdl_trigger();
    // Work order.  0: items ascending (query-major).  1: group-major, LAST group first -- the parameter GEMM writes its
    // N-tiles (= groups) in ascending order, so the highest groups' parameters are the ones still resident in L2 when this
    // kernel starts; and this kernel then finishes with group 0, whose output is what the out-projection's first split-K
    // slices read.
    const int BQ = num_items / G;
    auto item_of = [&](long long i) -> long long {
        if (!order) return i;
        const int gi = (int)(i / BQ);
        return (i - (long long)gi * BQ) * G + (G - 1 - gi);
    };
    if (tid == 0 && (long long)blockIdx.x < num_items) prefetch(item_of(blockIdx.x), 0);
    __syncthreads();

    const int lm_r = lane & 7, lm_id = lane >> 3;
    const int a_row = lm_r + 8 * (lm_id & 1), a_col = 8 * (lm_id >> 1);

    int n = 0;
    for (long long it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
        const long long qg = item_of(it);
        const int slot = n & 1;
        {
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full_bar[slot]);
            const uint32_t parity = (n >> 1) & 1;
            uint32_t done = 0;
            for (uint32_t spins = 0; !done; ++spins) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                if (!done && spins > (1u << 26)) __trap();
            }
        }
        const uint32_t mh = (uint32_t)__cvta_generic_to_shared(sets + slot * SET_BYTES);       // M hi: [64 c][128 B], 128B swizzle
        const uint32_t ml = mh + 8192;
        const uint32_t sh = mh + 16384;                                                     // S hi: [128 o][64 B], 64B swizzle
        const uint32_t sl = mh + 24576;
        {   // x tile -> bf16 (hi, lo), padded rows for ldmatrix
            const float* rx = xraw + slot * Pin * MIX_C;
            for (int i = tid; i < Pin * 16; i += 256) {
                const int p = i >> 4, c = (i & 15) * 4;
                const float4 v = *reinterpret_cast<const float4*>(rx + p * MIX_C + c);
                uint32_t h0, l0, h1, l1;
                split2(v.x, v.y, h0, l0); split2(v.z, v.w, h1, l1);
                *reinterpret_cast<uint2*>(xh + p * MX_LD + c) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(xl + p * MX_LD + c) = make_uint2(l0, l1);
            }
        }
        __syncthreads();
        if (tid == 0 && it + gridDim.x < num_items) {        // the other operand set and x buffer are free: fetch the next item now
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // ... once the previous item's output tile has left it
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            prefetch(item_of(it + gridDim.x), slot ^ 1);
        }

        // ---- stage 1: h[p][c'] = sum_c x[p][c] M[c][c'], warp w owns columns c' = 8w..8w+7
        float acc1[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) { acc1[m][0] = acc1[m][1] = acc1[m][2] = acc1[m][3] = 0.f; }
#pragma unroll
        for (int k0 = 0; k0 < MIX_C; k0 += 16) {
            uint32_t bh[2], bl[2];
            const int row = k0 + (lane & 15);                                       // k index = row of M
            const uint32_t off = (uint32_t)(row * 128 + ((warp ^ (row & 7)) << 4));   // 16-byte chunk `warp` = columns 8w..8w+7
            asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(bh[0]), "=r"(bh[1]) : "r"(mh + off));
            asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(bl[0]), "=r"(bl[1]) : "r"(ml + off));
            // fragments of both row tiles first, then the three bf16x3 products with the two accumulators interleaved (the products of
            // ONE accumulator depend on each other; same order per accumulator as mma3: lo.hi, hi.lo, hi.hi)
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                ldsm_x4(ah[m], xh + (16 * m + a_row) * MX_LD + k0 + a_col);
                ldsm_x4(al[m], xl + (16 * m + a_row) * MX_LD + k0 + a_col);
            }
#pragma unroll
            for (int m = 0; m < 2; ++m) mma_bf16(acc1[m], al[m], bh);
#pragma unroll
            for (int m = 0; m < 2; ++m) mma_bf16(acc1[m], ah[m], bl);
#pragma unroll
            for (int m = 0; m < 2; ++m) mma_bf16(acc1[m], ah[m], bh);
        }
        {
            Stat st;
            {
                float sum = 0.f;
#pragma unroll
                for (int m = 0; m < 2; ++m) sum += (acc1[m][0] + acc1[m][1]) + (acc1[m][2] + acc1[m][3]);
                const float lm = sum * 0.125f;
                float m2 = 0.f;
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int i = 0; i < 4; ++i) { const float d = acc1[m][i] - lm; m2 += d * d; }
                st.n = 8.f; st.mean = lm; st.m2 = m2;
            }
            st = block_stat_256_eq<8>(st.mean, st.m2, red);
            const float mean = st.mean, rstd = rsqrtf(st.m2 / (float)(Pin * MIX_C) + 1e-5f);
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const int p = 16 * m + g8 + 8 * hrow, c = 8 * warp + 2 * t4;
                    const float v0 = fmaxf((acc1[m][2 * hrow] - mean) * rstd, 0.f), v1 = fmaxf((acc1[m][2 * hrow + 1] - mean) * rstd, 0.f);
                    uint32_t h, l;
                    split2(v0, v1, h, l);
                    *reinterpret_cast<uint32_t*>(hh + p * MX_LD + c) = h;
                    *reinterpret_cast<uint32_t*>(hl + p * MX_LD + c) = l;
                }
        }
        __syncthreads();

        // ---- stage 2: y[o][c'] = sum_p S[o][p] h[p][c'], warp w owns rows o = 16w..16w+15
        float acc2[8][4];
#pragma unroll
        for (int nn = 0; nn < 8; ++nn) { acc2[nn][0] = acc2[nn][1] = acc2[nn][2] = acc2[nn][3] = 0.f; }
#pragma unroll
        for (int k0 = 0; k0 < PK; k0 += 16) {
            uint32_t ah[4], al[4];
            const int row = 16 * warp + a_row;
            const int chunk = (k0 >> 3) + (lm_id >> 1);                              // 16-byte chunk of the 64-byte S row
            const uint32_t off = (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));   // TMA 64-byte swizzle
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(ah[0]), "=r"(ah[1]), "=r"(ah[2]), "=r"(ah[3]) : "r"(sh + off));
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(al[0]), "=r"(al[1]), "=r"(al[2]), "=r"(al[3]) : "r"(sl + off));
            uint32_t bh[8][2], bl[8][2];
#pragma unroll
            for (int nn = 0; nn < 8; ++nn) {
                ldsm_x2_trans(bh[nn], hh + (k0 + (lane & 15)) * MX_LD + 8 * nn);
                ldsm_x2_trans(bl[nn], hl + (k0 + (lane & 15)) * MX_LD + 8 * nn);
            }
#pragma unroll
            for (int nn = 0; nn < 8; ++nn) mma_bf16(acc2[nn], al, bh[nn]);
#pragma unroll
            for (int nn = 0; nn < 8; ++nn) mma_bf16(acc2[nn], ah, bl[nn]);
#pragma unroll
            for (int nn = 0; nn < 8; ++nn) mma_bf16(acc2[nn], ah, bh[nn]);
        }
        Stat st2;
        {
            float sum = 0.f;
#pragma unroll
            for (int nn = 0; nn < 8; ++nn) sum += (acc2[nn][0] + acc2[nn][1]) + (acc2[nn][2] + acc2[nn][3]);
            const float lm = sum * (1.f / 32.f);
            float m2 = 0.f;
#pragma unroll
            for (int nn = 0; nn < 8; ++nn)
#pragma unroll
                for (int i = 0; i < 4; ++i) { const float d = acc2[nn][i] - lm; m2 += d * d; }
            st2.n = 32.f; st2.mean = lm; st2.m2 = m2;
        }
        st2 = block_stat_256_eq<32>(st2.mean, st2.m2, red);                  // its barriers: every warp is past its last read of this operand set
        const float mean = st2.mean, rstd = rsqrtf(st2.m2 / (float)(MIX_POUT * MIX_C) + 1e-5f);

        // ---- epilogue: ReLU(LN) -> (hi, lo) staged in THIS item's (now dead) operand set, XOR-swizzled 128-byte rows
        uint8_t* oh = sets + slot * SET_BYTES;            // [128 rows][128 B]
        uint8_t* ol = oh + 16384;
        const long long obase = qg * (MIX_POUT * MIX_C);
#pragma unroll
        for (int nn = 0; nn < 8; ++nn)
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int o = 16 * warp + g8 + 8 * hrow, c = 8 * nn + 2 * t4;
                const float v0 = fmaxf((acc2[nn][2 * hrow] - mean) * rstd, 0.f), v1 = fmaxf((acc2[nn][2 * hrow + 1] - mean) * rstd, 0.f);
                uint32_t h, l;
                split2(v0, v1, h, l);
                const uint32_t off = (uint32_t)(o * 128 + ((nn ^ (o & 7)) << 4) + 4 * t4);     // chunk nn = columns 8nn..8nn+7
                *reinterpret_cast<uint32_t*>(oh + off) = h;
                *reinterpret_cast<uint32_t*>(ol + off) = l;
                if (y_f32) *reinterpret_cast<float2*>(y_f32 + obase + o * MIX_C + c) = make_float2(v0, v1);
            }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // staging writes -> visible to the TMA engine
        __syncthreads();
        if (y_hi && tid == 0) {      // the staged tiles are exactly TMA's 128 B-swizzled image of [128 rows][64 bf16]: two tensor stores
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(&maps.y_hi), "r"((uint32_t)__cvta_generic_to_shared(oh)), "r"(0), "r"((int)(qg * MIX_POUT)) : "memory");
            if (y_lo)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                             ::"l"(&maps.y_lo), "r"((uint32_t)__cvta_generic_to_shared(ol)), "r"(0), "r"((int)(qg * MIX_POUT)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        // no second barrier: this set is refilled by the prefetch of the NEXT iteration, which first waits (thread 0,
        // cp.async.bulk.wait_group.read) until the stores above have read it, behind that iteration's own barrier
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // shared memory must outlive the last store
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
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(y_hi);
    __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(y_lo);
    const int grid = BQ * G;
    const int impl = get_option(OPT_MIX_IMPL);   // 1 selects the fp32 FFMA kernel (exact fp32 like the reference)
    // tensor-core path needs raw staging + operand arrays in shared memory; very large in_points (> ~96) fall back to FFMA
    const size_t need_mma = ((size_t)MIX_C * MIX_C + (size_t)MIX_POUT * Pin + (size_t)Pin * MIX_C) * 4 +
                            2 * ((size_t)2 * 16 * ((Pin + 15) / 16 > 4 ? 8 : (Pin + 15) / 16 > 2 ? 4 : (Pin + 15) / 16 > 1 ? 2 : 1) * MX_LD + (size_t)MIX_C * MX_LD +
                                 (size_t)MIX_POUT * (16 * ((Pin + 15) / 16 > 4 ? 8 : (Pin + 15) / 16 > 2 ? 4 : (Pin + 15) / 16 > 1 ? 2 : 1) + 8)) * 2;
    if (impl == 0 && (Pin & 3) == 0 && need_mma <= 220 * 1024) {
        const int NM = (Pin + 15) / 16;
        const int PK = 16 * (NM <= 1 ? 1 : NM <= 2 ? 2 : NM <= 4 ? 4 : 8);
        size_t ops_bytes = 2 * ((size_t)PK * MX_LD + (size_t)MIX_C * MX_LD + (size_t)MIX_POUT * (PK + 8) + (size_t)PK * MX_LD) * 2;
        const size_t out_stage = 2 * (size_t)MIX_POUT * MX_LD * 2;
        if (ops_bytes < out_stage) ops_bytes = out_stage;
        const size_t raw_bytes = ((size_t)MIX_C * MIX_C + (size_t)MIX_POUT * Pin + (size_t)Pin * MIX_C) * 4;
        const size_t smem = raw_bytes + ops_bytes;
        SBEV_REQUIRE((reinterpret_cast<uintptr_t>(params) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, SBEV_ERR_INVALID,
                     "sbev_mix_fwd: params / x must be 16-byte aligned");
        const int num_sms = device_num_sms();
        const int per_sm = smem <= 110 * 1024 ? 2 : 1;
        const int pgrid = grid < per_sm * num_sms ? grid : per_sm * num_sms;
#define SBEV_LAUNCH_MMA(R)                                                                                          \
        do {                                                                                                        \
            if (smem > 48 * 1024) cudaFuncSetAttribute(mix_mma_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            launch_pdl(mix_mma_kernel<R>, dim3(pgrid), dim3(256), smem, st, params, x, G, Pin, grid, hi, lo, y_f32);      \
        } while (0)
        if (PK == 16) SBEV_LAUNCH_MMA(1);
        else if (PK == 32) SBEV_LAUNCH_MMA(2);
        else if (PK == 64) SBEV_LAUNCH_MMA(4);
        else SBEV_LAUNCH_MMA(8);
#undef SBEV_LAUNCH_MMA
        return check_launch("sbev_mix_fwd(mma)");
    }
    const size_t smem = sizeof(float) * ((size_t)MIX_C * MIX_C + (size_t)Pin * MIX_C * 2 + (size_t)Pin * MIX_ST_LD);
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

// params given as the bf16 (hi, lo) pair written by sbev_gemm_bf16_tn_split; in_points must be 32.
extern "C" int sbev_mix_presplit_fwd(const uint16_t* params_hi, const uint16_t* params_lo, const float* x, int BQ, int G, int Pin,
                                     int Pout, int C, uint16_t* y_hi, uint16_t* y_lo, float* y_f32, void* stream) {
    SBEV_REQUIRE(params_hi && params_lo && x && (y_hi || y_f32), SBEV_ERR_INVALID, "sbev_mix_presplit_fwd: null pointer");
    SBEV_REQUIRE(C == MIX_C && Pout == MIX_POUT && Pin == 32, SBEV_ERR_UNSUPPORTED, "sbev_mix_presplit_fwd: needs C=64, out_points=128, in_points=32");
    SBEV_REQUIRE(BQ >= 0 && G > 0, SBEV_ERR_INVALID, "sbev_mix_presplit_fwd: bad sizes");
    SBEV_REQUIRE((reinterpret_cast<uintptr_t>(params_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(params_lo) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(x) & 15) == 0, SBEV_ERR_INVALID, "sbev_mix_presplit_fwd: operands must be 16-byte aligned");
    if (BQ == 0) return SBEV_OK;
    const long long items = (long long)BQ * G;
    SBEV_REQUIRE(items * 256 < (1ll << 31), SBEV_ERR_UNSUPPORTED, "sbev_mix_presplit_fwd: too many items");
    MixMaps maps;
    // M view: rows of 64 bf16 (the item's 64x64 block starts at row 128*item); S view: rows of 32 bf16 (block at row 256*item + 128)
    int rc = make_bf16_map_ex(&maps.m_hi, params_hi, items * 128, 64, 64, 64, 128);   if (rc) return rc;
    rc = make_bf16_map_ex(&maps.m_lo, params_lo, items * 128, 64, 64, 64, 128);       if (rc) return rc;
    rc = make_bf16_map_ex(&maps.s_hi, params_hi, items * 256, 32, 128, 32, 64);       if (rc) return rc;
    rc = make_bf16_map_ex(&maps.s_lo, params_lo, items * 256, 32, 128, 32, 64);       if (rc) return rc;
    if (y_hi) {
        SBEV_REQUIRE((reinterpret_cast<uintptr_t>(y_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(y_lo) & 15) == 0, SBEV_ERR_INVALID,
                     "sbev_mix_presplit_fwd: y_hi / y_lo must be 16-byte aligned");
        rc = make_bf16_map_ex(&maps.y_hi, y_hi, items * MIX_POUT, 64, MIX_POUT, 64, 128);               if (rc) return rc;
        rc = make_bf16_map_ex(&maps.y_lo, y_lo ? y_lo : y_hi, items * MIX_POUT, 64, MIX_POUT, 64, 128); if (rc) return rc;
    } else {
        maps.y_hi = maps.m_hi; maps.y_lo = maps.m_lo;        // unused
    }
    const int num_sms = device_num_sms();
    const size_t smem = 2 * 32768 + 2 * 8192 + 4 * 32 * MX_LD * 2 + 1024;
    SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(mix_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = items < 2 * num_sms ? (int)items : 2 * num_sms;
    launch_pdl(mix_tma_kernel, dim3(grid), dim3(256), smem, (cudaStream_t)stream, maps, x, (int)items, G, get_option(OPT_MIX_ORDER),
               reinterpret_cast<__nv_bfloat16*>(y_hi),
               reinterpret_cast<__nv_bfloat16*>(y_lo), y_f32);
    return check_launch("sbev_mix_presplit_fwd");
}
