// Scale-adaptive self-attention core for sm_100a: flash-style, the [B*H,Q,Q] distance mask of the
// reference is never materialised -- the bias -tau[b,i,h] * ||c_i - c_j|| is rebuilt on the fly from
// the decoded box centres held in shared memory.
//
// Behavioural reference: /root/reference/models/sparsebev_transformer.py:210-248
// (SparseBEVSelfAttention.inner_forward / calc_bbox_dists) + torch.nn.MultiheadAttention's core
// (q scaled by 1/sqrt(head_dim) before the dot product, additive float mask, softmax, @ v).
//
// fp32 throughout (the op is 0.8 GFLOP at Q=900: latency-, not FLOP-bound).  CTA = 256 threads owns a
// 64-query tile of one (batch, head); keys/values stream through shared memory in tiles of 64; S and
// O are register-tiled 4x4 / 4x2 per thread with operands read as float4 from d-major shared tiles.
#include "common.cuh"
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include <mutex>

namespace sbev {

constexpr int SA_BQ = 64;   // queries per CTA
constexpr int SA_BK = 64;   // keys per tile
constexpr int SA_HD = 32;   // head dim (embed 256 / 8 heads)

__global__ void __launch_bounds__(256)
sasa_hd32_kernel(const float* __restrict__ qkv, int ld_qkv, const float* __restrict__ query_bbox, const float* __restrict__ tau, int ld_tau,
                 const uint8_t* __restrict__ dn_mask, float x_lo, float x_hi, float y_lo, float y_hi,
                 int B, int Q, int H, float* __restrict__ out) {
    __shared__ __align__(16) float Qt[SA_HD][SA_BQ];        // [d][q]   (q pre-scaled)
    __shared__ __align__(16) float Kt[SA_HD][SA_BK];        // [d][k]
    __shared__ __align__(16) float Vs[SA_BK][SA_HD];        // [k][d]
    __shared__ __align__(16) float Pt[SA_BK][SA_BQ + 4];    // [k][q]
    __shared__ float qcx[SA_BQ], qcy[SA_BQ], qtau[SA_BQ], kcx[SA_BK], kcy[SA_BK];

    const int D = H * SA_HD;
    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * SA_BQ;
    const int h = blockIdx.y, b = blockIdx.z;
    const float scale = 0.17677669529663687f;               // 1/sqrt(32) rounded to fp32, as math.sqrt(1/E) -> float
    const float* base = qkv + (long long)b * Q * ld_qkv;

    // stage the query tile (transposed, scaled), centres and tau
    for (int i = tid; i < SA_BQ * SA_HD; i += 256) {
        const int q = i >> 5, d = i & 31;
        const int gq = q0 + q;
        Qt[d][q] = (gq < Q) ? __ldg(base + (long long)gq * ld_qkv + h * SA_HD + d) * scale : 0.f;
    }
    if (tid < SA_BQ) {
        const int gq = q0 + tid;
        const bool ok = gq < Q;
        // decode_bbox centre (bbox/utils.py:63-71): c*(hi-lo)+lo, separate multiply and add
        qcx[tid] = ok ? __fadd_rn(__fmul_rn(__ldg(query_bbox + ((long long)b * Q + gq) * 10), __fsub_rn(x_hi, x_lo)), x_lo) : 0.f;
        qcy[tid] = ok ? __fadd_rn(__fmul_rn(__ldg(query_bbox + ((long long)b * Q + gq) * 10 + 1), __fsub_rn(y_hi, y_lo)), y_lo) : 0.f;
        qtau[tid] = ok ? __ldg(tau + ((long long)b * Q + gq) * ld_tau + h) : 0.f;
    }

    const int ty = tid >> 4, tx = tid & 15;       // S rows 4ty..4ty+3, S cols 4tx..4tx+3, O cols 2tx..2tx+1
    float m_run[4], l_run[4], o_acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { m_run[i] = -INFINITY; l_run[i] = 0.f; o_acc[i][0] = 0.f; o_acc[i][1] = 0.f; }

    for (int k0 = 0; k0 < Q; k0 += SA_BK) {
        __syncthreads();      // previous tile fully consumed (also covers the Qt staging on the first trip)
        for (int i = tid; i < SA_BK * SA_HD; i += 256) {
            const int k = i >> 5, d = i & 31;
            const int gk = k0 + k;
            const bool ok = gk < Q;
            const float* row = base + (long long)gk * ld_qkv + h * SA_HD + d;
            Kt[d][k] = ok ? __ldg(row + D) : 0.f;
            Vs[k][d] = ok ? __ldg(row + 2 * D) : 0.f;
        }
        if (tid < SA_BK) {
            const int gk = k0 + tid;
            const bool ok = gk < Q;
            kcx[tid] = ok ? __fadd_rn(__fmul_rn(__ldg(query_bbox + ((long long)b * Q + gk) * 10), __fsub_rn(x_hi, x_lo)), x_lo) : 0.f;
            kcy[tid] = ok ? __fadd_rn(__fmul_rn(__ldg(query_bbox + ((long long)b * Q + gk) * 10 + 1), __fsub_rn(y_hi, y_lo)), y_lo) : 0.f;
        }
        __syncthreads();

        // S = (q*scale) . k
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) s[i][jj] = 0.f;
#pragma unroll 8
        for (int d = 0; d < SA_HD; ++d) {
            const float4 qv = *reinterpret_cast<const float4*>(&Qt[d][4 * ty]);
            const float4 kv = *reinterpret_cast<const float4*>(&Kt[d][4 * tx]);
            const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) s[i][jj] = fmaf(qa[i], ka[jj], s[i][jj]);
        }
        // + bias, online softmax (rows are shared by the 16 threads of a half-warp)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = 4 * ty + i;
            const float cx = qcx[q], cy = qcy[q], tq = qtau[q];
            float mx = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int k = 4 * tx + jj;
                const float dx = cx - kcx[k], dy = cy - kcy[k];
                const float dist = sqrtf(dx * dx + dy * dy);
                float v = s[i][jj] + (-dist) * tq;
                if (dn_mask != nullptr && q0 + q < Q && k0 + k < Q && dn_mask[(long long)(q0 + q) * Q + k0 + k]) v = -INFINITY;
                if (k0 + k >= Q) v = -INFINITY;
                s[i][jj] = v;
                mx = fmaxf(mx, v);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float m_new = fmaxf(m_run[i], mx);
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;       // fully masked so far
            const float alpha = expf(m_run[i] - m_use);                    // exp(-inf) = 0 on the first tile
            float rs = 0.f;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const float p = expf(s[i][jj] - m_use);
                Pt[4 * tx + jj][q] = p;
                rs += p;
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
            l_run[i] = l_run[i] * alpha + rs;
            m_run[i] = m_new;
            o_acc[i][0] *= alpha; o_acc[i][1] *= alpha;
        }
        __syncthreads();
        // O += P V
#pragma unroll 8
        for (int k = 0; k < SA_BK; ++k) {
            const float4 pv = *reinterpret_cast<const float4*>(&Pt[k][4 * ty]);
            const float2 vv = *reinterpret_cast<const float2*>(&Vs[k][2 * tx]);
            o_acc[0][0] = fmaf(pv.x, vv.x, o_acc[0][0]); o_acc[0][1] = fmaf(pv.x, vv.y, o_acc[0][1]);
            o_acc[1][0] = fmaf(pv.y, vv.x, o_acc[1][0]); o_acc[1][1] = fmaf(pv.y, vv.y, o_acc[1][1]);
            o_acc[2][0] = fmaf(pv.z, vv.x, o_acc[2][0]); o_acc[2][1] = fmaf(pv.z, vv.y, o_acc[2][1]);
            o_acc[3][0] = fmaf(pv.w, vv.x, o_acc[3][0]); o_acc[3][1] = fmaf(pv.w, vv.y, o_acc[3][1]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gq = q0 + 4 * ty + i;
        if (gq < Q) {
            const float inv = 1.f / l_run[i];
            float2 o = make_float2(o_acc[i][0] * inv, o_acc[i][1] * inv);
            *reinterpret_cast<float2*>(out + ((long long)b * Q + gq) * D + h * SA_HD + 2 * tx) = o;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Tensor-core version (default): flash-attention-2 structure on mma.sync.m16n8k16 with the bf16x3 split
// (q.k and p.v each as hi.hi + hi.lo + lo.hi, fp32 accumulate) so logits / outputs stay fp32-grade.
// CTA = 4 warps x 16 query rows = 64 queries of one (batch, head); keys / values stream through shared
// memory in tiles of 64, split into bf16 (hi, lo) as they are staged; S = QK^T and O += PV never leave
// registers (the C fragments of S are re-used in place as the A fragments of P).
constexpr int SM_LD = 40;      // bf16 row stride of the 32-wide head-dim tiles: 80 B = 20 words (ldmatrix conflict-free)

__device__ __forceinline__ void sa_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void sa_ldsm_x4(uint32_t (&r)[4], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void sa_ldsm_x2(uint32_t (&r)[2], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void sa_ldsm_x2_trans(uint32_t (&r)[2], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void sa_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void sa_mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                        const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    sa_mma(d, al, bh);
    sa_mma(d, ah, bl);
    sa_mma(d, ah, bh);
}

__global__ void __launch_bounds__(128)
sasa_mma_kernel(const float* __restrict__ qkv, int ld_qkv, const float* __restrict__ query_bbox, const float* __restrict__ tau, int ld_tau,
                const uint8_t* __restrict__ dn_mask, float x_lo, float x_hi, float y_lo, float y_hi,
                int B, int Q, int H, float* __restrict__ out) {
    __shared__ __align__(16) __nv_bfloat16 Qh[SA_BQ * SM_LD], Ql[SA_BQ * SM_LD];
    __shared__ __align__(16) __nv_bfloat16 Kh[SA_BK * SM_LD], Kl[SA_BK * SM_LD];
    __shared__ __align__(16) __nv_bfloat16 Vh[SA_BK * SM_LD], Vl[SA_BK * SM_LD];
    __shared__ float kcx[SA_BK], kcy[SA_BK];

    const int D = H * SA_HD;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g8 = lane >> 2, t4 = lane & 3;
    const int q0 = blockIdx.x * SA_BQ;
    const int h = blockIdx.y, b = blockIdx.z;
    const float scale = 0.17677669529663687f;
    const float* base = qkv + (long long)b * Q * ld_qkv;

    for (int i = tid; i < SA_BQ * 8; i += 128) {                 // Q tile, pre-scaled like nn.MultiheadAttention
        const int q = i >> 3, d4 = (i & 7) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + q < Q) v = ldg4(base + (long long)(q0 + q) * ld_qkv + h * SA_HD + d4);
        uint32_t h0, l0, h1, l1;
        sa_split2(v.x * scale, v.y * scale, h0, l0); sa_split2(v.z * scale, v.w * scale, h1, l1);
        *reinterpret_cast<uint2*>(Qh + q * SM_LD + d4) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(Ql + q * SM_LD + d4) = make_uint2(l0, l1);
    }
    // this thread's two query rows: centres and tau
    float rcx[2], rcy[2], rtau[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int gq = q0 + 16 * warp + g8 + 8 * r;
        const bool ok = gq < Q;
        rcx[r] = ok ? __fadd_rn(__fmul_rn(__ldg(query_bbox + ((long long)b * Q + gq) * 10), __fsub_rn(x_hi, x_lo)), x_lo) : 0.f;
        rcy[r] = ok ? __fadd_rn(__fmul_rn(__ldg(query_bbox + ((long long)b * Q + gq) * 10 + 1), __fsub_rn(y_hi, y_lo)), y_lo) : 0.f;
        rtau[r] = ok ? __ldg(tau + ((long long)b * Q + gq) * ld_tau + h) : 0.f;
    }
    __syncthreads();

    const int lm_r = lane & 7, lm_id = lane >> 3;
    const int a_row = lm_r + 8 * (lm_id & 1), a_col = 8 * (lm_id >> 1);
    const int b_row = lm_r, b_col = 8 * (lm_id & 1);
    uint32_t qh[2][4], ql[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        sa_ldsm_x4(qh[ks], Qh + (16 * warp + a_row) * SM_LD + 16 * ks + a_col);
        sa_ldsm_x4(ql[ks], Ql + (16 * warp + a_row) * SM_LD + 16 * ks + a_col);
    }

    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float oacc[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) { oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f; }

    for (int k0 = 0; k0 < Q; k0 += SA_BK) {
        __syncthreads();                                          // previous K/V tile fully consumed
        for (int i = tid; i < SA_BK * 8; i += 128) {
            const int k = i >> 3, d4 = (i & 7) * 4;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (k0 + k < Q) {
                const float* row = base + (long long)(k0 + k) * ld_qkv + h * SA_HD + d4;
                kv = ldg4(row + D); vv = ldg4(row + 2 * D);
            }
            uint32_t h0, l0, h1, l1;
            sa_split2(kv.x, kv.y, h0, l0); sa_split2(kv.z, kv.w, h1, l1);
            *reinterpret_cast<uint2*>(Kh + k * SM_LD + d4) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(Kl + k * SM_LD + d4) = make_uint2(l0, l1);
            sa_split2(vv.x, vv.y, h0, l0); sa_split2(vv.z, vv.w, h1, l1);
            *reinterpret_cast<uint2*>(Vh + k * SM_LD + d4) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(Vl + k * SM_LD + d4) = make_uint2(l0, l1);
        }
        if (tid < SA_BK) {
            const int gk = k0 + tid;
            const bool ok = gk < Q;
            kcx[tid] = ok ? __fadd_rn(__fmul_rn(__ldg(query_bbox + ((long long)b * Q + gk) * 10), __fsub_rn(x_hi, x_lo)), x_lo) : 0.f;
            kcy[tid] = ok ? __fadd_rn(__fmul_rn(__ldg(query_bbox + ((long long)b * Q + gk) * 10 + 1), __fsub_rn(y_hi, y_lo)), y_lo) : 0.f;
        }
        __syncthreads();

        // S = (q*scale) . k^T   (16 rows x 64 keys per warp)
        float sacc[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) { sacc[n][0] = sacc[n][1] = sacc[n][2] = sacc[n][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                uint32_t bh[2], bl[2];
                sa_ldsm_x2(bh, Kh + (8 * n + b_row) * SM_LD + 16 * ks + b_col);
                sa_ldsm_x2(bl, Kl + (8 * n + b_row) * SM_LD + 16 * ks + b_col);
                sa_mma3(sacc[n], qh[ks], ql[ks], bh, bl);
            }
        // + distance bias, masks, online softmax (row r of this thread = g8 + 8r; a row lives in one quad)
        float alpha[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float mx = -INFINITY;
#pragma unroll
            for (int n = 0; n < 8; ++n)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 8 * n + 2 * t4 + e;
                    const float dx = rcx[r] - kcx[k], dy = rcy[r] - kcy[k];
                    float v = sacc[n][2 * r + e] + (-sqrtf(dx * dx + dy * dy)) * rtau[r];
                    const int gq = q0 + 16 * warp + g8 + 8 * r;
                    if (dn_mask != nullptr && gq < Q && k0 + k < Q && dn_mask[(long long)gq * Q + k0 + k]) v = -INFINITY;
                    if (k0 + k >= Q) v = -INFINITY;
                    sacc[n][2 * r + e] = v;
                    mx = fmaxf(mx, v);
                }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float m_new = fmaxf(m_run[r], mx);
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
            alpha[r] = expf(m_run[r] - m_use);
            float rs = 0.f;
#pragma unroll
            for (int n = 0; n < 8; ++n)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float pv = expf(sacc[n][2 * r + e] - m_use);
                    sacc[n][2 * r + e] = pv;
                    rs += pv;
                }
            l_run[r] = l_run[r] * alpha[r] + rs;                 // per-thread partial; quad-reduced once at the end
            m_run[r] = m_new;
        }
#pragma unroll
        for (int n = 0; n < 4; ++n) { oacc[n][0] *= alpha[0]; oacc[n][1] *= alpha[0]; oacc[n][2] *= alpha[1]; oacc[n][3] *= alpha[1]; }
        // O += P . V : the C fragments of S tiles (2j, 2j+1) are exactly the A fragment of key block j
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t ph[4], pl[4];
            sa_split2(sacc[2 * j][0], sacc[2 * j][1], ph[0], pl[0]);
            sa_split2(sacc[2 * j][2], sacc[2 * j][3], ph[1], pl[1]);
            sa_split2(sacc[2 * j + 1][0], sacc[2 * j + 1][1], ph[2], pl[2]);
            sa_split2(sacc[2 * j + 1][2], sacc[2 * j + 1][3], ph[3], pl[3]);
#pragma unroll
            for (int nd = 0; nd < 4; ++nd) {
                uint32_t bh[2], bl[2];
                sa_ldsm_x2_trans(bh, Vh + (16 * j + (lane & 15)) * SM_LD + 8 * nd);
                sa_ldsm_x2_trans(bl, Vl + (16 * j + (lane & 15)) * SM_LD + 8 * nd);
                sa_mma3(oacc[nd], ph, pl, bh, bl);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float l = l_run[r];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        const int gq = q0 + 16 * warp + g8 + 8 * r;
        if (gq < Q) {
            const float inv = 1.f / l;
#pragma unroll
            for (int nd = 0; nd < 4; ++nd)
                *reinterpret_cast<float2*>(out + ((long long)b * Q + gq) * D + h * SA_HD + 8 * nd + 2 * t4) =
                    make_float2(oacc[nd][2 * r] * inv, oacc[nd][2 * r + 1] * inv);
        }
    }
}


// ------------------------------------------------------------------------------------------------
// v3 (default): every warp is a flash-attention worker on (16 queries, one head, a quarter of the keys); the two
// query tiles of a CTA that scan the same key quarter share one double-buffered cp.async pipeline that streams 32-key K/V tiles that were pre-split into bf16 (hi, lo)
// ONCE for the whole layer (sbev_split_bf16 on the in_proj output), straight into ldmatrix-ready shared memory --
// no block-level barriers, no per-CTA re-conversion.  The four partial (max, sum, O) results of a CTA are merged
// through shared memory at the end.  456 CTAs x 4 warps at Q = 900 (vs 120 x 4 before).
constexpr int S3_KT = 32;                                  // keys per tile
constexpr int S3_ARR = S3_KT * SM_LD;                      // bf16 elements of one [32][40] array
constexpr int S3_BUF_BYTES = 4 * S3_ARR * 2 + 2 * S3_KT * 4;   // Kh,Kl,Vh,Vl + key centres (x,y)

// MUFU-based square root / exponential for the attention logits: sqrt.approx (max relative error 2^-23) and
// ex2.approx on the log2(e)-scaled argument (relative error < 1e-6 for |x| < 100) -- well inside the 1e-4 parity bar, and
// a third of the instructions of the IEEE sqrtf / expf sequences that were ~25 % of this kernel's instruction stream.
__device__ __forceinline__ float sa_sqrt(float x) { float y; asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sa_exp(float x) { float y; asm("ex2.approx.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f)); return y; }
// v3 keeps its logits in the log2 domain (scale and tau pre-multiplied by log2(e): one FFMA per logit instead of FMUL + FFMA + FMUL) and
// uses the flush-to-zero forms: without .ftz ptxas wraps every MUFU in a denormal-range rescue (FSETP + 2 FMUL + FSEL per call -- the
// profile showed 10 FMUL per logit); a softmax weight below 1e-38 or a centre distance below 1e-19 m is zero for every purpose here
__device__ __forceinline__ float sa_sqrt_ftz(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sa_exp2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// KQ = key splits per CTA (4: 256 threads, two CTAs per SM -- the full-size layer; 8: 512 threads, half as many key tiles per
// warp -- when only a shard of the queries attends and the grid is far below one wave, latency per warp is what counts)
template <bool HAS_MASK, int KQ>
__global__ void __launch_bounds__(64 * KQ, KQ == 4 ? 2 : 1)
sasa_v3_kernel(const __nv_bfloat16* __restrict__ qkv_hi, const __nv_bfloat16* __restrict__ qkv_lo, int ld,
               const float* __restrict__ query_bbox, const float* __restrict__ tau, int ld_tau,
               const uint8_t* __restrict__ dn_mask, float x_lo, float x_hi, float y_lo, float y_hi,
               int B, int Q, int H, int qa, int qb, float* __restrict__ out) {
    // 2*KQ warps = KQ key splits x 2 query tiles; the two warps of a key-quarter share one K/V double buffer
    // (each fills half of it) and synchronise on their own named barrier (64 threads) -- never the whole CTA.
    extern __shared__ __align__(16) unsigned char s3_smem[];
    const int D = H * SA_HD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kq = warp >> 1, mt = warp & 1;
    const int g8 = lane >> 2, t4 = lane & 3;
    const int q0 = qa + blockIdx.x * 32 + 16 * mt;      // queries [qa, qb) of this launch; keys are always all Q
    const int h = blockIdx.y, b = blockIdx.z;
    const float scale2 = 0.17677669529663687f * 1.4426950408889634f;      // 1 / sqrt(32) * log2(e): logits live in the log2 domain
    unsigned char* mybuf = s3_smem + kq * 2 * S3_BUF_BYTES;
    const long long rowbase = (long long)b * Q;
    const int bar_id = 1 + kq;
    pdl_wait();
    pdl_trigger();

    // Q fragments straight from global (A operand: a0 (row g, k 2t..), a1 (row g+8), a2 (row g, k 2t+8..), a3 (row g+8, k+8))
    uint32_t qh[2][4], ql[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = q0 + g8 + 8 * (i & 1), k = 16 * ks + 2 * t4 + 8 * (i >> 1);
            uint32_t vh = 0, vl = 0;
            if (r < qb) {
                vh = __ldg(reinterpret_cast<const uint32_t*>(qkv_hi + (rowbase + r) * ld + h * SA_HD + k));
                vl = __ldg(reinterpret_cast<const uint32_t*>(qkv_lo + (rowbase + r) * ld + h * SA_HD + k));
            }
            qh[ks][i] = vh; ql[ks][i] = vl;
        }
    float rcx[2], rcy[2], rtau[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int gq = q0 + g8 + 8 * r;
        const bool ok = gq < qb;
        rcx[r] = ok ? __fadd_rn(__fmul_rn(__ldg(query_bbox + (rowbase + gq) * 10), __fsub_rn(x_hi, x_lo)), x_lo) : 0.f;
        rcy[r] = ok ? __fadd_rn(__fmul_rn(__ldg(query_bbox + (rowbase + gq) * 10 + 1), __fsub_rn(y_hi, y_lo)), y_lo) : 0.f;
        rtau[r] = ok ? -1.4426950408889634f * __ldg(tau + (rowbase + gq) * ld_tau + h) : 0.f;      // -tau * log2(e)
    }

    const int num_tiles = (Q + S3_KT - 1) / S3_KT;
    // this warp's half of a tile: mt == 0 -> K (hi, lo) + key centres, mt == 1 -> V (hi, lo); lane = key, zero-filled beyond Q
    auto issue = [&](int tile, int slot) {
        unsigned char* buf = mybuf + slot * S3_BUF_BYTES;
        const int key = tile * S3_KT + lane;
        const bool ok = key < Q;
        const long long src_row = (rowbase + (ok ? key : 0)) * ld + h * SA_HD + (mt ? 2 * D : D);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(buf) + (mt ? 2 * S3_ARR * 2 : 0) + lane * SM_LD * 2;
        const uint32_t nbytes = ok ? 16u : 0u;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const __nv_bfloat16* src = (a ? qkv_lo : qkv_hi) + src_row;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst + a * S3_ARR * 2 + c * 16), "l"(src + c * 8), "r"(nbytes) : "memory");
        }
        if (mt == 0) {
            float* kc = reinterpret_cast<float*>(buf + 4 * S3_ARR * 2);
            kc[lane] = ok ? __fadd_rn(__fmul_rn(__ldg(query_bbox + (rowbase + key) * 10), __fsub_rn(x_hi, x_lo)), x_lo) : 0.f;
            kc[S3_KT + lane] = ok ? __fadd_rn(__fmul_rn(__ldg(query_bbox + (rowbase + key) * 10 + 1), __fsub_rn(y_hi, y_lo)), y_lo) : 0.f;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int lm_r = lane & 7, lm_id = lane >> 3;
    const int b_row = lm_r, b_col = 8 * (lm_id & 1);
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float oacc[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) { oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f; }

    int slot = 0;
    if (kq < num_tiles) issue(kq, 0);
    for (int tile = kq; tile < num_tiles; tile += KQ, slot ^= 1) {
        if (tile + KQ < num_tiles) { issue(tile + KQ, slot ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");          // both halves of this tile have landed
        const unsigned char* buf = mybuf + slot * S3_BUF_BYTES;
        const __nv_bfloat16* Kh = reinterpret_cast<const __nv_bfloat16*>(buf);
        const __nv_bfloat16* Kl = Kh + S3_ARR;
        const __nv_bfloat16* Vh = Kl + S3_ARR;
        const __nv_bfloat16* Vl = Vh + S3_ARR;
        const float* kcx = reinterpret_cast<const float*>(buf + 4 * S3_ARR * 2);
        const float* kcy = kcx + S3_KT;
        const int k0 = tile * S3_KT;

        float sacc[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n) { sacc[n][0] = sacc[n][1] = sacc[n][2] = sacc[n][3] = 0.f; }
        {
            // all K fragments of the tile first, then the MMAs with the four key blocks interleaved: the three bf16x3 products of one
            // accumulator depend on each other (~35 cycles apiece), consecutive MMAs on DIFFERENT accumulators pipeline.  Every
            // accumulator still sees lo.hi, hi.lo, hi.hi per k-step in this order (bit-identical to sa_mma3)
            uint32_t kbh[2][4][2], kbl[2][4][2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    sa_ldsm_x2(kbh[ks][n], Kh + (8 * n + b_row) * SM_LD + 16 * ks + b_col);
                    sa_ldsm_x2(kbl[ks][n], Kl + (8 * n + b_row) * SM_LD + 16 * ks + b_col);
                }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                for (int n = 0; n < 4; ++n) sa_mma(sacc[n], ql[ks], kbh[ks][n]);
#pragma unroll
                for (int n = 0; n < 4; ++n) sa_mma(sacc[n], qh[ks], kbl[ks][n]);
#pragma unroll
                for (int n = 0; n < 4; ++n) sa_mma(sacc[n], qh[ks], kbh[ks][n]);
            }
        }
        float alpha[2];
        const bool ragged = k0 + S3_KT > Q;                       // (warp-uniform) only the last tile has keys beyond Q
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float mx = -INFINITY;
            const int gq = q0 + g8 + 8 * r;
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 8 * n + 2 * t4 + e;
                    const float dx = rcx[r] - kcx[k], dy = rcy[r] - kcy[k];
                    float v = fmaf(sa_sqrt_ftz(fmaf(dx, dx, dy * dy)), rtau[r], sacc[n][2 * r + e] * scale2);
                    if (HAS_MASK && gq < qb && k0 + k < Q && dn_mask[(long long)gq * Q + k0 + k]) v = -INFINITY;
                    if (ragged && k0 + k >= Q) v = -INFINITY;
                    sacc[n][2 * r + e] = v;
                    mx = fmaxf(mx, v);
                }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float m_new = fmaxf(m_run[r], mx);
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
            alpha[r] = sa_exp2_ftz(m_run[r] - m_use);
            float rs = 0.f;
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float pv = sa_exp2_ftz(sacc[n][2 * r + e] - m_use);
                    sacc[n][2 * r + e] = pv;
                    rs += pv;
                }
            l_run[r] = l_run[r] * alpha[r] + rs;
            m_run[r] = m_new;
        }
#pragma unroll
        for (int n = 0; n < 4; ++n) { oacc[n][0] *= alpha[0]; oacc[n][1] *= alpha[0]; oacc[n][2] *= alpha[1]; oacc[n][3] *= alpha[1]; }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            uint32_t ph[4], pl[4];
            sa_split2(sacc[2 * j][0], sacc[2 * j][1], ph[0], pl[0]);
            sa_split2(sacc[2 * j][2], sacc[2 * j][3], ph[1], pl[1]);
            sa_split2(sacc[2 * j + 1][0], sacc[2 * j + 1][1], ph[2], pl[2]);
            sa_split2(sacc[2 * j + 1][2], sacc[2 * j + 1][3], ph[3], pl[3]);
            uint32_t vbh[4][2], vbl[4][2];
#pragma unroll
            for (int nd = 0; nd < 4; ++nd) {
                sa_ldsm_x2_trans(vbh[nd], Vh + (16 * j + (lane & 15)) * SM_LD + 8 * nd);
                sa_ldsm_x2_trans(vbl[nd], Vl + (16 * j + (lane & 15)) * SM_LD + 8 * nd);
            }
#pragma unroll
            for (int nd = 0; nd < 4; ++nd) sa_mma(oacc[nd], pl, vbh[nd]);
#pragma unroll
            for (int nd = 0; nd < 4; ++nd) sa_mma(oacc[nd], ph, vbl[nd]);
#pragma unroll
            for (int nd = 0; nd < 4; ++nd) sa_mma(oacc[nd], ph, vbh[nd]);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");          // both warps done with this slot before it is refilled
    }

    // ---- merge the four key-quarter partials of each query tile
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    __syncthreads();                                              // every warp is done with its pipeline buffers
    float* mo = reinterpret_cast<float*>(s3_smem);                // [2 mt][KQ][16 rows][32 + 2]  (O row, m, l)
    constexpr int MLD = SA_HD + 2;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float* row = mo + ((mt * KQ + kq) * 16 + g8 + 8 * r) * MLD;
#pragma unroll
        for (int nd = 0; nd < 4; ++nd) { row[8 * nd + 2 * t4] = oacc[nd][2 * r]; row[8 * nd + 2 * t4 + 1] = oacc[nd][2 * r + 1]; }
        if (t4 == 0) { row[SA_HD] = m_run[r]; row[SA_HD + 1] = l_run[r]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * SA_HD; i += 64 * KQ) {
        const int r32 = i >> 5, d = i & 31;
        const int m2 = r32 >> 4, r = r32 & 15;
        float mmax = -INFINITY;
#pragma unroll
        for (int w = 0; w < KQ; ++w) mmax = fmaxf(mmax, mo[((m2 * KQ + w) * 16 + r) * MLD + SA_HD]);
        const float muse = (mmax == -INFINITY) ? 0.f : mmax;
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int w = 0; w < KQ; ++w) {
            const float* row = mo + ((m2 * KQ + w) * 16 + r) * MLD;
            const float f = sa_exp2_ftz(row[SA_HD] - muse);
            num += f * row[d];
            den += f * row[SA_HD + 1];
        }
        const int gq = qa + blockIdx.x * 32 + r32;
        if (gq < qb) out[(rowbase + gq) * D + h * SA_HD + d] = num / den;
    }
}

}  // namespace sbev

using namespace sbev;

extern "C" int sbev_sasa_fwd(const float* qkv, int ld_qkv, const float* query_bbox, const float* tau, int ld_tau,
                             const uint8_t* dn_mask, const float* pc_range, int B, int Q, int H, int D, float* out, void* stream) {
    SBEV_REQUIRE(qkv && query_bbox && tau && pc_range && out, SBEV_ERR_INVALID, "sbev_sasa_fwd: null pointer");
    SBEV_REQUIRE(B >= 0 && Q >= 0 && H > 0, SBEV_ERR_INVALID, "sbev_sasa_fwd: bad sizes");
    SBEV_REQUIRE(ld_qkv >= 3 * D && ld_tau >= H, SBEV_ERR_INVALID, "sbev_sasa_fwd: row strides too small");
    SBEV_REQUIRE(D == H * SA_HD, SBEV_ERR_UNSUPPORTED, "sbev_sasa_fwd: head dim must be 32 (D=%d, H=%d)", D, H);
    if (B == 0 || Q == 0) return SBEV_OK;
    dim3 grid((Q + SA_BQ - 1) / SA_BQ, H, B);
    const int impl = get_option(OPT_SASA_IMPL);  // 1 selects the fp32 FFMA kernel
    SBEV_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (ld_qkv & 3) == 0, SBEV_ERR_INVALID, "sbev_sasa_fwd: qkv must be 16-byte aligned with ld_qkv % 4 == 0");
    if (impl == 0)
        sasa_mma_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(qkv, ld_qkv, query_bbox, tau, ld_tau, dn_mask, pc_range[0], pc_range[3], pc_range[1], pc_range[4], B, Q, H, out);
    else
        sasa_hd32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(qkv, ld_qkv, query_bbox, tau, ld_tau, dn_mask, pc_range[0], pc_range[3], pc_range[1], pc_range[4], B, Q, H, out);
    return check_launch("sbev_sasa_fwd");
}

extern "C" int sbev_sasa_split_range_fwd(const uint16_t* qkv_hi, const uint16_t* qkv_lo, int ld, const float* query_bbox,
                                         const float* tau, int ld_tau, const uint8_t* dn_mask, const float* pc_range,
                                         int B, int Q, int H, int D, int q_begin, int q_end, float* out, void* stream) {
    SBEV_REQUIRE(q_begin >= 0 && q_begin <= q_end && q_end <= Q, SBEV_ERR_INVALID, "sbev_sasa_split_range_fwd: query range [%d,%d) outside [0,%d]", q_begin, q_end, Q);
    SBEV_REQUIRE(qkv_hi && qkv_lo && query_bbox && tau && pc_range && out, SBEV_ERR_INVALID, "sbev_sasa_split_fwd: null pointer");
    SBEV_REQUIRE(B >= 0 && Q >= 0 && H > 0 && ld >= 3 * D && ld_tau >= H, SBEV_ERR_INVALID, "sbev_sasa_split_fwd: bad sizes");
    SBEV_REQUIRE(D == H * SA_HD, SBEV_ERR_UNSUPPORTED, "sbev_sasa_split_fwd: head dim must be 32 (D=%d, H=%d)", D, H);
    SBEV_REQUIRE((ld & 7) == 0 && (reinterpret_cast<uintptr_t>(qkv_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(qkv_lo) & 15) == 0,
                 SBEV_ERR_INVALID, "sbev_sasa_split_fwd: bf16 operands must be 16-byte aligned with ld % 8 == 0");
    if (B == 0 || q_end == q_begin) return SBEV_OK;
    // key splits per CTA: 8 (512 threads) when the grid would not fill the GPU with 4 (a query shard), option "sasa_kq" overrides
    int kq = get_option(OPT_SASA_KQ);
    if (kq != 4 && kq != 8) kq = ((long long)((q_end - q_begin + 31) / 32) * H * B <= device_num_sms()) ? 8 : 4;
    const size_t smem = (size_t)kq * 2 * S3_BUF_BYTES;
    SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(sasa_v3_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * 2 * S3_BUF_BYTES));
        cudaFuncSetAttribute(sasa_v3_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * 2 * S3_BUF_BYTES));
        cudaFuncSetAttribute(sasa_v3_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * 2 * S3_BUF_BYTES));
        cudaFuncSetAttribute(sasa_v3_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * 2 * S3_BUF_BYTES)));
    dim3 grid((q_end - q_begin + 31) / 32, H, B);
#define SBEV_SASA_LAUNCH(MASK, KQV)                                                                                                      \
    launch_pdl(sasa_v3_kernel<MASK, KQV>, grid, dim3(64 * KQV), smem, (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(qkv_hi), \
               reinterpret_cast<const __nv_bfloat16*>(qkv_lo), ld, query_bbox, tau, ld_tau, dn_mask, pc_range[0], pc_range[3], pc_range[1],  \
               pc_range[4], B, Q, H, q_begin, q_end, out)
    if (kq == 8) { if (dn_mask != nullptr) SBEV_SASA_LAUNCH(true, 8); else SBEV_SASA_LAUNCH(false, 8); return check_launch("sbev_sasa_split_fwd"); }
    if (dn_mask != nullptr)
        launch_pdl(sasa_v3_kernel<true, 4>, grid, dim3(256), smem, (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(qkv_hi), reinterpret_cast<const __nv_bfloat16*>(qkv_lo), ld,
                   query_bbox, tau, ld_tau, dn_mask, pc_range[0], pc_range[3], pc_range[1], pc_range[4], B, Q, H, q_begin, q_end, out);
    else
        launch_pdl(sasa_v3_kernel<false, 4>, grid, dim3(256), smem, (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(qkv_hi), reinterpret_cast<const __nv_bfloat16*>(qkv_lo), ld,
                   query_bbox, tau, ld_tau, dn_mask, pc_range[0], pc_range[3], pc_range[1], pc_range[4], B, Q, H, q_begin, q_end, out);
    return check_launch("sbev_sasa_split_fwd");
}

extern "C" int sbev_sasa_split_fwd(const uint16_t* qkv_hi, const uint16_t* qkv_lo, int ld, const float* query_bbox,
                                   const float* tau, int ld_tau, const uint8_t* dn_mask, const float* pc_range,
                                   int B, int Q, int H, int D, float* out, void* stream) {
    return sbev_sasa_split_range_fwd(qkv_hi, qkv_lo, ld, query_bbox, tau, ld_tau, dn_mask, pc_range, B, Q, H, D, 0, Q, out, stream);
}
