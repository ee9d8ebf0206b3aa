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
#include <math.h>

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
    sasa_hd32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(qkv, ld_qkv, query_bbox, tau, ld_tau, dn_mask, pc_range[0], pc_range[3], pc_range[1], pc_range[4], B, Q, H, out);
    return check_launch("sbev_sasa_fwd");
}
