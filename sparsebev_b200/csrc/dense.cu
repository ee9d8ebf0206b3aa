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
#include <mutex>

namespace sbev {

constexpr int DENSE_ROWS = 8;
constexpr int CHAIN_STAGES = 3;
constexpr int CHAIN_STAGE_FLOATS = 8192;          // 32 KB weight chunk per pipeline stage
constexpr int CHAIN_MAX_LAYERS = 6;
constexpr int CHAIN_MAX_PASS = 4;                 // N <= 4 * 256

struct ChainLayer {
    const float* Wt; const float* bias; const float* ln_w; const float* ln_b; const float* residual; float* y;
    __nv_bfloat16* y_hi; __nv_bfloat16* y_lo;     // optional bf16 (hi, lo) split of the stored output, row stride ldy
    int ldw, K, N, flags, ldy, kc;                // kc = weight rows per chunk
    const uint8_t* wpack;                         // optional pre-tiled (hi | lo) weight stream: 32 KB per (128-feature block, 64-k chunk), see the header
};
struct ChainParams {
    const float* x; int ldx, M, n_layers, act_ld; // act_ld: row stride (floats) of the shared activation buffers
    const float* aux_proposal; const float* aux_time_diff; int aux_Q, aux_T;   // SBEV_DENSE_REFINE epilogue
    // optional input stage (tensor-core kernel only): the chain's input rows are
    //   LN(sum_z in_partial[z][row] + in_bias + in_res[row])   (split-K partials of the preceding GEMM, [in_nsplit][M][K0])
    // computed in the prologue and also stored to in_out [M][K0] (x / ldx are then unused)
    const float* in_partial; int in_nsplit; const float* in_bias; const float* in_res; const float* in_ln_w; const float* in_ln_b; float* in_out;
    // optional sample-points epilogue on the last layer (CHAIN_FLAG_POINTS; tensor-core kernel only): its output row
    // [.. | GP*3 offsets at sp_off_col | GP*L scale logits at sp_log_col | ..] -> sp_points [M][GP][3], sp_scale_w [M][GP][L]
    const float* sp_bbox; float* sp_points; float* sp_scale_w; float sp_r[6]; int sp_GP, sp_L, sp_off_col, sp_log_col; float sp_ssign;
    ChainLayer layer[CHAIN_MAX_LAYERS];
};
constexpr int CHAIN_FLAG_POINTS = 1 << 9;

// One sample point: box decode + offset scaling + yaw rotation + softmax over the L scale logits
// (sparsebev_sampling.py:8-24, bbox/utils.py:63-77, models/utils.py:49-84, sparsebev_transformer.py:298-299).
// Shared by sample_points_kernel and the chain kernel's fused epilogue, so both produce identical bits.
__device__ __forceinline__ void sample_point_one(const float* bb, const float* op, const float* lp, int L,
                                                 float r0, float r1, float r2, float r3, float r4, float r5,
                                                 float* pt, float* w, float ssign = 1.f) {
    // decode_bbox: xyz*(hi-lo)+lo, exp(log-size), atan2(sin, cos)   (separate mul and add as in torch)
    const float cx = __fadd_rn(__fmul_rn(bb[0], __fsub_rn(r3, r0)), r0);
    const float cy = __fadd_rn(__fmul_rn(bb[1], __fsub_rn(r4, r1)), r1);
    const float cz = __fadd_rn(__fmul_rn(bb[2], __fsub_rn(r5, r2)), r2);
    const float sw = expf(bb[3]), sl = expf(bb[4]), sh = expf(bb[5]);
    const float yaw = atan2f(bb[6], bb[7]);
    // ssign = -1: the legacy 'v0.17.1' checkpoint convention rotates the other way (models/utils.py:66-71)
    const float s = ssign * sinf(yaw), c = cosf(yaw);
    const float dx = __fmul_rn(sw, op[0]), dy = __fmul_rn(sl, op[1]), dz = __fmul_rn(sh, op[2]);
    // rotate counter-clockwise by yaw about z: x' = x*c + y*(-s), y' = x*s + y*c
    const float rx = __fadd_rn(__fmul_rn(dx, c), __fmul_rn(dy, -s));
    const float ry = __fadd_rn(__fmul_rn(dx, s), __fmul_rn(dy, c));
    pt[0] = __fadd_rn(cx, rx);
    pt[1] = __fadd_rn(cy, ry);
    pt[2] = __fadd_rn(cz, dz);
    // softmax over L
    float mx = -INFINITY;
    for (int l = 0; l < L; ++l) mx = fmaxf(mx, lp[l]);
    float e[SBEV_MAX_LEVELS], sum = 0.f;
    for (int l = 0; l < L; ++l) { e[l] = expf(lp[l] - mx); sum += e[l]; }
    for (int l = 0; l < L; ++l) w[l] = __fdiv_rn(e[l], sum);
}

__device__ __forceinline__ uint32_t dsmem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void chain_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(dsmem_u32(bar)), "r"(parity) : "memory");
        if (!done && spins > (1u << 26)) __trap();
    }
}

// Row epilogue shared by the FFMA and the tensor-core chain kernels: one warp finishes one output row held in
// shared memory (yr[0..N)): + bias -> (+ residual) -> LayerNorm -> ReLU -> (+ residual) -> (refine) -> yr and global.
// sv / sres (optional): shared-memory copies of the layer's bias | ln_w | ln_b (3 x CHAIN_VEC_LD floats) and of this row's
// residual, prefetched with cp.async while the layer's GEMM ran -- the epilogue then waits on no global load.
constexpr int CHAIN_VEC_LD = 1024;
// U = columns per lane per batch (their global operands are loaded together): the unrolled batch is U instructions long whatever N
// is, so narrow rows (the 10-wide cls / reg outputs) take the U = 1 instance -- see chain_row_epilogue below.
template <int U>
__device__ __forceinline__ void chain_row_epilogue_u(const ChainParams& prm, const ChainLayer& L, int row, float* yr, int lane,
                                                     const float* sv, const float* sres) {
    const int N = L.N;
    const bool live = row < prm.M;
    const bool pre_res = (L.flags & SBEV_DENSE_RES_PRE_LN) && L.residual != nullptr;
    const bool has_res = live && L.residual != nullptr;
    const float* res_row = L.residual ? L.residual + (long long)row * N : nullptr;
    float s = 0.f;
    for (int n0 = 0; n0 < N; n0 += 32 * U) {
        float bv[U], rv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int n = n0 + lane + 32 * u;
            bv[u] = (L.bias && n < N) ? (sv ? sv[n] : __ldg(L.bias + n)) : 0.f;
            rv[u] = (pre_res && has_res && n < N) ? (sres ? sres[n] : __ldg(res_row + n)) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int n = n0 + lane + 32 * u;
            if (n < N) { const float v = (yr[n] + bv[u]) + rv[u]; yr[n] = v; s += v; }
        }
    }
    float mean = 0.f, rstd = 1.f;
    if (L.ln_w != nullptr) {
        mean = warp_sum(s) / (float)N;
        float ss = 0.f;
        for (int n = lane; n < N; n += 32) { const float d = yr[n] - mean; ss += d * d; }
        rstd = rsqrtf(warp_sum(ss) / (float)N + 1e-5f);
    }
    const bool refine = (L.flags & SBEV_DENSE_REFINE) && live;
    for (int n0 = 0; n0 < N; n0 += 32 * U) {
        float gv[U], ov[U], rv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int n = n0 + lane + 32 * u;
            const bool ok = n < N;
            gv[u] = (L.ln_w && ok) ? (sv ? sv[CHAIN_VEC_LD + n] : __ldg(L.ln_w + n)) : 1.f;
            ov[u] = (L.ln_w && ok) ? (sv ? sv[2 * CHAIN_VEC_LD + n] : __ldg(L.ln_b + n)) : 0.f;
            rv[u] = (!pre_res && has_res && ok) ? (sres ? sres[n] : __ldg(res_row + n)) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int n = n0 + lane + 32 * u;
            if (n < N) {
                float v = yr[n];
                if (L.ln_w != nullptr) v = (v - mean) * rstd * gv[u] + ov[u];
                if (L.flags & SBEV_DENSE_RELU) v = fmaxf(v, 0.f);
                v += rv[u];
                if (refine) {
                    // refine_bbox + velocity rescale (sparsebev_transformer.py:155-160,179-183)
                    if (n < 3) {
                        const float x = fminf(fmaxf(__ldg(prm.aux_proposal + (long long)row * N + n), 0.f), 1.f);
                        v = v + logf(__fdiv_rn(fmaxf(x, 1e-5f), fmaxf(1.f - x, 1e-5f)));
                        v = __fdiv_rn(1.f, 1.f + expf(-v));
                    } else if (n >= 8 && prm.aux_T > 1) {
                        float td = __ldg(prm.aux_time_diff + (row / prm.aux_Q) * prm.aux_T + 1);
                        if (td < 1e-5f) td = 1.0f;
                        v = __fdiv_rn(v, td);
                    }
                }
                yr[n] = v;
                if (live) {
                    if (L.y != nullptr) L.y[(long long)row * L.ldy + n] = v;
                    if (L.y_hi != nullptr) {
                        const __nv_bfloat16 h = __float2bfloat16_rn(v);
                        L.y_hi[(long long)row * L.ldy + n] = h;
                        L.y_lo[(long long)row * L.ldy + n] = __float2bfloat16_rn(v - __bfloat162float(h));
                    }
                }
            }
        }
    }
}

__device__ __forceinline__ void chain_row_epilogue(const ChainParams& prm, const ChainLayer& L, int row, float* yr, int lane,
                                                   const float* sv = nullptr, const float* sres = nullptr) {
    if (L.N <= 32) chain_row_epilogue_u<1>(prm, L, row, yr, lane, sv, sres);      // (warp-uniform)
    else chain_row_epilogue_u<8>(prm, L, row, yr, lane, sv, sres);
}

__device__ __forceinline__ void mc_split2(float a, float b, uint32_t& hi, uint32_t& lo);
// Vectorised form of chain_row_epilogue for the tensor-core chains (profile: the scalar epilogue was ~45 % of the chain
// kernels' warp-stall samples -- the kernels are bound by the length of each warp's instruction stream, not by weight
// streaming).  Same arithmetic; a lane owns the 16-byte column groups lane, lane + 32, ...; the row stays in registers
// between the three passes, every global / shared access is 16 bytes (8 for the bf16 outputs), and the next layer's
// bf16 (hi, lo) operand row (nh / nl, zero padded up to Kn) is produced in the same pass.  Requires N % 4 == 0,
// N <= 1024, ldy % 4 == 0, 16-byte aligned operands and no refine epilogue (host sets CHAIN_FLAG_VEC4 when that holds).
// PER = 16-byte column groups per lane the three passes are unrolled for (N <= 128 PER): the loops used to be unrolled for the
// maximum (8 groups, N = 1024) with the dead groups predicated off -- 462 issued instructions per warp and layer at N = 256, where
// two groups are live (ncu source view of the cls chain: the epilogue was 36 % of all executed instructions and 45 % of the stall
// samples).  The dispatcher below picks 2 / 4 / 8.
constexpr int CHAIN_FLAG_VEC4 = 1 << 8;
template <int PER>
__device__ __forceinline__ void chain_row_epilogue_v4_p(const ChainParams& prm, const ChainLayer& L, int row, const float* yr, int lane,
                                                        const float* sv, const float* sres, __nv_bfloat16* nh, __nv_bfloat16* nl, int Kn) {
    const int N = L.N, ng = N >> 2;
    const bool live = row < prm.M;
    const bool pre_res = (L.flags & SBEV_DENSE_RES_PRE_LN) && L.residual != nullptr;
    const bool has_res = live && L.residual != nullptr;
    const float* res_row = L.residual ? L.residual + (long long)row * N : nullptr;
    float4 v[PER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int g = lane + 32 * i;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < ng) {
            a = *reinterpret_cast<const float4*>(yr + 4 * g);
            if (L.bias != nullptr) {
                const float4 b = sv ? *reinterpret_cast<const float4*>(sv + 4 * g) : ldg4(L.bias + 4 * g);
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            if (pre_res && has_res) {
                const float4 r = sres ? *reinterpret_cast<const float4*>(sres + 4 * g) : ldg4(res_row + 4 * g);
                a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
            }
            s += (a.x + a.y) + (a.z + a.w);
        }
        v[i] = a;
    }
    float mean = 0.f, rstd = 1.f;
    const bool has_ln = L.ln_w != nullptr;
    if (has_ln) {
        mean = warp_sum(s) / (float)N;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i)
            if (lane + 32 * i < ng) {
                const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
                ss += (a * a + b * b) + (c * c + d * d);
            }
        rstd = rsqrtf(warp_sum(ss) / (float)N + 1e-5f);
    }
    const long long yoff = (long long)row * L.ldy;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int g = lane + 32 * i;
        if (g < ng) {
            float4 o = v[i];
            if (has_ln) {
                const float4 gm = sv ? *reinterpret_cast<const float4*>(sv + CHAIN_VEC_LD + 4 * g) : ldg4(L.ln_w + 4 * g);
                const float4 bt = sv ? *reinterpret_cast<const float4*>(sv + 2 * CHAIN_VEC_LD + 4 * g) : ldg4(L.ln_b + 4 * g);
                o.x = (o.x - mean) * rstd * gm.x + bt.x; o.y = (o.y - mean) * rstd * gm.y + bt.y;
                o.z = (o.z - mean) * rstd * gm.z + bt.z; o.w = (o.w - mean) * rstd * gm.w + bt.w;
            }
            if (L.flags & SBEV_DENSE_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (!pre_res && has_res) {
                const float4 r = sres ? *reinterpret_cast<const float4*>(sres + 4 * g) : ldg4(res_row + 4 * g);
                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
            }
            uint32_t h0, l0, h1, l1;
            mc_split2(o.x, o.y, h0, l0); mc_split2(o.z, o.w, h1, l1);
            if (live) {
                if (L.y != nullptr) *reinterpret_cast<float4*>(L.y + yoff + 4 * g) = o;
                if (L.y_hi != nullptr) {
                    *reinterpret_cast<uint2*>(L.y_hi + yoff + 4 * g) = make_uint2(h0, h1);
                    *reinterpret_cast<uint2*>(L.y_lo + yoff + 4 * g) = make_uint2(l0, l1);
                }
            }
            if (nh != nullptr) {
                *reinterpret_cast<uint2*>(nh + 4 * g) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(nl + 4 * g) = make_uint2(l0, l1);
            }
        }
    }
    if (nh != nullptr)
        for (int k = N + 4 * lane; k < Kn; k += 128) {
            *reinterpret_cast<uint2*>(nh + k) = make_uint2(0u, 0u);
            *reinterpret_cast<uint2*>(nl + k) = make_uint2(0u, 0u);
        }
}
__device__ __forceinline__ void chain_row_epilogue_v4(const ChainParams& prm, const ChainLayer& L, int row, const float* yr, int lane,
                                                      const float* sv, const float* sres, __nv_bfloat16* nh, __nv_bfloat16* nl, int Kn) {
    const int per = ((L.N >> 2) + 31) >> 5;                   // (warp-uniform)
    if (per <= 2) chain_row_epilogue_v4_p<2>(prm, L, row, yr, lane, sv, sres, nh, nl, Kn);
    else if (per <= 4) chain_row_epilogue_v4_p<4>(prm, L, row, yr, lane, sv, sres, nh, nl, Kn);
    else chain_row_epilogue_v4_p<8>(prm, L, row, yr, lane, sv, sres, nh, nl, Kn);
}

// A chain of up to 6 Linear(+bias)(+residual)(+LayerNorm)(+ReLU) layers (fp32 FFMA version; exact fp32).  One CTA owns 8 full rows from the
// first layer to the last (activations never leave shared memory, LayerNorm never leaves the CTA); the
// pre-transposed weights of ALL layers are streamed back-to-back through a 3-stage ring of 32 KB chunks by
// 1-D bulk copies (cp.async.bulk + mbarrier complete_tx), so the next layer's weights are already in flight
// while the current layer's epilogue runs.  fp32 FFMA, 2 rows x (up to 4 x 4) columns per thread.
__global__ void __launch_bounds__(256, 2)
dense_chain_kernel(const __grid_constant__ ChainParams prm) {
    extern __shared__ __align__(128) float smem[];
    float* wbuf = smem;                                              // [STAGES][8192]
    float* act0 = smem + CHAIN_STAGES * CHAIN_STAGE_FLOATS;           // [8][act_ld]
    float* act1 = act0 + DENSE_ROWS * prm.act_ld;
    __shared__ uint64_t full_bar[CHAIN_STAGES], empty_bar[CHAIN_STAGES];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * DENSE_ROWS;
    if (tid == 0) {
        for (int s = 0; s < CHAIN_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dsmem_u32(&full_bar[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dsmem_u32(&empty_bar[s])), "r"(8));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // stage the input rows
    {
        const int K0 = prm.layer[0].K;
        for (int i = tid; i < DENSE_ROWS * prm.act_ld; i += 256) {
            const int r = i / prm.act_ld, k = i - r * prm.act_ld;
            act0[i] = (row0 + r < prm.M && k < K0) ? __ldg(prm.x + (long long)(row0 + r) * prm.ldx + k) : 0.f;
        }
    }
    __syncthreads();

    // producer state (thread 0): walks the global chunk sequence over all layers
    int p_layer = 0, p_k0 = 0, p_idx = 0;
    auto produce_one = [&]() {
        if (p_layer >= prm.n_layers) return;
        const ChainLayer& L = prm.layer[p_layer];
        const int kc = min(L.kc, L.K - p_k0);
        const int stage = p_idx % CHAIN_STAGES;
        if (p_idx >= CHAIN_STAGES) chain_mbar_wait(&empty_bar[stage], ((p_idx / CHAIN_STAGES) - 1) & 1);
        const uint32_t bytes = (uint32_t)kc * L.ldw * 4u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dsmem_u32(&full_bar[stage])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dsmem_u32(wbuf + stage * CHAIN_STAGE_FLOATS)), "l"(L.Wt + (long long)p_k0 * L.ldw), "r"(bytes),
                       "r"(dsmem_u32(&full_bar[stage])) : "memory");
        p_k0 += kc; ++p_idx;
        if (p_k0 >= L.K) { p_k0 = 0; ++p_layer; }
    };
    if (tid == 0) { for (int s = 0; s < CHAIN_STAGES - 1; ++s) produce_one(); }

    const int cg0 = tid & 63, rg = tid >> 6;                          // rows 2rg, 2rg+1; column groups cg0 + 64*pass
    int c_idx = 0;                                                    // consumer chunk counter
    float* xin = act0;
    float* yout = act1;
    for (int li = 0; li < prm.n_layers; ++li) {
        const ChainLayer& L = prm.layer[li];
        const int ncg = L.ldw >> 2;
        float4 acc[CHAIN_MAX_PASS][2];
#pragma unroll
        for (int ps = 0; ps < CHAIN_MAX_PASS; ++ps) { acc[ps][0] = make_float4(0.f, 0.f, 0.f, 0.f); acc[ps][1] = acc[ps][0]; }
        const float* x0 = xin + (2 * rg) * prm.act_ld;
        const float* x1 = x0 + prm.act_ld;
        for (int k0 = 0; k0 < L.K; k0 += L.kc, ++c_idx) {
            if (tid == 0) produce_one();
            const int stage = c_idx % CHAIN_STAGES;
            chain_mbar_wait(&full_bar[stage], (c_idx / CHAIN_STAGES) & 1);
            const float* wch = wbuf + stage * CHAIN_STAGE_FLOATS;
            const int kc = min(L.kc, L.K - k0);
#define SBEV_FMA4(a, xv, wv) a.x = fmaf(xv, wv.x, a.x); a.y = fmaf(xv, wv.y, a.y); a.z = fmaf(xv, wv.z, a.z); a.w = fmaf(xv, wv.w, a.w);
            int kk = 0;
            for (; kk + 4 <= kc; kk += 4) {
                const float4 xa = *reinterpret_cast<const float4*>(x0 + k0 + kk);
                const float4 xb = *reinterpret_cast<const float4*>(x1 + k0 + kk);
#pragma unroll
                for (int ps = 0; ps < CHAIN_MAX_PASS; ++ps) {
                    const int cg = cg0 + 64 * ps;
                    if (cg < ncg) {
                        const float* wp = wch + kk * L.ldw + 4 * cg;
                        const float4 w0 = *reinterpret_cast<const float4*>(wp);
                        const float4 w1 = *reinterpret_cast<const float4*>(wp + L.ldw);
                        const float4 w2 = *reinterpret_cast<const float4*>(wp + 2 * L.ldw);
                        const float4 w3 = *reinterpret_cast<const float4*>(wp + 3 * L.ldw);
                        SBEV_FMA4(acc[ps][0], xa.x, w0) SBEV_FMA4(acc[ps][0], xa.y, w1) SBEV_FMA4(acc[ps][0], xa.z, w2) SBEV_FMA4(acc[ps][0], xa.w, w3)
                        SBEV_FMA4(acc[ps][1], xb.x, w0) SBEV_FMA4(acc[ps][1], xb.y, w1) SBEV_FMA4(acc[ps][1], xb.z, w2) SBEV_FMA4(acc[ps][1], xb.w, w3)
                    }
                }
            }
            for (; kk < kc; ++kk) {
                const float xa = x0[k0 + kk], xb = x1[k0 + kk];
#pragma unroll
                for (int ps = 0; ps < CHAIN_MAX_PASS; ++ps) {
                    const int cg = cg0 + 64 * ps;
                    if (cg < ncg) {
                        const float4 w0 = *reinterpret_cast<const float4*>(wch + kk * L.ldw + 4 * cg);
                        SBEV_FMA4(acc[ps][0], xa, w0) SBEV_FMA4(acc[ps][1], xb, w0)
                    }
                }
            }
#undef SBEV_FMA4
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dsmem_u32(&empty_bar[stage])) : "memory");
        }
#pragma unroll
        for (int ps = 0; ps < CHAIN_MAX_PASS; ++ps) {
            const int cg = cg0 + 64 * ps;
            if (cg < ncg) {
                *reinterpret_cast<float4*>(yout + (2 * rg) * prm.act_ld + 4 * cg) = acc[ps][0];
                *reinterpret_cast<float4*>(yout + (2 * rg + 1) * prm.act_ld + 4 * cg) = acc[ps][1];
            }
        }
        __syncthreads();
        // epilogue: one warp per row; result stays in yout (next layer's input) and optionally goes to global
        chain_row_epilogue(prm, L, row0 + warp, yout + warp * prm.act_ld, lane);
        if (li + 1 < prm.n_layers) {   // zero the K-padding of the next layer's input (its K may not be a multiple of 4)
            float* yr = yout + warp * prm.act_ld;
            for (int n = L.N + lane; n < ((L.N + 3) & ~3); n += 32) yr[n] = 0.f;
        }
        __syncthreads();
        float* t = xin; xin = yout; yout = t;
    }
}


// ------------------------------------------------------------------------------------------------
// Tensor-core chain (default).  Same contract as dense_chain_kernel, but the matmuls run on
// mma.sync.m16n8k16 with the bf16x3 split and fp32 accumulation.  The problem is transposed so that the 8 rows a
// CTA owns sit on the MMA N dimension: y^T[n][row] = sum_k W[n][k] x[row][k] -- A = W (the nn.Linear layout, no
// transpose needed), B = x.  Pre-split bf16 (hi, lo) weights [Npad][Kpad] are streamed as 128-row x 64-k tiles by
// 2-D TMA (128-byte swizzle, so ldmatrix is conflict free) through a 4-stage mbarrier ring fed by a dedicated
// producer warp; warp w of the 8 consumer warps owns output features 16w..16w+15 of the current 128-wide block.
constexpr int MC_STAGES = 4;
constexpr int MC_TILE_BYTES = 128 * 64 * 2;            // one (hi or lo) 128 x 64 bf16 tile
constexpr int MC_XLD = 512 + 8;                        // bf16 row stride of the activation operand (K <= 512)
constexpr int MC_YLD = 1024 + 4;                       // fp32 row stride of the layer output (N <= 1024)
constexpr int MC_RES_LD = 256;                         // residual rows are staged in shared memory when N <= 256

struct ChainMaps { CUtensorMap hi[CHAIN_MAX_LAYERS]; CUtensorMap lo[CHAIN_MAX_LAYERS]; };

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void mc_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void mc_ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mc_ldsm_x2(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mc_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// CL = cluster size (1 = no cluster).  With CL > 1 the CL CTAs of a cluster own CL different row groups but need the
// SAME weight tiles: CTA r fetches rows [r*128/CL, (r+1)*128/CL) of each tile and TMA-multicasts them into the
// shared memory of all CL CTAs, so every weight byte leaves L2 once per cluster instead of once per CTA.  A stage is
// refilled only after the consumers of ALL CL CTAs released it (each consumer warp arrives on every CTA's barrier).
// R = rows per CTA (8, or 16 = two n8 MMA tiles per warp: half as many CTAs stream the weights -- for chains that run side by side
// with another chain on a second stream, cls || reg, so that both fit on the 148 SMs at once); XLD / YLD = row strides of the
// activation operand / the layer output (the 16-row form is instantiated for K, N <= 256 only: its buffers must fit shared memory).
template <int CL, int R = DENSE_ROWS, int XLD = MC_XLD, int YLD = MC_YLD>
__global__ void __launch_bounds__(288, 1)
dense_chain_mma_kernel(const __grid_constant__ ChainParams prm, const __grid_constant__ ChainMaps maps) {
    extern __shared__ uint8_t mc_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mc_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* wring = smem;                                                             // [STAGES][hi tile | lo tile]
    __nv_bfloat16* xbuf = reinterpret_cast<__nv_bfloat16*>(smem + MC_STAGES * 2 * MC_TILE_BYTES);   // [2 ping-pong][hi,lo][R][XLD]
    float* ys = reinterpret_cast<float*>(xbuf + 2 * 2 * R * XLD);                      // [R][YLD]
    float* vecs = ys + R * YLD;                                            // [3][CHAIN_VEC_LD]  bias | ln_w | ln_b of the layer
    float* resb = vecs + 3 * CHAIN_VEC_LD;                                             // [8][MC_RES_LD]     residual rows of the layer
    __shared__ uint64_t full_bar[MC_STAGES], empty_bar[MC_STAGES];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * R;
    if (tid == 0) {
        for (int s = 0; s < MC_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dsmem_u32(&full_bar[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dsmem_u32(&empty_bar[s])), "r"(8 * CL));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
    pdl_wait();
    pdl_trigger();
    if (prm.in_partial != nullptr) {
        // input stage: warp w reduces the split-K partials of row row0 + w (K0 = 128 or 256 columns: 1 or 2 float4 per lane),
        // adds bias + residual, LayerNorms (two-pass, like torch), stores the fp32 row and stages it as bf16 (hi, lo)
        if (warp < 8)
        for (int rr = warp; rr < R; rr += 8) {
            const int K0 = prm.layer[0].K, per = K0 >> 7;
            const int row = row0 + rr;
            const bool live = row < prm.M;
            const long long zs = (long long)prm.M * K0;
            float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
            const float* base = prm.in_partial + (long long)row * K0 + 4 * lane;
            for (int z = 0; z < prm.in_nsplit; z += 6) {
                float4 t[2][6];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int u = 0; u < 6; ++u)
                        t[i][u] = (live && i < per && z + u < prm.in_nsplit) ? ldg4(base + 128 * i + (long long)(z + u) * zs) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int u = 0; u < 6; ++u) { acc[i].x += t[i][u].x; acc[i].y += t[i][u].y; acc[i].z += t[i][u].z; acc[i].w += t[i][u].w; }
            }
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (i < per) {
                    const int n = 4 * lane + 128 * i;
                    if (prm.in_bias) { const float4 b = ldg4(prm.in_bias + n); acc[i].x += b.x; acc[i].y += b.y; acc[i].z += b.z; acc[i].w += b.w; }
                    if (prm.in_res && live) { const float4 r = ldg4(prm.in_res + (long long)row * K0 + n); acc[i].x += r.x; acc[i].y += r.y; acc[i].z += r.z; acc[i].w += r.w; }
                    sum += (acc[i].x + acc[i].y) + (acc[i].z + acc[i].w);
                }
            if (prm.in_ln_w != nullptr) {
                const float mean = warp_sum(sum) / (float)K0;
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    if (i < per) {
                        const float a = acc[i].x - mean, b = acc[i].y - mean, c = acc[i].z - mean, d = acc[i].w - mean;
                        ss += (a * a + b * b) + (c * c + d * d);
                    }
                const float rstd = rsqrtf(warp_sum(ss) / (float)K0 + 1e-5f);
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    if (i < per) {
                        const int n = 4 * lane + 128 * i;
                        const float4 g = ldg4(prm.in_ln_w + n), b = ldg4(prm.in_ln_b + n);
                        acc[i].x = (acc[i].x - mean) * rstd * g.x + b.x; acc[i].y = (acc[i].y - mean) * rstd * g.y + b.y;
                        acc[i].z = (acc[i].z - mean) * rstd * g.z + b.z; acc[i].w = (acc[i].w - mean) * rstd * g.w + b.w;
                    }
            }
            __nv_bfloat16* xh = xbuf + rr * XLD;
            __nv_bfloat16* xl = xh + R * XLD;
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (i < per) {
                    const int n = 4 * lane + 128 * i;
                    const float4 v = live ? acc[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live && prm.in_out) *reinterpret_cast<float4*>(prm.in_out + (long long)row * K0 + n) = v;
                    uint32_t h0, l0, h1, l1;
                    mc_split2(v.x, v.y, h0, l0); mc_split2(v.z, v.w, h1, l1);
                    *reinterpret_cast<uint2*>(xh + n) = make_uint2(h0, h1);
                    *reinterpret_cast<uint2*>(xl + n) = make_uint2(l0, l1);
                }
        }
        __threadfence_block();             // in_out rows are re-read by this CTA (residual of a later layer)
    } else
    // stage the input rows as bf16 (hi, lo), zero-padded to the first layer's K rounded up to 64
    {
        const int K0 = prm.layer[0].K, K0p = (K0 + 63) & ~63;
        __nv_bfloat16* xh = xbuf;
        __nv_bfloat16* xl = xbuf + R * XLD;
        if ((K0 & 63) == 0 && (prm.ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(prm.x) & 15) == 0) {
            const int k4 = K0 >> 2;
            for (int i = tid; i < R * k4; i += 288) {
                const int r = i / k4, k = (i - r * k4) * 4;
                const float4 v = (row0 + r < prm.M) ? ldg4(prm.x + (long long)(row0 + r) * prm.ldx + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                uint32_t h0, l0, h1, l1;
                mc_split2(v.x, v.y, h0, l0); mc_split2(v.z, v.w, h1, l1);
                *reinterpret_cast<uint2*>(xh + r * XLD + k) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(xl + r * XLD + k) = make_uint2(l0, l1);
            }
        } else
        for (int i = tid; i < R * (K0p / 2); i += 288) {
            const int r = i / (K0p / 2), k = (i - r * (K0p / 2)) * 2;
            const bool ok = row0 + r < prm.M;
            const float a = (ok && k < K0) ? __ldg(prm.x + (long long)(row0 + r) * prm.ldx + k) : 0.f;
            const float b = (ok && k + 1 < K0) ? __ldg(prm.x + (long long)(row0 + r) * prm.ldx + k + 1) : 0.f;
            uint32_t h, l;
            mc_split2(a, b, h, l);
            *reinterpret_cast<uint32_t*>(xh + r * XLD + k) = h;
            *reinterpret_cast<uint32_t*>(xl + r * XLD + k) = l;
        }
    }
    __syncthreads();
    if (CL > 1) cluster_sync_all();            // every CTA's mbarriers are initialised before any remote arrive / multicast

    if (warp == 8) {
        // ---- producer warp: stream every layer's weight tiles in consumption order
        if (lane == 0) {
            int it = 0;
            for (int li = 0; li < prm.n_layers; ++li) {
                const ChainLayer& L = prm.layer[li];
                const int kchunks = (L.K + 63) >> 6, nblocks = (L.N + 127) >> 7;
                for (int nb = 0; nb < nblocks; ++nb)
                    for (int kc = 0; kc < kchunks; ++kc, ++it) {
                        const int stage = it % MC_STAGES;
                        if (it >= MC_STAGES) chain_mbar_wait(&empty_bar[stage], ((it / MC_STAGES) - 1) & 1);
                        uint8_t* dst = wring + stage * 2 * MC_TILE_BYTES;
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dsmem_u32(&full_bar[stage])), "r"(2 * MC_TILE_BYTES) : "memory");
                        if (CL == 1 && L.wpack != nullptr) {          // pre-tiled stream: the whole stage is one contiguous 32 KB bulk copy
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         ::"r"(dsmem_u32(dst)), "l"(L.wpack + (size_t)(nb * kchunks + kc) * (2 * MC_TILE_BYTES)), "r"(2 * MC_TILE_BYTES),
                                           "r"(dsmem_u32(&full_bar[stage])) : "memory");
                        } else if (CL == 1) {
                            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                         ::"r"(dsmem_u32(dst)), "l"(&maps.hi[li]), "r"(dsmem_u32(&full_bar[stage])), "r"(kc * 64), "r"(nb * 128) : "memory");
                            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                         ::"r"(dsmem_u32(dst + MC_TILE_BYTES)), "l"(&maps.lo[li]), "r"(dsmem_u32(&full_bar[stage])), "r"(kc * 64), "r"(nb * 128) : "memory");
                        } else {
                            constexpr int SL = 128 / CL;                       // tile rows fetched by this CTA
                            const uint16_t mask = (uint16_t)((1u << CL) - 1);
                            const uint32_t off = crank * SL * 128;
                            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                                         ::"r"(dsmem_u32(dst + off)), "l"(&maps.hi[li]), "r"(dsmem_u32(&full_bar[stage])), "r"(kc * 64), "r"(nb * 128 + (int)crank * SL), "h"(mask) : "memory");
                            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                                         ::"r"(dsmem_u32(dst + MC_TILE_BYTES + off)), "l"(&maps.lo[li]), "r"(dsmem_u32(&full_bar[stage])), "r"(kc * 64), "r"(nb * 128 + (int)crank * SL), "h"(mask) : "memory");
                        }
                    }
            }
        }
        __syncwarp();                      // lanes 1..31 wait for the producing lane: the cluster barrier below is warp-aligned
        if (CL > 1) cluster_sync_all();   // matches the consumers' final cluster barrier: no CTA exits while peers may still touch it
        return;          // the producer warp takes no part in the consumer barriers below (named barrier 1, 256 threads)
    }
    // ---- consumers (warps 0..7)
    const int g8 = lane >> 2, t4 = lane & 3;
    const int lm_r = lane & 7, lm_id = lane >> 3;
    const int a_row = 16 * warp + lm_r + 8 * (lm_id & 1);               // row of the 128-row weight tile this lane addresses
    const int a_chunk = lm_id >> 1;                                     // + 2*kstep = 16-byte chunk inside the 128 B row
    int it = 0, ping = 0;
    for (int li = 0; li < prm.n_layers; ++li) {
        const ChainLayer& L = prm.layer[li];
        const int kchunks = (L.K + 63) >> 6, nblocks = (L.N + 127) >> 7;
        const __nv_bfloat16* xh = xbuf + ping * 2 * R * XLD;
        const __nv_bfloat16* xl = xh + R * XLD;
        // epilogue operands of this layer -> shared memory, asynchronously (LDGSTS), while the GEMM below runs
        const bool res_staged = L.residual != nullptr && L.N <= MC_RES_LD;
        {
            auto cp4 = [](float* dst, const float* src) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dsmem_u32(dst)), "l"(src) : "memory");
            };
            if (L.flags & CHAIN_FLAG_VEC4) {          // 16-byte copies (N % 4 == 0, aligned operands)
                auto cp16 = [](float* dst, const float* src) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dsmem_u32(dst)), "l"(src) : "memory");
                };
                const int n4 = L.N >> 2;
                for (int g = tid; g < n4; g += 256) {
                    if (L.bias) cp16(vecs + 4 * g, L.bias + 4 * g);
                    if (L.ln_w) { cp16(vecs + CHAIN_VEC_LD + 4 * g, L.ln_w + 4 * g); cp16(vecs + 2 * CHAIN_VEC_LD + 4 * g, L.ln_b + 4 * g); }
                }
                if (res_staged)
                    for (int i = tid; i < R * n4; i += 256) {
                        const int r = i / n4, g = i - r * n4;
                        if (row0 + r < prm.M) cp16(resb + r * MC_RES_LD + 4 * g, L.residual + (long long)(row0 + r) * L.N + 4 * g);
                    }
            } else {
            for (int n = tid; n < L.N; n += 256) {
                if (L.bias) cp4(vecs + n, L.bias + n);
                if (L.ln_w) { cp4(vecs + CHAIN_VEC_LD + n, L.ln_w + n); cp4(vecs + 2 * CHAIN_VEC_LD + n, L.ln_b + n); }
            }
            if (res_staged)
                for (int i = tid; i < R * L.N; i += 256) {
                    const int r = i / L.N, n = i - r * L.N;
                    if (row0 + r < prm.M) cp4(resb + r * MC_RES_LD + n, L.residual + (long long)(row0 + r) * L.N + n);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        {
            constexpr int NT = R / 8;                                     // n8 row tiles per warp
            const uint32_t xh_addr = dsmem_u32(xh + lm_r * XLD + 8 * (lm_id & 1));
            const uint32_t xl_addr = dsmem_u32(xl + lm_r * XLD + 8 * (lm_id & 1));
            for (int nb = 0; nb < nblocks; ++nb) {
                // six independent accumulator chains per row tile (main / cross terms x k-step parity): one chain would serialise every
                // MMA of the tile behind the ~30-cycle MMA latency, this warp has no other tile to interleave with
                float acc6[NT][6][4];
#pragma unroll
                for (int t = 0; t < NT; ++t)
#pragma unroll
                    for (int c = 0; c < 6; ++c) { acc6[t][c][0] = acc6[t][c][1] = acc6[t][c][2] = acc6[t][c][3] = 0.f; }
                for (int kc = 0; kc < kchunks; ++kc, ++it) {
                    const int stage = it % MC_STAGES;
                    chain_mbar_wait(&full_bar[stage], (it / MC_STAGES) & 1);
                    const uint32_t wh = dsmem_u32(wring + stage * 2 * MC_TILE_BYTES) + a_row * 128;
                    const uint32_t wl = wh + MC_TILE_BYTES;
                    // all fragment loads of the stage first, then its MMAs: the asm statements keep their order, so interleaving them per
                    // k-step exposed one ldmatrix round trip (~30 cycles) in front of every group of three MMAs
                    uint32_t ah[4][4], al[4][4], bh[4][NT][2], bl[4][NT][2];
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t sw = (uint32_t)(((2 * ks + a_chunk) ^ (a_row & 7)) << 4);      // TMA 128-byte swizzle
                        mc_ldsm_x4(ah[ks], wh + sw);
                        mc_ldsm_x4(al[ks], wl + sw);
                        const uint32_t xo = (uint32_t)((kc * 64 + ks * 16) * 2);
#pragma unroll
                        for (int t = 0; t < NT; ++t) {
                            mc_ldsm_x2(bh[ks][t], xh_addr + xo + (uint32_t)(t * 8 * XLD * 2));
                            mc_ldsm_x2(bl[ks][t], xl_addr + xo + (uint32_t)(t * 8 * XLD * 2));
                        }
                    }
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                        for (int t = 0; t < NT; ++t) {
                            mc_mma(acc6[t][ks & 1], ah[ks], bh[ks][t]); mc_mma(acc6[t][2 + (ks & 1)], al[ks], bh[ks][t]); mc_mma(acc6[t][4 + (ks & 1)], ah[ks], bl[ks][t]);
                        }
                    __syncwarp();
                    if (CL == 1) {
                        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dsmem_u32(&empty_bar[stage])) : "memory");
                    } else if (lane < CL) {                  // lane r releases the stage towards CTA r of the cluster
                        uint32_t remote;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(dsmem_u32(&empty_bar[stage])), "r"(lane));
                        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
                    }
                }
                // D fragment: (feature 16w+g8 [+8], row 8t + 2t4 [+1])
                const int n = nb * 128 + 16 * warp + g8;
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    float acc[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i] = ((acc6[t][2][i] + acc6[t][3][i]) + (acc6[t][4][i] + acc6[t][5][i])) + (acc6[t][0][i] + acc6[t][1][i]);
                    float* y0 = ys + (8 * t + 2 * t4) * YLD;
                    if (n < YLD - 4) { y0[n] = acc[0]; y0[YLD + n] = acc[1]; }
                    if (n + 8 < YLD - 4) { y0[n + 8] = acc[2]; y0[YLD + n + 8] = acc[3]; }
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int rr = warp; rr < R; rr += 8) {
        if (L.flags & CHAIN_FLAG_VEC4) {
            const bool more = li + 1 < prm.n_layers;
            __nv_bfloat16* nh = more ? xbuf + (ping ^ 1) * 2 * R * XLD + rr * XLD : nullptr;
            chain_row_epilogue_v4(prm, L, row0 + rr, ys + rr * YLD, lane, vecs, res_staged ? resb + rr * MC_RES_LD : nullptr,
                                  nh, more ? nh + R * XLD : nullptr, more ? ((prm.layer[li + 1].K + 63) & ~63) : 0);
        } else {
            float* yr = ys + rr * YLD;
            chain_row_epilogue(prm, L, row0 + rr, yr, lane, vecs, res_staged ? resb + rr * MC_RES_LD : nullptr);
            __syncwarp();
            if ((L.flags & CHAIN_FLAG_POINTS) && row0 + rr < prm.M) {           // lane gp: sample point gp of this query
                const long long row = row0 + rr;
                for (int gp = lane; gp < prm.sp_GP; gp += 32)
                    sample_point_one(prm.sp_bbox + row * 10, yr + prm.sp_off_col + gp * 3, yr + prm.sp_log_col + gp * prm.sp_L, prm.sp_L,
                                     prm.sp_r[0], prm.sp_r[1], prm.sp_r[2], prm.sp_r[3], prm.sp_r[4], prm.sp_r[5],
                                     prm.sp_points + (row * prm.sp_GP + gp) * 3, prm.sp_scale_w + (row * prm.sp_GP + gp) * prm.sp_L, prm.sp_ssign);
            }
            if (li + 1 < prm.n_layers) {          // next layer's activation operand: bf16 (hi, lo), zero beyond N up to its padded K
                __nv_bfloat16* nh = xbuf + (ping ^ 1) * 2 * R * XLD + rr * XLD;
                __nv_bfloat16* nl = nh + R * XLD;
                const int Kn = (prm.layer[li + 1].K + 63) & ~63;
                for (int k = 2 * lane; k < Kn; k += 64) {
                    const float a = (k < L.N) ? yr[k] : 0.f, b = (k + 1 < L.N) ? yr[k + 1] : 0.f;
                    uint32_t h, l;
                    mc_split2(a, b, h, l);
                    *reinterpret_cast<uint32_t*>(nh + k) = h;
                    *reinterpret_cast<uint32_t*>(nl + k) = l;
                }
            }
        }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        ping ^= 1;
    }
    if (CL > 1) cluster_sync_all();
}

// ------------------------------------------------------------------------------------------------
// N-split cluster chain (option dense_nsplit = 2 | 4).  The kernels above give every CTA 8 rows and make it stream ALL
// the weights of the chain (3.9 MB per row group and decoder layer, x 113 CTAs = 440 MB of L2->SM traffic): that stream,
// not the math, is what they wait for.  Here a cluster of CL CTAs owns R = 8*CL rows and splits every layer's OUTPUT
// FEATURES: CTA r computes the FB = 256/CL-wide feature blocks r, r+CL, ... for all R rows, so it streams 1/CL of the
// weights and every weight byte leaves L2 once per R rows instead of once per 8.  Rows sit on the MMA N dimension as
// before (two n8 tiles per warp, so each A = weight fragment is used twice); warp w owns m16 tile w % MB of the block
// and rows 16*(w / MB) .. +15.
//   Exchange: a CTA adds bias (+ the pre-LN residual) to its slice, writes it to its own ys[R][N] (fp32) and pushes it
// with 16-byte st.shared::cluster stores into the ys of its CL-1 peers.  Hand-over is by two cluster-scope mbarriers per
// CTA, each expecting one arrival per warp of the cluster: ys_full (all slices of the layer have landed; waited with
// acquire.cluster before the row epilogue) and ys_free (every CTA finished reading the previous layer's ys; waited
// before the next push, normally long satisfied).  Every CTA then runs the row epilogue (LayerNorm over the full row,
// ReLU, post-LN residual) for all R rows redundantly -- a few hundred cycles, and it leaves the next layer's bf16
// (hi, lo) operand in local shared memory without a second exchange -- while global outputs of row r are written by
// CTA r / 8 only.  A last layer without LayerNorm needs no full rows: each CTA finishes and stores its own feature
// blocks ("slice mode": in-projection + tau, sampling heads, cls / reg outputs, refine).
//   The dedicated producer warp never takes part in the hand-over (mbarriers, not barrier.cluster), so weight tiles of
// the following layers keep streaming while the consumers exchange: the ring is refilled from the moment a stage drains.
template <int CL>
struct NsCfg {
    static constexpr int R = 8 * CL;                 // rows per cluster
    static constexpr int FB = 256 / CL;              // features per weight block (TMA box rows)
    static constexpr int MB = FB / 16;               // m16 tiles per block
    static constexpr int STAGES = CL == 4 ? 6 : 4;
    static constexpr int TILE_BYTES = FB * 128;      // one (hi or lo) FB x 64 bf16 tile, 128-byte swizzled rows
    static constexpr int XLD = 512 + 8;              // bf16 row stride of the activation operand (K <= 512)
    static constexpr int YLD = 512 + 4;              // fp32 row stride of ys (exchanged layers: N <= 512)
    static constexpr size_t SMEM = (size_t)STAGES * 2 * TILE_BYTES + (size_t)2 * R * XLD * 2 + (size_t)R * YLD * 4 + 1024;
};

__device__ __forceinline__ void ns_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(dsmem_u32(bar)), "r"(parity) : "memory");
        if (!done && spins > (1u << 26)) __trap();
    }
}
// one arrival (release, cluster scope) on the copy of `bar` that lives in CTA `rank` of the cluster
__device__ __forceinline__ void ns_arrive_remote(uint64_t* bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(dsmem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ uint32_t ns_mapa(uint32_t addr, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(rank));
    return remote;
}

template <int CL>
__global__ void __launch_bounds__(288, 1)
dense_chain_ns_kernel(const __grid_constant__ ChainParams prm, const __grid_constant__ ChainMaps maps) {
    using C = NsCfg<CL>;
    constexpr int R = C::R, FB = C::FB, MB = C::MB, STAGES = C::STAGES, TILE = C::TILE_BYTES, XLD = C::XLD, YLD = C::YLD;
    extern __shared__ uint8_t ns_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ns_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* wring = smem;                                                                  // [STAGES][hi tile | lo tile]
    __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(smem + STAGES * 2 * TILE);         // [R][XLD]
    __nv_bfloat16* xl = xh + R * XLD;                                                       // [R][XLD]
    float* ys = reinterpret_cast<float*>(xl + R * XLD);                                     // [R][YLD]
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], ys_full, ys_free, xin_bar;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = cluster_ctarank();
    const int row0 = (blockIdx.x / CL) * R;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dsmem_u32(&full_bar[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dsmem_u32(&empty_bar[s])), "r"(8));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dsmem_u32(&ys_full)), "r"(8 * CL));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dsmem_u32(&ys_free)), "r"(8 * CL));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dsmem_u32(&xin_bar)), "r"(8 * CL));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();                 // every CTA's mbarriers exist before any remote arrive / store (no global access yet)
    pdl_wait();
    pdl_trigger();

    if (warp == 8) {
        // ---- producer warp: this CTA's weight tiles of every layer, in consumption order
        if (lane == 0) {
            int it = 0;
            for (int li = 0; li < prm.n_layers; ++li) {
                const ChainLayer& L = prm.layer[li];
                const int kchunks = (L.K + 63) >> 6, nblk = (L.N + FB - 1) / FB;
                for (int gb = (int)crank; gb < nblk; gb += CL)
                    for (int kc = 0; kc < kchunks; ++kc, ++it) {
                        const int stage = it % STAGES;
                        if (it >= STAGES) chain_mbar_wait(&empty_bar[stage], ((it / STAGES) - 1) & 1);
                        uint8_t* dst = wring + stage * 2 * TILE;
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dsmem_u32(&full_bar[stage])), "r"(2 * TILE) : "memory");
                        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                     ::"r"(dsmem_u32(dst)), "l"(&maps.hi[li]), "r"(dsmem_u32(&full_bar[stage])), "r"(kc * 64), "r"(gb * FB) : "memory");
                        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                     ::"r"(dsmem_u32(dst + TILE)), "l"(&maps.lo[li]), "r"(dsmem_u32(&full_bar[stage])), "r"(kc * 64), "r"(gb * FB) : "memory");
                    }
            }
        }
        __syncwarp();
        cluster_sync_all();             // matches the consumers' final cluster barrier
        return;
    }

    // ---- consumers (warps 0..7)
    const int g8 = lane >> 2, t4 = lane & 3;
    const int lm_r = lane & 7, lm_id = lane >> 3;
    const int mt = warp % MB, rg = warp / MB;                           // m16 tile of the block, 16-row group
    const int a_row = 16 * mt + lm_r + 8 * (lm_id & 1);                 // weight-tile row this lane addresses (ldmatrix.x4)
    const int a_chunk = lm_id >> 1;
    const int b_row = 16 * rg + 8 * (lm_id >> 1) + lm_r;                // activation row this lane addresses (ldmatrix.x4: two n8 tiles)
    const uint32_t xh_addr = dsmem_u32(xh + b_row * XLD + 8 * (lm_id & 1));
    const uint32_t xl_addr = dsmem_u32(xl + b_row * XLD + 8 * (lm_id & 1));
    const uint32_t ys_u32 = dsmem_u32(ys), xh_u32 = dsmem_u32(xh), xl_u32 = dsmem_u32(xl);

    // ---- input stage
    if (prm.in_partial != nullptr) {
        // warp w of CTA r reduces the split-K partials of row 8r + w (K0 = 128 or 256), adds bias + residual, LayerNorms
        // (two-pass, like torch), stores the fp32 row and pushes it as bf16 (hi, lo) into the operand buffer of every CTA
        const int K0 = prm.layer[0].K, per = K0 >> 7;
        const int lr = 8 * (int)crank + warp;
        const int row = row0 + lr;
        const bool live = row < prm.M;
        const long long zs = (long long)prm.M * K0;
        float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
        const float* base = prm.in_partial + (long long)row * K0 + 4 * lane;
        for (int z = 0; z < prm.in_nsplit; z += 6) {
            float4 t[2][6];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int u = 0; u < 6; ++u)
                    t[i][u] = (live && i < per && z + u < prm.in_nsplit) ? ldg4(base + 128 * i + (long long)(z + u) * zs) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int u = 0; u < 6; ++u) { acc[i].x += t[i][u].x; acc[i].y += t[i][u].y; acc[i].z += t[i][u].z; acc[i].w += t[i][u].w; }
        }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (i < per) {
                const int n = 4 * lane + 128 * i;
                if (prm.in_bias) { const float4 b = ldg4(prm.in_bias + n); acc[i].x += b.x; acc[i].y += b.y; acc[i].z += b.z; acc[i].w += b.w; }
                if (prm.in_res && live) { const float4 r = ldg4(prm.in_res + (long long)row * K0 + n); acc[i].x += r.x; acc[i].y += r.y; acc[i].z += r.z; acc[i].w += r.w; }
                sum += (acc[i].x + acc[i].y) + (acc[i].z + acc[i].w);
            }
        if (prm.in_ln_w != nullptr) {
            const float mean = warp_sum(sum) / (float)K0;
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (i < per) {
                    const float a = acc[i].x - mean, b = acc[i].y - mean, c = acc[i].z - mean, d = acc[i].w - mean;
                    ss += (a * a + b * b) + (c * c + d * d);
                }
            const float rstd = rsqrtf(warp_sum(ss) / (float)K0 + 1e-5f);
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (i < per) {
                    const int n = 4 * lane + 128 * i;
                    const float4 g = ldg4(prm.in_ln_w + n), b = ldg4(prm.in_ln_b + n);
                    acc[i].x = (acc[i].x - mean) * rstd * g.x + b.x; acc[i].y = (acc[i].y - mean) * rstd * g.y + b.y;
                    acc[i].z = (acc[i].z - mean) * rstd * g.z + b.z; acc[i].w = (acc[i].w - mean) * rstd * g.w + b.w;
                }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (i < per) {
                const int n = 4 * lane + 128 * i;
                const float4 v = live ? acc[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                if (live && prm.in_out) *reinterpret_cast<float4*>(prm.in_out + (long long)row * K0 + n) = v;
                uint32_t h0, l0, h1, l1;
                mc_split2(v.x, v.y, h0, l0); mc_split2(v.z, v.w, h1, l1);
                const uint32_t off = (uint32_t)((lr * XLD + n) * 2);
#pragma unroll
                for (int p = 0; p < CL; ++p) {
                    asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(ns_mapa(xh_u32 + off, p)), "r"(h0), "r"(h1) : "memory");
                    asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(ns_mapa(xl_u32 + off, p)), "r"(l0), "r"(l1) : "memory");
                }
            }
        // the in_out rows are re-read by OTHER CTAs of the cluster (pre-LN residual of a later layer, L2 loads): the
        // release below orders the global stores too
        __syncwarp();
        if (lane < CL) ns_arrive_remote(&xin_bar, lane);
        ns_wait_cluster(&xin_bar, 0);
    } else {
        // every CTA stages all R input rows as bf16 (hi, lo), zero-padded to the first layer's K rounded up to 64
        const int K0 = prm.layer[0].K, K0p = (K0 + 63) & ~63;
        for (int i = tid; i < R * (K0p / 2); i += 256) {
            const int r = i / (K0p / 2), k = (i - r * (K0p / 2)) * 2;
            const bool ok = row0 + r < prm.M;
            const float a = (ok && k < K0) ? __ldg(prm.x + (long long)(row0 + r) * prm.ldx + k) : 0.f;
            const float b = (ok && k + 1 < K0) ? __ldg(prm.x + (long long)(row0 + r) * prm.ldx + k + 1) : 0.f;
            uint32_t h, l;
            mc_split2(a, b, h, l);
            *reinterpret_cast<uint32_t*>(xh + r * XLD + k) = h;
            *reinterpret_cast<uint32_t*>(xl + r * XLD + k) = l;
        }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");

    int it = 0, ex = 0;                  // weight tiles consumed; exchanged layers completed
    for (int li = 0; li < prm.n_layers; ++li) {
        const ChainLayer& L = prm.layer[li];
        const int N = L.N;
        const int kchunks = (L.K + 63) >> 6, nblk = (N + FB - 1) / FB;
        const bool last = li + 1 == prm.n_layers;
        const bool exchange = !(last && L.ln_w == nullptr);
        const bool pre_res = (L.flags & SBEV_DENSE_RES_PRE_LN) && L.residual != nullptr;
        int bi = 0;
        for (int gb = (int)crank; gb < nblk; gb += CL, ++bi) {
            // D fragment of n8 tile j: (feature f0 [+8], row r0 + 8j [+1])
            const int f0 = gb * FB + 16 * mt + g8;
            const int r0 = 16 * rg + 2 * t4;
            float pre[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int f = f0 + 8 * (i >> 1), r = row0 + r0 + 8 * j + (i & 1);
                    float v = (L.bias != nullptr && f < N) ? __ldg(L.bias + f) : 0.f;
                    if (pre_res && f < N && r < prm.M) v += __ldcg(L.residual + (long long)r * N + f);
                    pre[j][i] = v;
                }
            float acc[2][2][2][4];       // [k-step parity][main | cross][n8 tile][fragment]: 8 independent accumulator chains
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[a][b][j][i] = 0.f;
            for (int kc = 0; kc < kchunks; ++kc, ++it) {
                const int stage = it % STAGES;
                chain_mbar_wait(&full_bar[stage], (it / STAGES) & 1);
                const uint32_t wh = dsmem_u32(wring + stage * 2 * TILE) + a_row * 128;
                const uint32_t wl = wh + TILE;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t ah[4], al[4], bh[4], bl[4];
                    const uint32_t sw = (uint32_t)(((2 * ks + a_chunk) ^ (a_row & 7)) << 4);      // TMA 128-byte swizzle
                    mc_ldsm_x4(ah, wh + sw);
                    mc_ldsm_x4(al, wl + sw);
                    const uint32_t xo = (uint32_t)((kc * 64 + ks * 16) * 2);
                    mc_ldsm_x4(bh, xh_addr + xo);
                    mc_ldsm_x4(bl, xl_addr + xo);
                    const uint32_t bh0[2] = {bh[0], bh[1]}, bh1[2] = {bh[2], bh[3]}, bl0[2] = {bl[0], bl[1]}, bl1[2] = {bl[2], bl[3]};
                    mc_mma(acc[ks & 1][0][0], ah, bh0); mc_mma(acc[ks & 1][0][1], ah, bh1);
                    mc_mma(acc[ks & 1][1][0], al, bh0); mc_mma(acc[ks & 1][1][1], al, bh1);
                    mc_mma(acc[ks & 1][1][0], ah, bl0); mc_mma(acc[ks & 1][1][1], ah, bl1);
                }
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dsmem_u32(&empty_bar[stage])) : "memory");
            }
            // own slice (+ bias + pre-LN residual) -> local ys; exchanged layers index ys by global feature, slice mode by own block
            const int c0 = exchange ? f0 : bi * FB + 16 * mt + g8;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float v = ((acc[0][1][j][i] + acc[1][1][j][i]) + (acc[0][0][j][i] + acc[1][0][j][i])) + pre[j][i];
                    ys[(r0 + 8 * j + (i & 1)) * YLD + c0 + 8 * (i >> 1)] = v;
                }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");

        if (!exchange) {
            // ---- slice mode: finish and store this CTA's feature blocks (elementwise epilogue)
            const bool refine = (L.flags & SBEV_DENSE_REFINE) != 0;
            bi = 0;
            for (int gb = (int)crank; gb < nblk; gb += CL, ++bi)
                for (int i = tid; i < R * FB; i += 256) {
                    const int r = i / FB, fl = i - r * FB;
                    const int n = gb * FB + fl, row = row0 + r;
                    if (n >= N || row >= prm.M) continue;
                    float v = ys[r * YLD + bi * FB + fl];
                    if (L.flags & SBEV_DENSE_RELU) v = fmaxf(v, 0.f);
                    if (!pre_res && L.residual != nullptr) v += __ldcg(L.residual + (long long)row * N + n);
                    if (refine) {
                        // refine_bbox + velocity rescale (sparsebev_transformer.py:155-160,179-183)
                        if (n < 3) {
                            const float x = fminf(fmaxf(__ldg(prm.aux_proposal + (long long)row * N + n), 0.f), 1.f);
                            v = v + logf(__fdiv_rn(fmaxf(x, 1e-5f), fmaxf(1.f - x, 1e-5f)));
                            v = __fdiv_rn(1.f, 1.f + expf(-v));
                        } else if (n >= 8 && prm.aux_T > 1) {
                            float td = __ldg(prm.aux_time_diff + (row / prm.aux_Q) * prm.aux_T + 1);
                            if (td < 1e-5f) td = 1.0f;
                            v = __fdiv_rn(v, td);
                        }
                    }
                    if (L.y != nullptr) L.y[(long long)row * L.ldy + n] = v;
                    if (L.y_hi != nullptr) {
                        const __nv_bfloat16 h = __float2bfloat16_rn(v);
                        L.y_hi[(long long)row * L.ldy + n] = h;
                        L.y_lo[(long long)row * L.ldy + n] = __float2bfloat16_rn(v - __bfloat162float(h));
                    }
                }
            continue;                    // (slice mode is the last layer)
        }

        // ---- exchange: push this CTA's feature blocks into every peer's ys
        if (ex > 0) ns_wait_cluster(&ys_free, (ex - 1) & 1);            // every CTA is done reading the previous layer's ys
        for (int gb = (int)crank; gb < nblk; gb += CL)
            for (int i = tid; i < R * (FB / 4); i += 256) {
                const int r = i / (FB / 4), c = gb * FB + 4 * (i - r * (FB / 4));
                if (c >= N) continue;
                const uint32_t off = (uint32_t)((r * YLD + c) * 4);
                const float4 v = *reinterpret_cast<const float4*>(ys + r * YLD + c);
#pragma unroll
                for (int p = 0; p < CL; ++p)
                    if (p != (int)crank)
                        asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ns_mapa(ys_u32 + off, p)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
        const int per = (N + 127) >> 7;                                 // float4 per lane per row (N <= 512, N % 4 == 0)
        __syncwarp();
        if (lane < CL) ns_arrive_remote(&ys_full, lane);
        ns_wait_cluster(&ys_full, ex & 1);

        // ---- row epilogue on full rows: warp w finishes rows w, w + 8, ...; row r's global outputs belong to CTA r / 8
        const bool has_ln = L.ln_w != nullptr;
        const bool post_res = !pre_res && L.residual != nullptr;
        const int Kn = last ? 0 : ((prm.layer[last ? li : li + 1].K + 63) & ~63);
        // (the CL rows of a warp are processed side by side so that their shuffle reductions overlap)
        float4 v[CL][4];
        float mean[CL], rstd[CL];
        bool act[CL];
#pragma unroll
        for (int j = 0; j < CL; ++j) {
            act[j] = !last || j == (int)crank;
            const int r = warp + 8 * j;
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int n = 4 * lane + 128 * i;
                v[j][i] = (act[j] && i < per && n < N) ? *reinterpret_cast<const float4*>(ys + r * YLD + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                s += (v[j][i].x + v[j][i].y) + (v[j][i].z + v[j][i].w);
            }
            mean[j] = s;
            rstd[j] = 1.f;
        }
        if (has_ln) {
#pragma unroll
            for (int j = 0; j < CL; ++j) mean[j] = warp_sum(mean[j]) / (float)N;
#pragma unroll
            for (int j = 0; j < CL; ++j) {
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n = 4 * lane + 128 * i;
                    if (i < per && n < N) {
                        const float a = v[j][i].x - mean[j], b = v[j][i].y - mean[j], c = v[j][i].z - mean[j], d = v[j][i].w - mean[j];
                        ss += (a * a + b * b) + (c * c + d * d);
                    }
                }
                rstd[j] = ss;
            }
#pragma unroll
            for (int j = 0; j < CL; ++j) rstd[j] = rsqrtf(warp_sum(rstd[j]) / (float)N + 1e-5f);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = 4 * lane + 128 * i;
            if (i < per && n < N) {
                float4 g = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (has_ln) { g = ldg4(L.ln_w + n); b = ldg4(L.ln_b + n); }
#pragma unroll
                for (int j = 0; j < CL; ++j) {
                    if (!act[j]) continue;
                    const bool owner = j == (int)crank;
                    const int r = warp + 8 * j, row = row0 + r;
                    const bool live = row < prm.M;
                    float4 o = v[j][i];
                    if (has_ln) {
                        o.x = (o.x - mean[j]) * rstd[j] * g.x + b.x; o.y = (o.y - mean[j]) * rstd[j] * g.y + b.y;
                        o.z = (o.z - mean[j]) * rstd[j] * g.z + b.z; o.w = (o.w - mean[j]) * rstd[j] * g.w + b.w;
                    }
                    if (L.flags & SBEV_DENSE_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    if (post_res && live) {
                        const float4 q = __ldcg(reinterpret_cast<const float4*>(L.residual + (long long)row * N + n));
                        o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
                    }
                    if (!live) o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (owner && live) {
                        const float ov[4] = {o.x, o.y, o.z, o.w};
                        if (L.y != nullptr) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) L.y[(long long)row * L.ldy + n + e] = ov[e];
                        }
                        if (L.y_hi != nullptr) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const __nv_bfloat16 h = __float2bfloat16_rn(ov[e]);
                                L.y_hi[(long long)row * L.ldy + n + e] = h;
                                L.y_lo[(long long)row * L.ldy + n + e] = __float2bfloat16_rn(ov[e] - __bfloat162float(h));
                            }
                        }
                    }
                    if (!last) {
                        uint32_t h0, l0, h1, l1;
                        mc_split2(o.x, o.y, h0, l0); mc_split2(o.z, o.w, h1, l1);
                        *reinterpret_cast<uint2*>(xh + r * XLD + n) = make_uint2(h0, h1);
                        *reinterpret_cast<uint2*>(xl + r * XLD + n) = make_uint2(l0, l1);
                    }
                }
            }
        }
        if (!last)                                                      // zero beyond N up to the next layer's padded K
#pragma unroll
            for (int j = 0; j < CL; ++j)
                for (int k = N + 4 * lane; k < Kn; k += 128) {
                    *reinterpret_cast<uint2*>(xh + (warp + 8 * j) * XLD + k) = make_uint2(0u, 0u);
                    *reinterpret_cast<uint2*>(xl + (warp + 8 * j) * XLD + k) = make_uint2(0u, 0u);
                }
        __syncwarp();
        if (lane < CL) ns_arrive_remote(&ys_free, lane);                // this warp is done reading ys
        ++ex;
        asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    cluster_sync_all();                  // no CTA exits while peers may still store into it or arrive on its barriers
}

// Box decode + offsets -> lidar-frame sample points; softmax over levels.  Thread per point.
__global__ void __launch_bounds__(128)
sample_points_kernel(const float* __restrict__ query_bbox, const float* __restrict__ offset, int ld_off,
                     const float* __restrict__ logits, int ld_log, float r0, float r1, float r2, float r3, float r4, float r5,
                     int BQ, int GP, int L, float* __restrict__ points, float* __restrict__ scale_w, float ssign) {
    pdl_wait();
    pdl_trigger();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)BQ * GP) return;
    const long long bq = idx / GP;
    const int gp = (int)(idx - bq * GP);
    sample_point_one(query_bbox + bq * 10, offset + bq * ld_off + gp * 3, logits + bq * ld_log + gp * L, L, r0, r1, r2, r3, r4, r5,
                     points + idx * 3, scale_w + idx * L, ssign);
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

// out[row] = LN(sum_z partial[z][row] + bias + residual[row]); one 256-thread CTA per row, N <= 1024, N % 4 == 0.
// Thread t owns the 16-byte column group t % (N/4) of the split-K slices z = t / (N/4), + zpar, ...: every partial of the
// row is requested at once (the slices were just written by the GEMM and sit in L2), then the slice sums meet in shared
// memory and the first N/4 threads finish bias + residual + LayerNorm (two-pass variance, like torch).
__global__ void __launch_bounds__(256)
reduce_ln_kernel(const float* __restrict__ partial, int nsplit, const float* __restrict__ bias,
                 const float* __restrict__ residual, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                 int M, int N, float* __restrict__ out) {
    __shared__ float4 part[256];
    __shared__ float red[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = blockIdx.x;
    const int ng = N >> 2;                       // 16-byte column groups (<= 256)
    const int zpar = 256 / ng;                   // slices summed side by side (>= 1)
    const int c = tid % ng, h = tid / ng;
    const long long zs = (long long)M * N;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    pdl_wait();
    pdl_trigger();
    if (h < zpar) {
        const float* p = partial + (long long)row * N + 4 * c;
        for (int z = h; z < nsplit; z += 8 * zpar) {
            float4 t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = (z + u * zpar < nsplit) ? ldg4(p + (long long)(z + u * zpar) * zs) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) { acc.x += t[u].x; acc.y += t[u].y; acc.z += t[u].z; acc.w += t[u].w; }
        }
    }
    part[tid] = acc;
    __syncthreads();
    const bool owner = tid < ng;
    float s = 0.f;
    if (owner) {
        for (int j = 1; j < zpar; ++j) { const float4 t = part[tid + j * ng]; acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
        if (bias) { const float4 t = ldg4(bias + 4 * c); acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
        if (residual) { const float4 t = ldg4(residual + (long long)row * N + 4 * c); acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
        s = (acc.x + acc.y) + (acc.z + acc.w);
    }
    if (ln_w != nullptr) {                       // uniform branch: block reductions over the (<= 8) warps
        s = warp_sum(s);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += red[w];
        const float mean = tot / (float)N;
        __syncthreads();
        float ss = 0.f;
        if (owner) {
            const float a = acc.x - mean, b = acc.y - mean, cc = acc.z - mean, d = acc.w - mean;
            ss = (a * a + b * b) + (cc * cc + d * d);
        }
        ss = warp_sum(ss);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        float var = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) var += red[w];
        const float rstd = rsqrtf(var / (float)N + 1e-5f);
        if (owner) {
            const float4 g = ldg4(ln_w + 4 * c), b = ldg4(ln_b + 4 * c);
            acc.x = (acc.x - mean) * rstd * g.x + b.x; acc.y = (acc.y - mean) * rstd * g.y + b.y;
            acc.z = (acc.z - mean) * rstd * g.z + b.z; acc.w = (acc.w - mean) * rstd * g.w + b.w;
        }
    }
    if (owner) *reinterpret_cast<float4*>(out + (long long)row * N + 4 * c) = acc;
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

struct ChainInputReduce { const float* partial; int nsplit; const float* bias; const float* residual; const float* ln_w; const float* ln_b; float* out; };
struct ChainPoints { const float* bbox; const float* pc_range; int GP, L, off_col, log_col; float* points; float* scale_w; };
extern "C" int sbev_sample_points_fwd(const float* query_bbox, const float* offset, int ld_off, const float* scale_logits, int ld_log,
                                      const float* pc_range, int BQ, int GP, int L, float* points, float* scale_w, void* stream);

static int dense_chain_impl(const float* x, int ldx, const ChainInputReduce* in, int M, int n_layers, const sbev_dense_layer* layers,
                            const float* refine_proposal, const float* refine_time_diff, int refine_Q, int refine_T,
                            void* stream, const ChainPoints* pts = nullptr) {
    SBEV_REQUIRE((x || in) && layers, SBEV_ERR_INVALID, "sbev_dense_chain_fwd: null pointer");
    SBEV_REQUIRE(n_layers >= 1 && n_layers <= CHAIN_MAX_LAYERS, SBEV_ERR_UNSUPPORTED, "sbev_dense_chain_fwd: 1..%d layers", CHAIN_MAX_LAYERS);
    SBEV_REQUIRE(M >= 0 && ldx >= layers[0].K, SBEV_ERR_INVALID, "sbev_dense_chain_fwd: bad sizes");
    ChainParams prm;
    bool use_mma = get_option(OPT_DENSE_IMPL) == 0;
    prm.x = x; prm.ldx = ldx; prm.M = M; prm.n_layers = n_layers;
    prm.in_partial = nullptr; prm.in_nsplit = 0; prm.in_bias = prm.in_res = prm.in_ln_w = prm.in_ln_b = nullptr; prm.in_out = nullptr;
    if (in) {
        prm.in_partial = in->partial; prm.in_nsplit = in->nsplit; prm.in_bias = in->bias; prm.in_res = in->residual;
        prm.in_ln_w = in->ln_w; prm.in_ln_b = in->ln_b; prm.in_out = in->out;
    }
    prm.aux_proposal = refine_proposal; prm.aux_time_diff = refine_time_diff; prm.aux_Q = refine_Q > 0 ? refine_Q : 1; prm.aux_T = refine_T;
    prm.sp_bbox = nullptr; prm.sp_points = prm.sp_scale_w = nullptr; prm.sp_GP = prm.sp_L = prm.sp_off_col = prm.sp_log_col = 0;
    for (int i = 0; i < 6; ++i) prm.sp_r[i] = 0.f;
    prm.sp_ssign = get_option(OPT_LEGACY_ROTATION) ? -1.f : 1.f;
    int act = 4;
    for (int i = 0; i < n_layers; ++i) {
        const sbev_dense_layer& l = layers[i];
        SBEV_REQUIRE((l.Wt != nullptr || l.W_hi != nullptr) && l.K > 0 && l.N > 0, SBEV_ERR_INVALID, "sbev_dense_chain_fwd: layer %d: bad weight/sizes", i);
        SBEV_REQUIRE(l.ldw >= l.N && (l.ldw & 3) == 0 && l.ldw <= 256 * CHAIN_MAX_PASS, SBEV_ERR_UNSUPPORTED,
                     "sbev_dense_chain_fwd: layer %d: ldw must be a multiple of 4, >= N and <= %d", i, 256 * CHAIN_MAX_PASS);
        SBEV_REQUIRE((l.ln_w == nullptr) == (l.ln_b == nullptr), SBEV_ERR_INVALID, "sbev_dense_chain_fwd: ln_w and ln_b go together");
        SBEV_REQUIRE((reinterpret_cast<uintptr_t>(l.Wt) & 15) == 0, SBEV_ERR_INVALID, "sbev_dense_chain_fwd: Wt not 16-byte aligned");
        SBEV_REQUIRE(i == 0 || l.K == layers[i - 1].N, SBEV_ERR_INVALID, "sbev_dense_chain_fwd: layer %d K != previous N", i);
        SBEV_REQUIRE(l.y == nullptr || l.ldy >= l.N, SBEV_ERR_INVALID, "sbev_dense_chain_fwd: layer %d: ldy < N", i);
        if (l.flags & SBEV_DENSE_REFINE)
            SBEV_REQUIRE(refine_proposal && refine_time_diff && l.N >= 10 && refine_T >= 1, SBEV_ERR_INVALID, "sbev_dense_chain_fwd: refine needs proposal/time_diff");
        if (l.W_hi == nullptr || l.W_lo == nullptr || l.K > 512 || l.N > 1024) use_mma = false;
        ChainLayer& c = prm.layer[i];
        c.Wt = l.Wt; c.bias = l.bias; c.ln_w = l.ln_w; c.ln_b = l.ln_b; c.residual = l.residual; c.y = l.y;
        c.y_hi = reinterpret_cast<__nv_bfloat16*>(const_cast<uint16_t*>(l.y_hi)); c.y_lo = reinterpret_cast<__nv_bfloat16*>(const_cast<uint16_t*>(l.y_lo));
        SBEV_REQUIRE((l.y_hi == nullptr) == (l.y_lo == nullptr), SBEV_ERR_INVALID, "sbev_dense_chain_fwd: y_hi and y_lo go together");
        c.ldw = l.ldw; c.K = l.K; c.N = l.N; c.flags = l.flags & 0xff & ~SBEV_DENSE_WIDE_CTA; c.ldy = l.ldy;
        c.wpack = get_option(OPT_DENSE_PACK) ? reinterpret_cast<const uint8_t*>(l.W_pack) : nullptr;
        SBEV_REQUIRE((reinterpret_cast<uintptr_t>(l.W_pack) & 127) == 0, SBEV_ERR_INVALID, "sbev_dense_chain_fwd: layer %d: W_pack not 128-byte aligned", i);
        {
            const uintptr_t a16 = reinterpret_cast<uintptr_t>(l.bias) | reinterpret_cast<uintptr_t>(l.ln_w) | reinterpret_cast<uintptr_t>(l.ln_b) |
                                  reinterpret_cast<uintptr_t>(l.residual) | reinterpret_cast<uintptr_t>(l.y);
            const uintptr_t a8 = reinterpret_cast<uintptr_t>(l.y_hi) | reinterpret_cast<uintptr_t>(l.y_lo);
            if ((l.N & 3) == 0 && l.N <= 1024 && !(l.flags & SBEV_DENSE_REFINE) && (l.ldy & 3) == 0 && (a16 & 15) == 0 && (a8 & 7) == 0 &&
                get_option(OPT_DENSE_VEC4))
                c.flags |= CHAIN_FLAG_VEC4;
        }
        int kc = CHAIN_STAGE_FLOATS / l.ldw;
        if (kc >= 4) kc &= ~3;
        if (kc > l.K) kc = l.K;
        c.kc = kc;
        act = act > ((l.K + 3) & ~3) ? act : ((l.K + 3) & ~3);
        act = act > l.ldw ? act : l.ldw;
    }
    SBEV_REQUIRE(layers[n_layers - 1].y != nullptr, SBEV_ERR_INVALID, "sbev_dense_chain_fwd: last layer needs an output pointer");
    prm.act_ld = act + 4;
    if (M == 0) return SBEV_OK;
    bool fuse_points = false;
    if (pts != nullptr) {
        const sbev_dense_layer& ll = layers[n_layers - 1];
        SBEV_REQUIRE(pts->bbox && pts->pc_range && pts->points && pts->scale_w && pts->GP > 0 && pts->L >= 1 && pts->L <= SBEV_MAX_LEVELS &&
                     pts->off_col >= 0 && pts->log_col >= 0 && pts->off_col + pts->GP * 3 <= ll.N && pts->log_col + pts->GP * pts->L <= ll.N &&
                     ll.ln_w == nullptr && !(ll.flags & (SBEV_DENSE_RELU | SBEV_DENSE_REFINE)) && ll.residual == nullptr,
                     SBEV_ERR_INVALID, "sbev_dense_chain_points_fwd: bad sample-points arguments (plain Linear last layer holding the offsets and logits)");
        // fused only in the default tensor-core kernel; every other variant runs the chain, then sample_points_kernel
        bool mma_ok = use_mma && get_option(OPT_DENSE_NSPLIT) == 0 && get_option(OPT_DENSE_CLUSTER) <= 1 && get_option(OPT_DENSE_FUSE_POINTS);
        for (int i = 0; i < n_layers; ++i)
            if (layers[i].W_hi == nullptr || layers[i].W_lo == nullptr || layers[i].K > 512 || layers[i].N > 1024) mma_ok = false;
        if (mma_ok) {
            fuse_points = true;
            ChainLayer& c = prm.layer[n_layers - 1];
            c.flags = (c.flags & ~CHAIN_FLAG_VEC4) | CHAIN_FLAG_POINTS;          // scalar epilogue: the finished row stays in shared memory
            prm.sp_bbox = pts->bbox; prm.sp_points = pts->points; prm.sp_scale_w = pts->scale_w;
            for (int i = 0; i < 6; ++i) prm.sp_r[i] = pts->pc_range[i];
            prm.sp_GP = pts->GP; prm.sp_L = pts->L; prm.sp_off_col = pts->off_col; prm.sp_log_col = pts->log_col;
        }
    }
    if (pts != nullptr && !fuse_points) {
        int rc = dense_chain_impl(x, ldx, in, M, n_layers, layers, refine_proposal, refine_time_diff, refine_Q, refine_T, stream, nullptr);
        if (rc) return rc;
        const sbev_dense_layer& ll = layers[n_layers - 1];
        return sbev_sample_points_fwd(pts->bbox, ll.y + pts->off_col, ll.ldy, ll.y + pts->log_col, ll.ldy, pts->pc_range, M, pts->GP, pts->L,
                                      pts->points, pts->scale_w, stream);
    }
    if (use_mma) {
        // tensor-core path: pre-split bf16 weights [N][Kpad] streamed by TMA (box 64 k x 128 rows, 128-byte swizzle)
        ChainMaps maps;
        const int groups = (M + DENSE_ROWS - 1) / DENSE_ROWS;
        int cl = get_option(OPT_DENSE_CLUSTER);            // 0/1 = no cluster, 2 / 4 / 8 = CTAs per multicast cluster
        if (cl != 2 && cl != 4 && cl != 8) cl = 1;
        if (groups < cl) cl = 1;
        const int ns = get_option(OPT_DENSE_NSPLIT);        // 2 / 4 = N-split cluster chain (dense_chain_ns_kernel)
        bool ns_ok = ns == 2 || ns == 4;
        for (int i = 0; i < n_layers && ns_ok; ++i) {
            const sbev_dense_layer& l = layers[i];
            const bool slice = i + 1 == n_layers && l.ln_w == nullptr;
            if (l.K > 512 || l.N > 1024) ns_ok = false;
            if (!slice) {
                if (l.N > 512 || (l.N & 3) || (l.flags & SBEV_DENSE_REFINE)) ns_ok = false;
                if ((reinterpret_cast<uintptr_t>(l.ln_w) | reinterpret_cast<uintptr_t>(l.ln_b) | reinterpret_cast<uintptr_t>(l.residual)) & 15) ns_ok = false;
            }
        }
        if (ns_ok) {
            const int box = 256 / ns, rows = 8 * ns;
            for (int i = 0; i < n_layers; ++i) {
                const sbev_dense_layer& l = layers[i];
                SBEV_REQUIRE(l.Kpad >= l.K && (l.Kpad & 63) == 0, SBEV_ERR_INVALID, "sbev_dense_chain_fwd: layer %d: Kpad must be a multiple of 64 >= K", i);
                int rc = make_bf16_map(&maps.hi[i], l.W_hi, l.N, l.Kpad, box);
                if (rc) return rc;
                rc = make_bf16_map(&maps.lo[i], l.W_lo, l.N, l.Kpad, box);
                if (rc) return rc;
            }
            SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(dense_chain_ns_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NsCfg<2>::SMEM);
                cudaFuncSetAttribute(dense_chain_ns_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NsCfg<4>::SMEM));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((M + rows - 1) / rows * ns);
            cfg.blockDim = dim3(288);
            cfg.dynamicSmemBytes = ns == 2 ? NsCfg<2>::SMEM : NsCfg<4>::SMEM;
            cfg.stream = (cudaStream_t)stream;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = ns; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = get_option(OPT_PDL) ? 2 : 1;
            cudaError_t e = ns == 2 ? cudaLaunchKernelEx(&cfg, dense_chain_ns_kernel<2>, prm, maps)
                                    : cudaLaunchKernelEx(&cfg, dense_chain_ns_kernel<4>, prm, maps);
            if (e != cudaSuccess) { set_error("sbev_dense_chain_fwd(n-split cluster launch): %s", cudaGetErrorString(e)); return SBEV_ERR_CUDA; }
            return check_launch("sbev_dense_chain_fwd(n-split)");
        }
        const int box_rows = 128 / cl;
        for (int i = 0; i < n_layers; ++i) {
            const sbev_dense_layer& l = layers[i];
            SBEV_REQUIRE(l.Kpad >= l.K && (l.Kpad & 63) == 0, SBEV_ERR_INVALID, "sbev_dense_chain_fwd: layer %d: Kpad must be a multiple of 64 >= K", i);
            int rc = make_bf16_map(&maps.hi[i], l.W_hi, l.N, l.Kpad, box_rows);
            if (rc) return rc;
            rc = make_bf16_map(&maps.lo[i], l.W_lo, l.N, l.Kpad, box_rows);
            if (rc) return rc;
        }
        const size_t smem_mma = (size_t)MC_STAGES * 2 * MC_TILE_BYTES + (size_t)2 * 2 * DENSE_ROWS * MC_XLD * 2 + (size_t)DENSE_ROWS * MC_YLD * 4 +
                                (size_t)(3 * CHAIN_VEC_LD + DENSE_ROWS * MC_RES_LD) * 4 + 1024;
        SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(dense_chain_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma);
            cudaFuncSetAttribute(dense_chain_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma);
            cudaFuncSetAttribute(dense_chain_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma);
            cudaFuncSetAttribute(dense_chain_mma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma));
        // 16 rows per CTA (caller's hint on the first layer: the chain runs side by side with another one): narrow chains only
        bool wide_cta = (layers[0].flags & SBEV_DENSE_WIDE_CTA) && cl == 1 && in == nullptr && !fuse_points && M > DENSE_ROWS;
        for (int i = 0; i < n_layers; ++i)
            if (layers[i].K > 256 || layers[i].N > 256) wide_cta = false;
        if (wide_cta) {
            constexpr int R16 = 16, XLD16 = 256 + 8, YLD16 = 256 + 4;
            const size_t smem16 = (size_t)MC_STAGES * 2 * MC_TILE_BYTES + (size_t)2 * 2 * R16 * XLD16 * 2 + (size_t)R16 * YLD16 * 4 +
                                  (size_t)(3 * CHAIN_VEC_LD + R16 * MC_RES_LD) * 4 + 1024;
            SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(dense_chain_mma_kernel<1, R16, XLD16, YLD16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
            launch_pdl(dense_chain_mma_kernel<1, R16, XLD16, YLD16>, dim3((M + R16 - 1) / R16), dim3(288), smem16, (cudaStream_t)stream, prm, maps);
            return check_launch("sbev_dense_chain_fwd(mma, 16 rows)");
        }
        if (cl > 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((groups + cl - 1) / cl * cl);
            cfg.blockDim = dim3(288);
            cfg.dynamicSmemBytes = smem_mma;
            cfg.stream = (cudaStream_t)stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cudaError_t e = cl == 2 ? cudaLaunchKernelEx(&cfg, dense_chain_mma_kernel<2>, prm, maps)
                          : cl == 4 ? cudaLaunchKernelEx(&cfg, dense_chain_mma_kernel<4>, prm, maps)
                                    : cudaLaunchKernelEx(&cfg, dense_chain_mma_kernel<8>, prm, maps);
            if (e != cudaSuccess) { set_error("sbev_dense_chain_fwd(cluster launch): %s", cudaGetErrorString(e)); return SBEV_ERR_CUDA; }
        } else {
            launch_pdl(dense_chain_mma_kernel<1>, dim3(groups), dim3(288), smem_mma, (cudaStream_t)stream, prm, maps);
        }
        return check_launch("sbev_dense_chain_fwd(mma)");
    }
    SBEV_REQUIRE(in == nullptr, SBEV_ERR_UNSUPPORTED, "sbev_dense_chain_reduce_fwd: the fused input stage needs the tensor-core chain");
    for (int i = 0; i < n_layers; ++i)
        SBEV_REQUIRE(layers[i].Wt != nullptr, SBEV_ERR_INVALID, "sbev_dense_chain_fwd: layer %d: fp32 weight missing for the FFMA path", i);
    const size_t smem = sizeof(float) * ((size_t)CHAIN_STAGES * CHAIN_STAGE_FLOATS + 2 * (size_t)DENSE_ROWS * prm.act_ld);
    SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(dense_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dense_chain_kernel<<<(M + DENSE_ROWS - 1) / DENSE_ROWS, 256, smem, (cudaStream_t)stream>>>(prm);
    return check_launch("sbev_dense_chain_fwd");
}

extern "C" int sbev_dense_chain_fwd(const float* x, int ldx, int M, int n_layers, const sbev_dense_layer* layers,
                                    const float* refine_proposal, const float* refine_time_diff, int refine_Q, int refine_T,
                                    void* stream) {
    SBEV_REQUIRE(x != nullptr, SBEV_ERR_INVALID, "sbev_dense_chain_fwd: null pointer");
    return dense_chain_impl(x, ldx, nullptr, M, n_layers, layers, refine_proposal, refine_time_diff, refine_Q, refine_T, stream);
}

extern "C" int sbev_dense_chain_points_fwd(const float* x, int ldx, int M, int n_layers, const sbev_dense_layer* layers,
                                           const float* query_bbox, const float* pc_range, int GP, int L, int off_col, int log_col,
                                           float* points, float* scale_w, void* stream) {
    SBEV_REQUIRE(x != nullptr, SBEV_ERR_INVALID, "sbev_dense_chain_points_fwd: null pointer");
    const ChainPoints pts{query_bbox, pc_range, GP, L, off_col, log_col, points, scale_w};
    return dense_chain_impl(x, ldx, nullptr, M, n_layers, layers, nullptr, nullptr, 0, 0, stream, &pts);
}

extern "C" int sbev_reduce_ln_fwd(const float* partial, int nsplit, const float* bias, const float* residual,
                                  const float* ln_w, const float* ln_b, int M, int N, float* out, void* stream);

extern "C" int sbev_dense_chain_reduce_fwd(const float* partial, int nsplit, const float* bias, const float* residual,
                                           const float* ln_w, const float* ln_b, float* x_out,
                                           int M, int n_layers, const sbev_dense_layer* layers,
                                           const float* refine_proposal, const float* refine_time_diff, int refine_Q, int refine_T,
                                           void* stream) {
    SBEV_REQUIRE(partial && x_out && layers && nsplit >= 1, SBEV_ERR_INVALID, "sbev_dense_chain_reduce_fwd: bad arguments");
    SBEV_REQUIRE((ln_w == nullptr) == (ln_b == nullptr), SBEV_ERR_INVALID, "sbev_dense_chain_reduce_fwd: ln_w and ln_b go together");
    SBEV_REQUIRE(n_layers >= 1 && n_layers <= CHAIN_MAX_LAYERS, SBEV_ERR_UNSUPPORTED, "sbev_dense_chain_reduce_fwd: 1..%d layers", CHAIN_MAX_LAYERS);
    const int K0 = layers[0].K;
    bool fusable = get_option(OPT_DENSE_IMPL) == 0 && (K0 == 128 || K0 == 256);
    for (int i = 0; i < n_layers; ++i)
        if (layers[i].W_hi == nullptr || layers[i].W_lo == nullptr || layers[i].K > 512 || layers[i].N > 1024) fusable = false;
    const void* al[5] = {partial, bias, residual, ln_w, ln_b};
    for (int i = 0; i < 5; ++i) SBEV_REQUIRE((reinterpret_cast<uintptr_t>(al[i]) & 15) == 0, SBEV_ERR_INVALID, "sbev_dense_chain_reduce_fwd: operands must be 16-byte aligned");
    SBEV_REQUIRE((reinterpret_cast<uintptr_t>(x_out) & 15) == 0, SBEV_ERR_INVALID, "sbev_dense_chain_reduce_fwd: x_out must be 16-byte aligned");
    if (!fusable) {                              // same result in two launches (fp32 FFMA chain, or an input width the fused stage does not cover)
        int rc = sbev_reduce_ln_fwd(partial, nsplit, bias, residual, ln_w, ln_b, M, K0, x_out, stream);
        if (rc) return rc;
        return dense_chain_impl(x_out, K0, nullptr, M, n_layers, layers, refine_proposal, refine_time_diff, refine_Q, refine_T, stream);
    }
    const ChainInputReduce in{partial, nsplit, bias, residual, ln_w, ln_b, x_out};
    return dense_chain_impl(nullptr, K0, &in, M, n_layers, layers, refine_proposal, refine_time_diff, refine_Q, refine_T, stream);
}

extern "C" int sbev_dense_fwd(const float* x, int ldx, const float* Wt, int ldw, const float* bias,
                              const float* ln_w, const float* ln_b, const float* residual,
                              int M, int K, int N, int flags, float* y, void* stream) {
    SBEV_REQUIRE(y != nullptr, SBEV_ERR_INVALID, "sbev_dense_fwd: null pointer");
    sbev_dense_layer l{};
    l.Wt = Wt; l.ldw = ldw; l.K = K; l.N = N; l.bias = bias; l.ln_w = ln_w; l.ln_b = ln_b; l.residual = residual;
    l.flags = flags & (SBEV_DENSE_RELU | SBEV_DENSE_RES_PRE_LN); l.y = y; l.ldy = N;
    l.W_hi = nullptr; l.W_lo = nullptr; l.Kpad = 0; l.y_hi = nullptr; l.y_lo = nullptr;
    return sbev_dense_chain_fwd(x, ldx, M, 1, &l, nullptr, nullptr, 0, 0, stream);
}

extern "C" int sbev_sample_points_fwd(const float* query_bbox, const float* offset, int ld_off, const float* scale_logits, int ld_log,
                                      const float* pc_range, int BQ, int GP, int L,
                                      float* points, float* scale_w, void* stream) {
    SBEV_REQUIRE(query_bbox && offset && scale_logits && pc_range && points && scale_w, SBEV_ERR_INVALID,
                 "sbev_sample_points_fwd: null pointer");
    SBEV_REQUIRE(BQ >= 0 && GP > 0 && L >= 1 && L <= SBEV_MAX_LEVELS && ld_off >= GP * 3 && ld_log >= GP * L, SBEV_ERR_INVALID, "sbev_sample_points_fwd: bad sizes");
    const long long total = (long long)BQ * GP;
    if (total == 0) return SBEV_OK;
    launch_pdl(sample_points_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, (cudaStream_t)stream,
        query_bbox, offset, ld_off, scale_logits, ld_log, pc_range[0], pc_range[1], pc_range[2], pc_range[3], pc_range[4], pc_range[5],
        BQ, GP, L, points, scale_w, get_option(OPT_LEGACY_ROTATION) ? -1.f : 1.f);
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
    SBEV_REQUIRE(N > 0 && N <= 1024 && (N & 3) == 0, SBEV_ERR_UNSUPPORTED, "sbev_reduce_ln_fwd: N must be a multiple of 4 in (0,1024]");
    SBEV_REQUIRE((ln_w == nullptr) == (ln_b == nullptr), SBEV_ERR_INVALID, "sbev_reduce_ln_fwd: ln_w and ln_b go together");
    if (M <= 0) return SBEV_OK;
    launch_pdl(reduce_ln_kernel, dim3(M), dim3(256), 0, (cudaStream_t)stream, partial, nsplit, bias, residual, ln_w, ln_b, M, N, out);
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
