// Weights-stationary dense chain (sbev_dense_chain_ws_fwd / sbev_dense_chain_ws_reduce_fwd): the same Linear (+bias) (+residual)
// (+LayerNorm) (+ReLU) chains as dense.cu, re-partitioned so that no CTA ever streams a whole chain's weights.
//
// Why: dense_chain_mma_kernel gives every CTA 8 rows and makes it stream ALL weights of the chain (0.4-1.1 MB of bf16 hi + lo);
// an SM ingests ~64 B/clk (~120 GB/s), so a chain costs 9-20 us NO MATTER how few rows there are (measured identical from 113
// to 900 rows, tests/perf/kernel_sweep.py).  Here a CLUSTER of 8 CTAs owns a block of rows and SPLITS EVERY LAYER'S OUTPUT
// FEATURES 8 ways: CTA c keeps features [c*SW, (c+1)*SW) of every layer resident in shared memory for the whole kernel -- one
// bulk copy of <= 144 KB per CTA, issued before griddepcontrol.wait, i.e. under the previous kernel's tail -- and row tiles of RT
// (16 or 32) rows stream through:
//     MMA (RT rows x this CTA's feature slice; rows on the MMA M dimension, bf16x3)
//     ->  + bias, 8-byte st.shared::cluster of every (row, feature pair) into the y tile of ALL 8 CTAs
//     ->  barrier.cluster (A)  ->  every CTA finishes the full rows (pre-LN residual, LayerNorm, ReLU, post residual) and writes the
//         next layer's bf16 (hi, lo) operand locally; row r's global outputs are written by CTA r % 8
//     ->  barrier.cluster arrive (B); its wait sits in front of the next layer's remote stores (split-phase: normally long satisfied)
// A last layer without LayerNorm needs no full rows: every CTA finishes and stores its own features (in-projection + tau,
// sampling heads, cls / reg outputs, box refinement).
// The split-K reduce + out_proj bias + identity + norm2 in front of the FFN (sbev_dense_chain_reduce_fwd's prologue) is distributed
// the same way: CTA c reduces tile rows r % 8 == c, stores the fp32 row (it is a later layer's residual) and broadcasts its bf16
// (hi, lo) operand to the 8 CTAs through distributed shared memory.
// Behavioural reference: the same lines dense.cu cites (/root/reference/models/sparsebev_transformer.py:113-183).
#include "common.cuh"
#include <cuda_bf16.h>
#include <mutex>
#include <unordered_map>

namespace sbev {

constexpr int WS_CL = 8;                 // CTAs per cluster (portable maximum)
constexpr int WS_MAX_LAYERS = 6;
constexpr int WS_THREADS = 256;
constexpr int WS_RED_K = 256;            // row width the reduce prologue handles

struct WsLayer {
    const float* bias; const float* ln_w; const float* ln_b; const float* residual; float* y;
    __nv_bfloat16* y_hi; __nv_bfloat16* y_lo;
    int K, Kpad, N, SW, flags, ldy, woff, vec;     // SW = features per CTA (multiple of 8); woff = byte offset of this layer inside a CTA's weight
                                                   // blob; vec = slice-mode outputs may be stored as aligned pairs
};
struct WsParams {
    const float* x; int ldx, M, n_layers, RT, rows_per_cluster, xld, yld, x_vec4;
    const float* red_partial; const float* red_bias; const float* red_res; const float* red_ln_w; const float* red_ln_b; float* red_out;
    int red_nsplit;
    const float* aux_proposal; const float* aux_time_diff; int aux_Q, aux_T;
    const uint8_t* blob; long long blob_stride; int blob_bytes;
    WsLayer layer[WS_MAX_LAYERS];
};

__device__ __forceinline__ uint32_t ws_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ws_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void ws_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void ws_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t ws_mapa(uint32_t addr, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(rank));
    return remote;
}
__device__ __forceinline__ void ws_st_remote_f2(uint32_t addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void ws_st_remote_u2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void ws_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void ws_ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ws_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float4 ws_ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// refine_bbox + velocity rescale on one output element (sparsebev_transformer.py:155-160,179-183); same arithmetic as dense.cu
__device__ __forceinline__ float ws_refine(const WsParams& prm, int row, int n, int N, float v) {
    if (n < 3) {
        const float x = fminf(fmaxf(__ldg(prm.aux_proposal + (long long)row * N + n), 0.f), 1.f);
        v = v + logf(__fdiv_rn(fmaxf(x, 1e-5f), fmaxf(1.f - x, 1e-5f)));
        v = __fdiv_rn(1.f, 1.f + expf(-v));
    } else if (n >= 8 && prm.aux_T > 1) {
        float td = __ldg(prm.aux_time_diff + (row / prm.aux_Q) * prm.aux_T + 1);
        if (td < 1e-5f) td = 1.0f;
        v = __fdiv_rn(v, td);
    }
    return v;
}

// slice mode: finish one output element of a last layer without LayerNorm
__device__ __forceinline__ float ws_finish(const WsParams& prm, const WsLayer& L, int row, int f, float v) {
    const bool pre_res = (L.flags & SBEV_DENSE_RES_PRE_LN) != 0;
    if (pre_res && L.residual != nullptr) v += __ldcg(L.residual + (long long)row * L.N + f);
    if (L.flags & SBEV_DENSE_RELU) v = fmaxf(v, 0.f);
    if (!pre_res && L.residual != nullptr) v += __ldcg(L.residual + (long long)row * L.N + f);
    if (L.flags & SBEV_DENSE_REFINE) v = ws_refine(prm, row, f, L.N, v);
    return v;
}

__global__ void __launch_bounds__(WS_THREADS, 1)
dense_chain_ws_kernel(const __grid_constant__ WsParams prm) {
    extern __shared__ uint8_t ws_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int RT = prm.RT, XLD = prm.xld, YLD = prm.yld;
    uint8_t* wsm = smem;                                                                   // this CTA's weight slices of every layer
    __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(smem + ((prm.blob_bytes + 1023) & ~1023));
    __nv_bfloat16* xl = xh + RT * XLD;
    float* ys = reinterpret_cast<float*>(xl + RT * XLD);                                   // [RT][YLD]: full rows of the current layer's output
    __shared__ uint64_t wbar;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = ws_ctarank();
    const int cluster = blockIdx.x / WS_CL;
    const int row_lo = cluster * prm.rows_per_cluster;
    const int row_hi = min(prm.M, row_lo + prm.rows_per_cluster);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_u32(&wbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // the weights are constants of the model: no preceding kernel writes them, so their copy may start before griddepcontrol.wait
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ws_u32(&wbar)), "r"(prm.blob_bytes) : "memory");
        const uint8_t* src = prm.blob + (long long)crank * prm.blob_stride;
        for (int off = 0; off < prm.blob_bytes; off += 32768) {
            const int n = min(32768, prm.blob_bytes - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(ws_u32(wsm + off)), "l"(src + off), "r"(n), "r"(ws_u32(&wbar)) : "memory");
        }
    }
    pdl_wait();
    pdl_trigger();
    // every CTA of the cluster is running (its shared memory may be written remotely from now on)
    ws_cluster_arrive();
    ws_cluster_wait();

    const int g8 = lane >> 2, t4 = lane & 3;
    const uint32_t ys_u32 = ws_u32(ys);
    uint32_t ys_peer[WS_CL];
#pragma unroll
    for (int p = 0; p < WS_CL; ++p) ys_peer[p] = ws_mapa(ys_u32, p);
    // ldmatrix lane roles.  A (rows x k, from the operand tile): matrix (lane >> 3) = (rows +0 / +8, k +0 / +8);
    // B (features x k, from the weight slice): matrix (lane >> 3) = k chunk (lane >> 3) of a 32-wide k group
    const int a_row = (lane & 7) + 8 * ((lane >> 3) & 1), a_kofs = 8 * (lane >> 4);
    const int b_row = lane & 7, b_sel = lane >> 3;
    const bool reduce_mode = prm.red_partial != nullptr;
    const bool last_exchange = prm.layer[prm.n_layers - 1].ln_w != nullptr;
    bool weights_ready = false, pending_b = false;

    for (int r0 = row_lo, tile = 0; r0 < row_hi; r0 += RT, ++tile) {
        if (reduce_mode) {
            if (pending_b) { ws_cluster_wait(); pending_b = false; }
            // the staging below writes the peers' operand tiles: their MMAs of the previous tile's last layer must be done
            if (tile > 0 && !last_exchange) { ws_cluster_arrive(); ws_cluster_wait(); }
        }
        __syncthreads();                        // this CTA's warps are done with the previous tile's operand / y rows
        if (reduce_mode) {
            // ---- input rows = LayerNorm(sum_z partial[z][row] + bias + residual[row]); CTA c owns tile rows r = 8 j + c
            const int nrows = RT >> 3, wpr = WS_CL / nrows;               // owned rows; warps per owned row (4 or 2)
            const int j = warp / wpr, part = warp - j * wpr;
            const int r = 8 * j + (int)crank, row = r0 + r;
            const bool live = row < row_hi;
            float* scratch = ys;                                          // [8 warps][256] (peers write ys only after the barrier below)
            float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
            const long long zs = (long long)prm.M * WS_RED_K;
            const float* base = prm.red_partial + (long long)row * WS_RED_K + 4 * lane;
            for (int zb = part; zb < prm.red_nsplit; zb += 6 * wpr) {
                float4 t[2][6];
#pragma unroll
                for (int u = 0; u < 6; ++u)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int z = zb + u * wpr;
                        t[i][u] = (live && z < prm.red_nsplit) ? ldg4(base + 128 * i + (long long)z * zs) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                for (int u = 0; u < 6; ++u)
#pragma unroll
                    for (int i = 0; i < 2; ++i) { acc[i].x += t[i][u].x; acc[i].y += t[i][u].y; acc[i].z += t[i][u].z; acc[i].w += t[i][u].w; }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(scratch + warp * WS_RED_K + 4 * lane + 128 * i) = acc[i];
            __syncthreads();
            if (part == 0) {
                float4 a[2];
                float sum = 0.f;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int n = 4 * lane + 128 * i;
                    a[i] = *reinterpret_cast<const float4*>(scratch + warp * WS_RED_K + n);
                    for (int p = 1; p < wpr; ++p) {
                        const float4 q = *reinterpret_cast<const float4*>(scratch + (warp + p) * WS_RED_K + n);
                        a[i].x += q.x; a[i].y += q.y; a[i].z += q.z; a[i].w += q.w;
                    }
                    if (prm.red_bias) { const float4 b = ldg4(prm.red_bias + n); a[i].x += b.x; a[i].y += b.y; a[i].z += b.z; a[i].w += b.w; }
                    if (prm.red_res && live) { const float4 q = ldg4(prm.red_res + (long long)row * WS_RED_K + n); a[i].x += q.x; a[i].y += q.y; a[i].z += q.z; a[i].w += q.w; }
                    sum += (a[i].x + a[i].y) + (a[i].z + a[i].w);
                }
                if (prm.red_ln_w != nullptr) {
                    const float mean = warp_sum(sum) / (float)WS_RED_K;
                    float ss = 0.f;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float p0 = a[i].x - mean, p1 = a[i].y - mean, p2 = a[i].z - mean, p3 = a[i].w - mean;
                        ss += (p0 * p0 + p1 * p1) + (p2 * p2 + p3 * p3);
                    }
                    const float rstd = rsqrtf(warp_sum(ss) / (float)WS_RED_K + 1e-5f);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int n = 4 * lane + 128 * i;
                        const float4 g = ldg4(prm.red_ln_w + n), b = ldg4(prm.red_ln_b + n);
                        a[i].x = (a[i].x - mean) * rstd * g.x + b.x; a[i].y = (a[i].y - mean) * rstd * g.y + b.y;
                        a[i].z = (a[i].z - mean) * rstd * g.z + b.z; a[i].w = (a[i].w - mean) * rstd * g.w + b.w;
                    }
                }
                const uint32_t xh_u32 = ws_u32(xh), xl_u32 = ws_u32(xl);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int n = 4 * lane + 128 * i;
                    const float4 v = live ? a[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live && prm.red_out) *reinterpret_cast<float4*>(prm.red_out + (long long)row * WS_RED_K + n) = v;
                    uint32_t h0, l0, h1, l1;
                    ws_split2(v.x, v.y, h0, l0); ws_split2(v.z, v.w, h1, l1);
                    const uint32_t off = (uint32_t)((r * XLD + n) * 2);
#pragma unroll
                    for (int p = 0; p < WS_CL; ++p) {
                        ws_st_remote_u2(ws_mapa(xh_u32 + off, p), h0, h1);
                        ws_st_remote_u2(ws_mapa(xl_u32 + off, p), l0, l1);
                    }
                }
            }
            // all 8 CTAs' rows have landed everywhere; the fp32 rows in global memory are visible to the cluster
            ws_cluster_arrive();
            ws_cluster_wait();
        } else {
            // ---- stage the tile's input rows as bf16 (hi, lo), zero-padded to the first layer's padded K; rows beyond row_hi are zero
            const int K0 = prm.layer[0].K, K0p = prm.layer[0].Kpad;
            if (prm.x_vec4) {
                const int k4 = K0p >> 2;
                for (int i = tid; i < RT * k4; i += WS_THREADS) {
                    const int r = i / k4, k = (i - r * k4) * 4;
                    const float4 v = (r0 + r < row_hi && k < K0) ? ldg4(prm.x + (long long)(r0 + r) * prm.ldx + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                    uint32_t h0, l0, h1, l1;
                    ws_split2(v.x, v.y, h0, l0); ws_split2(v.z, v.w, h1, l1);
                    *reinterpret_cast<uint2*>(xh + r * XLD + k) = make_uint2(h0, h1);
                    *reinterpret_cast<uint2*>(xl + r * XLD + k) = make_uint2(l0, l1);
                }
            } else {
                for (int i = tid; i < RT * (K0p / 2); i += WS_THREADS) {
                    const int r = i / (K0p / 2), k = (i - r * (K0p / 2)) * 2;
                    const bool ok = r0 + r < row_hi;
                    const float a = (ok && k < K0) ? __ldg(prm.x + (long long)(r0 + r) * prm.ldx + k) : 0.f;
                    const float b = (ok && k + 1 < K0) ? __ldg(prm.x + (long long)(r0 + r) * prm.ldx + k + 1) : 0.f;
                    uint32_t h, l;
                    ws_split2(a, b, h, l);
                    *reinterpret_cast<uint32_t*>(xh + r * XLD + k) = h;
                    *reinterpret_cast<uint32_t*>(xl + r * XLD + k) = l;
                }
            }
            __syncthreads();
        }

        for (int li = 0; li < prm.n_layers; ++li) {
            const WsLayer& L = prm.layer[li];
            const int N = L.N, SW = L.SW;
            const bool last = li + 1 == prm.n_layers;
            const bool exchange = !(last && L.ln_w == nullptr);
            const bool pre_res = (L.flags & SBEV_DENSE_RES_PRE_LN) && L.residual != nullptr;
            // residual rows of the rows this warp will finish -> registers while the MMAs run (N <= 256; wider ones are read in place)
            const bool res_reg = exchange && L.residual != nullptr && N <= 256;
            float4 resr[4][2];
            if (res_reg) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int r = warp + 8 * jj, n = 4 * lane + 128 * i;
                        resr[jj][i] = (r < RT && r0 + r < row_hi && n < N) ? ws_ldcg4(L.residual + (long long)(r0 + r) * N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
            }
            if (!weights_ready) {
                uint32_t done = 0;
                for (uint32_t spins = 0; !done; ++spins) {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(ws_u32(&wbar)), "r"(0) : "memory");
                    if (!done && spins > (1u << 26)) __trap();
                }
                weights_ready = true;
            }
            // ---- MMA: (m16 row tile, n8 feature tile) pairs of this CTA's slice, round-robin over the 8 warps
            const int ftiles = SW >> 3, rtiles = RT >> 4, kchunks = L.Kpad >> 6;
            const uint32_t wl_base = ws_u32(wsm + L.woff);
            const int tile_bytes = SW * 128;                                     // one (hi or lo) SW x 64 bf16 tile
            bool waited_b = false;
            for (int t = warp; t < ftiles * rtiles; t += 8) {
                const int ft = t % ftiles, rt = t / ftiles;
                const int f0 = (int)crank * SW + 8 * ft;
                if (f0 >= N) continue;                                           // (warp-uniform) a slice beyond the layer's features
                const uint32_t xh_addr = ws_u32(xh + (16 * rt + a_row) * XLD + a_kofs);
                const uint32_t xl_addr = ws_u32(xl + (16 * rt + a_row) * XLD + a_kofs);
                const uint32_t w_row = wl_base + (uint32_t)((8 * ft + b_row) * 128);
                float accA[4] = {0.f, 0.f, 0.f, 0.f}, accB[4] = {0.f, 0.f, 0.f, 0.f}, accC[4] = {0.f, 0.f, 0.f, 0.f};
                float accD[4] = {0.f, 0.f, 0.f, 0.f}, accE[4] = {0.f, 0.f, 0.f, 0.f}, accF[4] = {0.f, 0.f, 0.f, 0.f};
                for (int kc = 0; kc < kchunks; ++kc) {
                    const uint32_t wh = w_row + (uint32_t)(kc * 2 * tile_bytes);
                    const uint32_t wlo = wh + tile_bytes;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        uint32_t bh[4], bl[4];
                        const uint32_t sw = (uint32_t)(((4 * kk + b_sel) ^ b_row) << 4);
                        ws_ldsm_x4(bh, wh + sw);
                        ws_ldsm_x4(bl, wlo + sw);
                        uint32_t ah[4], al[4];
                        const uint32_t xo = (uint32_t)((kc * 64 + kk * 32) * 2);
                        ws_ldsm_x4(ah, xh_addr + xo);
                        ws_ldsm_x4(al, xl_addr + xo);
                        ws_mma(accA, ah, bh[0], bh[1]); ws_mma(accC, al, bh[0], bh[1]); ws_mma(accE, ah, bl[0], bl[1]);
                        ws_ldsm_x4(ah, xh_addr + xo + 32);
                        ws_ldsm_x4(al, xl_addr + xo + 32);
                        ws_mma(accB, ah, bh[2], bh[3]); ws_mma(accD, al, bh[2], bh[3]); ws_mma(accF, ah, bl[2], bl[3]);
                    }
                }
                // D fragment: rows 16 rt + g8 (+8), features f0 + 2 t4 (+1)
                float v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = ((accC[i] + accD[i]) + (accE[i] + accF[i])) + (accA[i] + accB[i]);
                const int f = f0 + 2 * t4;
                if (L.bias != nullptr) {
                    const float b0 = f < N ? __ldg(L.bias + f) : 0.f, b1 = f + 1 < N ? __ldg(L.bias + f + 1) : 0.f;
                    v[0] += b0; v[1] += b1; v[2] += b0; v[3] += b1;
                }
                if (exchange) {
                    if (pending_b && !waited_b) { ws_cluster_wait(); waited_b = true; }      // peers are done reading the previous y tile
                    if (f < N) {                                                             // (N % 4 == 0: the pair is in or out)
                        const uint32_t o0 = (uint32_t)(((16 * rt + g8) * YLD + f) * 4), o1 = o0 + (uint32_t)(8 * YLD * 4);
#pragma unroll
                        for (int p = 0; p < WS_CL; ++p) {
                            ws_st_remote_f2(ys_peer[p] + o0, v[0], v[1]);
                            ws_st_remote_f2(ys_peer[p] + o1, v[2], v[3]);
                        }
                    }
                } else {
                    // slice mode: finish and store this CTA's features of the tile's rows
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int row = r0 + 16 * rt + g8 + 8 * h;
                        if (row >= row_hi) continue;
                        const long long yoff = (long long)row * L.ldy + f;
                        if (L.vec && f + 1 < N) {
                            const float a = ws_finish(prm, L, row, f, v[2 * h]), b = ws_finish(prm, L, row, f + 1, v[2 * h + 1]);
                            if (L.y != nullptr) *reinterpret_cast<float2*>(L.y + yoff) = make_float2(a, b);
                            if (L.y_hi != nullptr) {
                                uint32_t hh, ll;
                                ws_split2(a, b, hh, ll);
                                *reinterpret_cast<uint32_t*>(L.y_hi + yoff) = hh;
                                *reinterpret_cast<uint32_t*>(L.y_lo + yoff) = ll;
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                if (f + e >= N) continue;
                                const float a = ws_finish(prm, L, row, f + e, v[2 * h + e]);
                                if (L.y != nullptr) L.y[yoff + e] = a;
                                if (L.y_hi != nullptr) {
                                    const __nv_bfloat16 hb = __float2bfloat16_rn(a);
                                    L.y_hi[yoff + e] = hb;
                                    L.y_lo[yoff + e] = __float2bfloat16_rn(a - __bfloat162float(hb));
                                }
                            }
                        }
                    }
                }
            }
            if (!exchange) continue;                 // (slice mode is the last layer of the chain)
            if (pending_b && !waited_b) ws_cluster_wait();          // warps without an MMA tile still take part in every barrier phase
            pending_b = false;
            // ---- barrier A: every CTA's slice of the y tile has landed everywhere
            ws_cluster_arrive();
            ws_cluster_wait();
            // ---- full rows: warp w finishes rows w, w + 8, ...; the next layer's operand is written locally, global outputs by CTA (row % 8)
            const bool has_ln = L.ln_w != nullptr;
            const bool post_res = !pre_res && L.residual != nullptr;
            const int Kn = last ? 0 : prm.layer[li + 1].Kpad;
            const int per = (N + 127) >> 7;                                      // float4 per lane per row (N <= 512, N % 4 == 0)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int r = warp + 8 * jj;
                if (r >= RT) break;
                const int row = r0 + r;
                const bool live = row < row_hi;
                float4 v[4], q[4];
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n = 4 * lane + 128 * i;
                    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    q[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < per && n < N) {
                        v[i] = *reinterpret_cast<const float4*>(ys + r * YLD + n);
                        if (L.residual != nullptr && live) {
                            if (res_reg) { if (i < 2) q[i] = resr[jj][i]; }
                            else q[i] = ws_ldcg4(L.residual + (long long)row * N + n);
                        }
                        if (pre_res) { v[i].x += q[i].x; v[i].y += q[i].y; v[i].z += q[i].z; v[i].w += q[i].w; }
                        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
                    }
                }
                float mean = 0.f, rstd = 1.f;
                if (has_ln) {
                    mean = warp_sum(s) / (float)N;
                    float ss = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int n = 4 * lane + 128 * i;
                        if (i < per && n < N) {
                            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
                            ss += (a * a + b * b) + (c * c + d * d);
                        }
                    }
                    rstd = rsqrtf(warp_sum(ss) / (float)N + 1e-5f);
                }
                const bool owner = live && (r & (WS_CL - 1)) == (int)crank;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n = 4 * lane + 128 * i;
                    if (i < per && n < N) {
                        float4 o = v[i];
                        if (has_ln) {
                            const float4 g = ldg4(L.ln_w + n), b = ldg4(L.ln_b + n);
                            o.x = (o.x - mean) * rstd * g.x + b.x; o.y = (o.y - mean) * rstd * g.y + b.y;
                            o.z = (o.z - mean) * rstd * g.z + b.z; o.w = (o.w - mean) * rstd * g.w + b.w;
                        }
                        if (L.flags & SBEV_DENSE_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        if (post_res) { o.x += q[i].x; o.y += q[i].y; o.z += q[i].z; o.w += q[i].w; }
                        if (!live) o = make_float4(0.f, 0.f, 0.f, 0.f);
                        uint32_t h0, l0, h1, l1;
                        ws_split2(o.x, o.y, h0, l0); ws_split2(o.z, o.w, h1, l1);
                        if (owner) {
                            const long long yoff = (long long)row * L.ldy + n;
                            if (L.y != nullptr) *reinterpret_cast<float4*>(L.y + yoff) = o;
                            if (L.y_hi != nullptr) {
                                *reinterpret_cast<uint2*>(L.y_hi + yoff) = make_uint2(h0, h1);
                                *reinterpret_cast<uint2*>(L.y_lo + yoff) = make_uint2(l0, l1);
                            }
                        }
                        if (!last) {
                            *reinterpret_cast<uint2*>(xh + r * XLD + n) = make_uint2(h0, h1);
                            *reinterpret_cast<uint2*>(xl + r * XLD + n) = make_uint2(l0, l1);
                        }
                    }
                }
                if (!last)
                    for (int k = N + 4 * lane; k < Kn; k += 128) {               // zero beyond N up to the next layer's padded K
                        *reinterpret_cast<uint2*>(xh + r * XLD + k) = make_uint2(0u, 0u);
                        *reinterpret_cast<uint2*>(xl + r * XLD + k) = make_uint2(0u, 0u);
                    }
            }
            // ---- barrier B (arrive only): this CTA is done reading its y tile; the matching wait precedes the next remote stores
            ws_cluster_arrive();
            pending_b = true;
            __syncthreads();                         // the next layer's operand rows are complete for all warps of this CTA
        }
    }
    if (pending_b) ws_cluster_wait();
    // no CTA may exit while a peer could still store into its shared memory
    ws_cluster_arrive();
    ws_cluster_wait();
}

// Features per CTA of a layer: ceil(N / 8) rounded up to the MMA's 8-feature tile.
static inline int ws_slice_width(int N) { return ((N + WS_CL - 1) / WS_CL + 7) & ~7; }

// Bytes of one CTA's weight blob for a chain; `woff` (optional) receives the per-layer offsets.
static long long ws_blob_bytes(int n_layers, const sbev_dense_layer* layers, int* woff, int* sw_out) {
    long long off = 0;
    for (int i = 0; i < n_layers; ++i) {
        const int SW = ws_slice_width(layers[i].N);
        if (woff) woff[i] = (int)off;
        if (sw_out) sw_out[i] = SW;
        off += (long long)(layers[i].Kpad / 64) * 2 * SW * 128;
    }
    return off;
}

struct WsReduce {
    const float* partial; int nsplit; const float* bias; const float* residual; const float* ln_w; const float* ln_b; float* x_out;
};

static int ws_launch(const float* x, int ldx, const WsReduce* red, int M, int n_layers, const sbev_dense_layer* layers,
                     const uint16_t* blob, long long blob_stride_bytes,
                     const float* refine_proposal, const float* refine_time_diff, int refine_Q, int refine_T, void* stream, const char* who) {
    SBEV_REQUIRE(layers && blob && (red != nullptr || x != nullptr), SBEV_ERR_INVALID, "%s: null pointer", who);
    SBEV_REQUIRE(n_layers >= 1 && n_layers <= WS_MAX_LAYERS, SBEV_ERR_UNSUPPORTED, "%s: 1..%d layers", who, WS_MAX_LAYERS);
    SBEV_REQUIRE(M >= 0, SBEV_ERR_INVALID, "%s: bad sizes", who);
    SBEV_REQUIRE((reinterpret_cast<uintptr_t>(blob) & 127) == 0 && (blob_stride_bytes & 127) == 0, SBEV_ERR_INVALID, "%s: blob must be 128-byte aligned", who);
    if (M == 0) return SBEV_OK;
    WsParams prm{};
    int woff[WS_MAX_LAYERS], sw[WS_MAX_LAYERS];
    for (int i = 0; i < n_layers; ++i)
        SBEV_REQUIRE(layers[i].K > 0 && layers[i].N > 0 && layers[i].Kpad >= layers[i].K && (layers[i].Kpad & 63) == 0 && layers[i].Kpad <= 512,
                     SBEV_ERR_UNSUPPORTED, "%s: layer %d: K <= 512, Kpad %% 64 == 0", who, i);
    const long long bytes = ws_blob_bytes(n_layers, layers, woff, sw);
    SBEV_REQUIRE(blob_stride_bytes >= bytes, SBEV_ERR_INVALID, "%s: blob stride %lld smaller than the chain's %lld bytes per CTA", who, blob_stride_bytes, bytes);
    int kmax = 64, nex = 4;
    for (int i = 0; i < n_layers; ++i) {
        const sbev_dense_layer& l = layers[i];
        const bool last = i + 1 == n_layers;
        const bool exchange = !(last && l.ln_w == nullptr);
        SBEV_REQUIRE(i == 0 || l.K == layers[i - 1].N, SBEV_ERR_INVALID, "%s: layer %d K != previous N", who, i);
        SBEV_REQUIRE((l.ln_w == nullptr) == (l.ln_b == nullptr) && (l.y_hi == nullptr) == (l.y_lo == nullptr), SBEV_ERR_INVALID, "%s: ln_w/ln_b and y_hi/y_lo go together", who);
        SBEV_REQUIRE((l.y == nullptr && l.y_hi == nullptr) || l.ldy >= l.N, SBEV_ERR_INVALID, "%s: layer %d: ldy < N", who, i);
        const uintptr_t f32_ptrs = reinterpret_cast<uintptr_t>(l.ln_w) | reinterpret_cast<uintptr_t>(l.ln_b) | reinterpret_cast<uintptr_t>(l.residual) | reinterpret_cast<uintptr_t>(l.y);
        const uintptr_t b16_ptrs = reinterpret_cast<uintptr_t>(l.y_hi) | reinterpret_cast<uintptr_t>(l.y_lo);
        if (exchange) {
            SBEV_REQUIRE(l.N <= 512 && (l.N & 3) == 0 && !(l.flags & SBEV_DENSE_REFINE), SBEV_ERR_UNSUPPORTED,
                         "%s: layer %d: a layer that feeds another one (or ends in LayerNorm) needs N <= 512, N %% 4 == 0", who, i);
            SBEV_REQUIRE((f32_ptrs & 15) == 0 && (b16_ptrs & 7) == 0 && ((l.y == nullptr && l.y_hi == nullptr) || (l.ldy & 3) == 0),
                         SBEV_ERR_INVALID, "%s: layer %d: operands must be 16-byte aligned", who, i);
            nex = nex > l.N ? nex : l.N;
        }
        if (l.flags & SBEV_DENSE_REFINE)
            SBEV_REQUIRE(refine_proposal && refine_time_diff && l.N >= 10 && refine_T >= 1, SBEV_ERR_INVALID, "%s: refine needs proposal/time_diff", who);
        WsLayer& c = prm.layer[i];
        c.bias = l.bias; c.ln_w = l.ln_w; c.ln_b = l.ln_b; c.residual = l.residual; c.y = l.y;
        c.y_hi = reinterpret_cast<__nv_bfloat16*>(l.y_hi); c.y_lo = reinterpret_cast<__nv_bfloat16*>(l.y_lo);
        c.K = l.K; c.Kpad = l.Kpad; c.N = l.N; c.SW = sw[i]; c.flags = l.flags & 0xff; c.ldy = l.ldy; c.woff = woff[i];
        c.vec = ((l.ldy & 1) == 0 && (reinterpret_cast<uintptr_t>(l.y) & 7) == 0 && (b16_ptrs & 3) == 0) ? 1 : 0;
        kmax = kmax > l.Kpad ? kmax : l.Kpad;
    }
    SBEV_REQUIRE(layers[n_layers - 1].y != nullptr, SBEV_ERR_INVALID, "%s: last layer needs an output pointer", who);
    if (red != nullptr) {
        SBEV_REQUIRE(red->partial && red->nsplit >= 1 && red->x_out, SBEV_ERR_INVALID, "%s: partial / x_out missing", who);
        SBEV_REQUIRE(layers[0].K == WS_RED_K, SBEV_ERR_UNSUPPORTED, "%s: the reduce prologue handles rows of %d floats", who, WS_RED_K);
        SBEV_REQUIRE((red->ln_w == nullptr) == (red->ln_b == nullptr), SBEV_ERR_INVALID, "%s: ln_w/ln_b go together", who);
        SBEV_REQUIRE(((reinterpret_cast<uintptr_t>(red->partial) | reinterpret_cast<uintptr_t>(red->bias) | reinterpret_cast<uintptr_t>(red->residual) |
                       reinterpret_cast<uintptr_t>(red->ln_w) | reinterpret_cast<uintptr_t>(red->ln_b) | reinterpret_cast<uintptr_t>(red->x_out)) & 15) == 0,
                     SBEV_ERR_INVALID, "%s: reduce operands must be 16-byte aligned", who);
        nex = nex > 252 ? nex : 252;          // the y tile doubles as the reduce scratch: 8 x 256 floats <= 16 x (nex + 4)
        prm.red_partial = red->partial; prm.red_nsplit = red->nsplit; prm.red_bias = red->bias; prm.red_res = red->residual;
        prm.red_ln_w = red->ln_w; prm.red_ln_b = red->ln_b; prm.red_out = red->x_out;
    } else {
        SBEV_REQUIRE(ldx >= layers[0].K, SBEV_ERR_INVALID, "%s: ldx < K", who);
        prm.x_vec4 = ((layers[0].K & 3) == 0 && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) ? 1 : 0;
    }
    prm.x = x; prm.ldx = ldx; prm.M = M; prm.n_layers = n_layers;
    prm.aux_proposal = refine_proposal; prm.aux_time_diff = refine_time_diff; prm.aux_Q = refine_Q > 0 ? refine_Q : 1; prm.aux_T = refine_T;
    prm.blob = reinterpret_cast<const uint8_t*>(blob); prm.blob_stride = blob_stride_bytes; prm.blob_bytes = (int)bytes;
    prm.xld = kmax + 8; prm.yld = nex + 4;
    // rows per tile: 32 when the activation buffers still fit next to the weights (and there are enough rows to fill such tiles), else 16
    auto smem_for = [&](int rt) {
        return (size_t)((bytes + 1023) & ~1023ll) + (size_t)2 * rt * prm.xld * 2 + (size_t)rt * prm.yld * 4 + 1024;
    };
    const size_t smem_cap = 226 * 1024;               // 227 KB opt-in maximum minus the kernel's static shared memory (the weight barrier)
    int RT = get_option(OPT_DENSE_WS_RT);
    if (RT != 16 && RT != 32) RT = (M <= 16 * 16) ? 16 : 32;
    if (smem_for(RT) > smem_cap) RT = 16;
    SBEV_REQUIRE(smem_for(RT) <= smem_cap, SBEV_ERR_UNSUPPORTED, "%s: the chain's weights (%lld bytes per CTA) do not fit in shared memory", who, bytes);
    prm.RT = RT;
    const size_t smem = smem_for(RT);
    {
        cudaError_t attr_err = cudaSuccess;
        SBEV_PER_DEVICE_ONCE(attr_err = cudaFuncSetAttribute(dense_chain_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        if (attr_err != cudaSuccess) { cudaGetLastError(); set_error("%s(shared-memory opt-in): %s", who, cudaGetErrorString(attr_err)); return SBEV_ERR_CUDA; }
    }
    // clusters: at most what can be resident at once (2 per GPC on a B200), at least one row tile each
    static std::mutex mu;
    static std::unordered_map<int, int> resident;
    int max_clusters = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(WS_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = WS_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        auto it = resident.find(dev);
        if (it == resident.end()) {
            cudaLaunchConfig_t probe = cfg;
            probe.gridDim = dim3(WS_CL * 16);
            probe.dynamicSmemBytes = smem_cap;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, dense_chain_ws_kernel, &probe) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 12; }
            it = resident.emplace(dev, n).first;
        }
        max_clusters = it->second;
    }
    int clusters = (M + RT - 1) / RT;
    if (clusters > max_clusters) clusters = max_clusters;
    prm.rows_per_cluster = (M + clusters - 1) / clusters;
    clusters = (M + prm.rows_per_cluster - 1) / prm.rows_per_cluster;
    cfg.gridDim = dim3(WS_CL * clusters);
    cfg.numAttrs = get_option(OPT_PDL) ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, dense_chain_ws_kernel, prm);
    if (e != cudaSuccess) { set_error("%s(cluster launch): %s", who, cudaGetErrorString(e)); return SBEV_ERR_CUDA; }
    return check_launch(who);
}

}  // namespace sbev

using namespace sbev;

extern "C" long long sbev_dense_chain_ws_blob_bytes(int n_layers, const sbev_dense_layer* layers) {
    if (n_layers < 1 || n_layers > WS_MAX_LAYERS || layers == nullptr) return -1;
    for (int i = 0; i < n_layers; ++i)
        if (layers[i].N <= 0 || layers[i].Kpad <= 0 || (layers[i].Kpad & 63)) return -1;
    return ws_blob_bytes(n_layers, layers, nullptr, nullptr);
}

extern "C" int sbev_dense_chain_ws_fwd(const float* x, int ldx, int M, int n_layers, const sbev_dense_layer* layers,
                                       const uint16_t* blob, long long blob_stride_bytes,
                                       const float* refine_proposal, const float* refine_time_diff, int refine_Q, int refine_T,
                                       void* stream) {
    SBEV_REQUIRE(x != nullptr, SBEV_ERR_INVALID, "sbev_dense_chain_ws_fwd: null pointer");
    return ws_launch(x, ldx, nullptr, M, n_layers, layers, blob, blob_stride_bytes, refine_proposal, refine_time_diff, refine_Q, refine_T,
                     stream, "sbev_dense_chain_ws_fwd");
}

extern "C" int sbev_dense_chain_ws_reduce_fwd(const float* partial, int nsplit, const float* bias, const float* residual,
                                              const float* ln_w, const float* ln_b, float* x_out,
                                              int M, int n_layers, const sbev_dense_layer* layers,
                                              const uint16_t* blob, long long blob_stride_bytes,
                                              const float* refine_proposal, const float* refine_time_diff, int refine_Q, int refine_T,
                                              void* stream) {
    WsReduce red{partial, nsplit, bias, residual, ln_w, ln_b, x_out};
    return ws_launch(nullptr, 0, &red, M, n_layers, layers, blob, blob_stride_bytes, refine_proposal, refine_time_diff, refine_Q, refine_T,
                     stream, "sbev_dense_chain_ws_reduce_fwd");
}
