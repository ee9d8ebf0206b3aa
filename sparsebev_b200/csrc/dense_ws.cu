// Weights-stationary dense chain (sbev_dense_chain_ws_fwd / sbev_dense_chain_ws_reduce_fwd): the same Linear (+bias) (+residual)
// (+LayerNorm) (+ReLU) chains as dense.cu, re-partitioned so that no CTA ever streams a whole chain's weights.
//
// Why: dense_chain_mma_kernel gives every CTA 8 rows and makes it stream ALL weights of the chain (0.4-1.1 MB of bf16 hi + lo);
// an SM ingests ~64 B/clk (~120 GB/s), so a chain costs 9-20 us NO MATTER how few rows there are (measured identical from 113
// to 900 rows, tests/perf/kernel_sweep.py).  Here a CLUSTER of 8 CTAs owns a block of rows and SPLITS EVERY LAYER'S OUTPUT
// FEATURES 8 ways: CTA c keeps features [c*SW, (c+1)*SW) of every layer resident in shared memory for the whole kernel -- one
// bulk copy of <= 144 KB per CTA, issued before griddepcontrol.wait, i.e. under the previous kernel's tail.  Row tiles of 16 rows
// stream through; a CTA runs up to TWO independent row groups (4 warps each, own buffers and barriers) so that one group's
// exchange latency hides under the other's MMAs.  Per layer and group:
//     MMA (16 rows x this CTA's feature slice; rows on the MMA M dimension, bf16x3)  ->  + bias  ->  compact [16][SW] fp32 slice in
//     shared memory  ->  ONE bulk shared->distributed-shared copy per peer (cp.async.bulk.shared::cluster.shared::cta), completion
//     counted in bytes on the RECEIVER's mbarrier  ->  every CTA waits for its own 8 slices, finishes the full rows (pre-LN residual,
//     LayerNorm, ReLU, post residual) and writes the next layer's bf16 (hi, lo) operand locally; row r's global outputs are written by
//     CTA r % 8  ->  "my y tile is free again" = one remote mbarrier arrive per peer (senders wait for 8 of them before the next send).
// There is no cluster-wide barrier inside the loop.  (The first version exchanged rows with 8-byte st.shared::cluster stores and
// barrier.cluster: 5.7 us per layer, 3x slower than the streaming kernel at 900 rows -- profiles/r02_ws_chain.md.)
// A last layer without LayerNorm needs no full rows: every CTA finishes and stores its own features (in-projection + tau,
// sampling heads, cls / reg outputs, box refinement).
// The split-K reduce + out_proj bias + identity + norm2 in front of the FFN (sbev_dense_chain_reduce_fwd's prologue) is distributed
// the same way: CTA c reduces tile rows c and c + 8, stores the fp32 rows (they are a later layer's residual) and bulk-copies their
// bf16 (hi, lo) operand rows to the 7 peers.
// Behavioural reference: the same lines dense.cu cites (/root/reference/models/sparsebev_transformer.py:113-183).
#include "common.cuh"
#include <cuda_bf16.h>
#include <mutex>
#include <unordered_map>

namespace sbev {

constexpr int WS_CL = 8;                 // CTAs per cluster (portable maximum)
constexpr int WS_MAX_LAYERS = 6;
constexpr int WS_THREADS = 256;
constexpr int WS_RT = 16;                // rows per tile (one m16 MMA tile)
constexpr int WS_GT = 128;               // threads per row group
constexpr int WS_RED_K = 256;            // row width the reduce prologue handles

struct WsLayer {
    const float* bias; const float* ln_w; const float* ln_b; const float* residual; float* y;
    __nv_bfloat16* y_hi; __nv_bfloat16* y_lo;
    int K, Kpad, N, SW, flags, ldy, woff, vec;     // SW = features per CTA (multiple of 8); woff = byte offset of this layer inside a CTA's weight
                                                   // blob; vec = slice-mode outputs may be stored as aligned pairs
};
struct WsParams {
    const float* x; int ldx, M, n_layers, groups, rows_per_cluster, xld, ys_floats, stg_floats, group_bytes, x_vec4;
    const float* red_partial; const float* red_bias; const float* red_res; const float* red_ln_w; const float* red_ln_b; float* red_out;
    int red_nsplit;
    const float* aux_proposal; const float* aux_time_diff; int aux_Q, aux_T;
    const uint8_t* blob; long long blob_stride; int blob_bytes;
    unsigned long long* dbg;                  // diagnostics (sbev_dense_chain_ws_debug): per CTA 64 clock stamps of thread 0, or NULL
    WsLayer layer[WS_MAX_LAYERS];
};

__device__ __forceinline__ uint32_t ws_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ws_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void ws_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void ws_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void ws_group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(WS_GT) : "memory"); }
__device__ __forceinline__ uint32_t ws_mapa(uint32_t addr, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(rank));
    return remote;
}
__device__ __forceinline__ void ws_mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_u32(b)), "r"(count));
}
__device__ __forceinline__ void ws_mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ws_u32(b)), "r"(bytes) : "memory");
}
// wait for the phase of the given parity.  Default (CTA-scope acquire) semantics, as in CUTLASS's cluster pipelines: the data of a remote
// bulk copy is ordered by its complete_tx on this barrier.  (An explicit .acquire.cluster makes ptxas add an L1 invalidate, CCTL.IVALL,
// to every wait, and .release.cluster on the remote arrive a GPU-scope MEMBAR: profiles/r02_ws_chain.md.)
__device__ __forceinline__ void ws_mbar_wait(uint64_t* b, uint32_t parity) {
    const uint32_t addr = ws_u32(b);
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void ws_mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// bulk copy: this CTA's shared memory -> a peer's, completion (bytes) on the peer's mbarrier
__device__ __forceinline__ void ws_bulk_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(mbar_cluster) : "memory");
}
__device__ __forceinline__ void ws_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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

// 16 rows x (NT x 8) features: v[j][.] = D fragment (rows g8 / g8 + 8, features 2 t4 / 2 t4 + 1 of feature tile j) of x . W^T, bf16x3.
// w_row = shared address of this lane's row of the FIRST feature tile in the layer's first (hi) k chunk; the next feature tile is
// 8 rows = 1024 bytes further; every 64-wide k chunk is [hi: SW rows x 128 B][lo: SW rows x 128 B].
template <int NT>
__device__ __forceinline__ void ws_mma_tiles(float (&v)[NT][4], uint32_t w_row, int tile_bytes, int kchunks,
                                             uint32_t xh_addr, uint32_t xl_addr, int b_row, int b_sel) {
    float acc[NT][6][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int a = 0; a < 6; ++a) { acc[j][a][0] = acc[j][a][1] = acc[j][a][2] = acc[j][a][3] = 0.f; }
    for (int kc = 0; kc < kchunks; ++kc) {
        const uint32_t wh = w_row + (uint32_t)(kc * 2 * tile_bytes);
        const uint32_t wlo = wh + tile_bytes;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            uint32_t bh[NT][4], bl[NT][4], ah[2][4], al[2][4];
            const uint32_t sw = (uint32_t)(((4 * kk + b_sel) ^ b_row) << 4);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                ws_ldsm_x4(bh[j], wh + j * 1024 + sw);
                ws_ldsm_x4(bl[j], wlo + j * 1024 + sw);
            }
            const uint32_t xo = (uint32_t)((kc * 64 + kk * 32) * 2);
            ws_ldsm_x4(ah[0], xh_addr + xo);
            ws_ldsm_x4(al[0], xl_addr + xo);
            ws_ldsm_x4(ah[1], xh_addr + xo + 32);
            ws_ldsm_x4(al[1], xl_addr + xo + 32);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                ws_mma(acc[j][0], ah[0], bh[j][0], bh[j][1]);
                ws_mma(acc[j][2], al[0], bh[j][0], bh[j][1]);
                ws_mma(acc[j][4], ah[0], bl[j][0], bl[j][1]);
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                ws_mma(acc[j][1], ah[1], bh[j][2], bh[j][3]);
                ws_mma(acc[j][3], al[1], bh[j][2], bh[j][3]);
                ws_mma(acc[j][5], ah[1], bl[j][2], bl[j][3]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) v[j][i] = ((acc[j][2][i] + acc[j][3][i]) + (acc[j][4][i] + acc[j][5][i])) + (acc[j][0][i] + acc[j][1][i]);
}

__global__ void __launch_bounds__(WS_THREADS, 1)
dense_chain_ws_kernel(const __grid_constant__ WsParams prm) {
    extern __shared__ uint8_t ws_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* wsm = smem;                                                                   // this CTA's weight slices of every layer
    __shared__ uint64_t wbar;                 // weights landed
    __shared__ uint64_t full_bar[2];          // per group: the 8 slices of the current exchange landed in ys (tx bytes)
    __shared__ uint64_t free_bar[2];          // per group: all 8 CTAs finished reading their ys of the last exchange (8 arrivals)
    __shared__ uint64_t xfull_bar[2];         // per group: the peers' operand rows of the reduce prologue landed (tx bytes)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = warp >> 2, gw = warp & 3, gtid = tid & (WS_GT - 1);
    const uint32_t crank = ws_ctarank();
    const int cluster = blockIdx.x / WS_CL;
    const int row_lo = cluster * prm.rows_per_cluster;
    const int row_hi = min(prm.M, row_lo + prm.rows_per_cluster);
    const int XLD = prm.xld;
    unsigned long long* dbg = (prm.dbg != nullptr && tid == 0) ? prm.dbg + (long long)blockIdx.x * 64 : nullptr;
    int dbg_n = 0;
#define WS_STAMP() do { if (dbg != nullptr && dbg_n < 64) dbg[dbg_n++] = (unsigned long long)clock64(); } while (0)
    WS_STAMP();                                                                            // 0: entry

    if (tid == 0) {
        ws_mbar_init(&wbar, 1);
        for (int i = 0; i < 2; ++i) { ws_mbar_init(&full_bar[i], 1); ws_mbar_init(&free_bar[i], WS_CL); ws_mbar_init(&xfull_bar[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ws_fence_async();
        // the weights are constants of the model: no preceding kernel writes them, so their copy may start before griddepcontrol.wait
        ws_mbar_expect(&wbar, (uint32_t)prm.blob_bytes);
        const uint8_t* src = prm.blob + (long long)crank * prm.blob_stride;
        for (int off = 0; off < prm.blob_bytes; off += 32768) {
            const int n = min(32768, prm.blob_bytes - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(ws_u32(wsm + off)), "l"(src + off), "r"(n), "r"(ws_u32(&wbar)) : "memory");
        }
    }
    WS_STAMP();                                                                            // 1: barriers initialised, weight copy issued
    pdl_wait();
    pdl_trigger();
    WS_STAMP();                                                                            // 2: the preceding kernel has completed
    // every CTA of the cluster is running and has initialised its barriers (remote copies / arrives may target it from now on)
    ws_cluster_arrive();
    ws_cluster_wait();
    WS_STAMP();                                                                            // 3: cluster up

    if (g < prm.groups) {
        uint8_t* gbase = smem + ((prm.blob_bytes + 1023) & ~1023) + g * prm.group_bytes;
        __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(gbase);
        __nv_bfloat16* xl = xh + WS_RT * XLD;
        float* ys = reinterpret_cast<float*>(xl + WS_RT * XLD);                            // [8 slices][16 rows][SW] of the current exchange
        float* stg = ys + prm.ys_floats;                                                   // [2][16 rows][SW]: this CTA's slice, double-buffered
        uint64_t* full = &full_bar[g];
        uint64_t* freeb = &free_bar[g];
        uint64_t* xfull = &xfull_bar[g];
        const int g8 = lane >> 2, t4 = lane & 3;
        // ldmatrix lane roles.  A (rows x k, from the operand tile): matrix (lane >> 3) = (rows +0 / +8, k +0 / +8);
        // B (features x k, from the weight slice): matrix (lane >> 3) = k chunk (lane >> 3) of a 32-wide k group
        const int a_row = (lane & 7) + 8 * ((lane >> 3) & 1), a_kofs = 8 * (lane >> 4);
        const int b_row = lane & 7, b_sel = lane >> 3;
        const uint32_t xh_addr = ws_u32(xh + a_row * XLD + a_kofs), xl_addr = ws_u32(xl + a_row * XLD + a_kofs);
        const bool reduce_mode = prm.red_partial != nullptr;
        bool weights_ready = false;
        uint32_t e = 0;                       // exchanges done by this group (phase counter of full / free)
        uint32_t xt = 0;                      // reduce-prologue tiles done by this group (phase counter of xfull)

        for (int tile = g; row_lo + WS_RT * tile < row_hi; tile += prm.groups) {
            const int r0 = row_lo + WS_RT * tile;
            ws_group_sync(g);                       // this group's warps are done with the previous tile's operand / y rows
            if (reduce_mode) {
                // ---- input rows = LayerNorm(sum_z partial[z][row] + bias + residual[row]); CTA c owns tile rows c and c + 8:
                // warp (2 j + part) sums the partials z = part, part + 2, ... of row 8 j + c
                const int j = gw >> 1, part = gw & 1;
                const int r = 8 * j + (int)crank, row = r0 + r;
                const bool live = row < row_hi;
                float* scratch = ys;                                          // [4 warps][256] (peers send into ys only after they got these rows)
                float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
                const long long zs = (long long)prm.M * WS_RED_K;
                const float* base = prm.red_partial + (long long)row * WS_RED_K + 4 * lane;
                for (int zb = part; zb < prm.red_nsplit; zb += 12) {
                    float4 t[2][6];
#pragma unroll
                    for (int u = 0; u < 6; ++u)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int z = zb + 2 * u;
                            t[i][u] = (live && z < prm.red_nsplit) ? ldg4(base + 128 * i + (long long)z * zs) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                    for (int u = 0; u < 6; ++u)
#pragma unroll
                        for (int i = 0; i < 2; ++i) { acc[i].x += t[i][u].x; acc[i].y += t[i][u].y; acc[i].z += t[i][u].z; acc[i].w += t[i][u].w; }
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(scratch + gw * WS_RED_K + 4 * lane + 128 * i) = acc[i];
                ws_group_sync(g);
                if (part == 0) {
                    float4 a[2];
                    float sum = 0.f;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int n = 4 * lane + 128 * i;
                        a[i] = *reinterpret_cast<const float4*>(scratch + gw * WS_RED_K + n);
                        const float4 q = *reinterpret_cast<const float4*>(scratch + (gw + 1) * WS_RED_K + n);
                        a[i].x += q.x; a[i].y += q.y; a[i].z += q.z; a[i].w += q.w;
                        if (prm.red_bias) { const float4 b = ldg4(prm.red_bias + n); a[i].x += b.x; a[i].y += b.y; a[i].z += b.z; a[i].w += b.w; }
                        if (prm.red_res && live) { const float4 s = ldg4(prm.red_res + (long long)row * WS_RED_K + n); a[i].x += s.x; a[i].y += s.y; a[i].z += s.z; a[i].w += s.w; }
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
                            const float4 gm = ldg4(prm.red_ln_w + n), b = ldg4(prm.red_ln_b + n);
                            a[i].x = (a[i].x - mean) * rstd * gm.x + b.x; a[i].y = (a[i].y - mean) * rstd * gm.y + b.y;
                            a[i].z = (a[i].z - mean) * rstd * gm.z + b.z; a[i].w = (a[i].w - mean) * rstd * gm.w + b.w;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int n = 4 * lane + 128 * i;
                        const float4 v = live ? a[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                        if (live && prm.red_out) *reinterpret_cast<float4*>(prm.red_out + (long long)row * WS_RED_K + n) = v;
                        uint32_t h0, l0, h1, l1;
                        ws_split2(v.x, v.y, h0, l0); ws_split2(v.z, v.w, h1, l1);
                        *reinterpret_cast<uint2*>(xh + r * XLD + n) = make_uint2(h0, h1);
                        *reinterpret_cast<uint2*>(xl + r * XLD + n) = make_uint2(l0, l1);
                    }
                    __threadfence();                // the fp32 row is read by the whole cluster (a later layer's residual)
                    ws_fence_async();               // the operand row is the source of the bulk copies below
                    __syncwarp();
                    if (lane < WS_CL && lane != (int)crank) {
                        // the peers' MMAs of their previous tile's last layer are done once they have released that exchange's y tile
                        if (e > 0) ws_mbar_wait(freeb, (e - 1) & 1);
                        const uint32_t bar = ws_mapa(ws_u32(xfull), lane);
                        ws_bulk_to_peer(ws_mapa(ws_u32(xh + r * XLD), lane), ws_u32(xh + r * XLD), WS_RED_K * 2, bar);
                        ws_bulk_to_peer(ws_mapa(ws_u32(xl + r * XLD), lane), ws_u32(xl + r * XLD), WS_RED_K * 2, bar);
                    }
                }
                if (gtid == 0) ws_mbar_expect(xfull, (uint32_t)((WS_CL - 1) * 2 * 2 * WS_RED_K * 2));     // 7 peers x 2 rows x (hi, lo) x 512 B
                ws_group_sync(g);                   // this CTA's own two rows
                ws_mbar_wait(xfull, xt & 1);        // the other fourteen
                ++xt;
                WS_STAMP();                         // per tile: input rows staged
            } else {
                // ---- stage the tile's input rows as bf16 (hi, lo), zero-padded to the first layer's padded K; rows beyond row_hi are zero
                const int K0 = prm.layer[0].K, K0p = prm.layer[0].Kpad;
                if (prm.x_vec4) {
                    const int k4 = K0p >> 2, total = WS_RT * k4;
                    for (int i0 = gtid; i0 < total; i0 += 8 * WS_GT) {       // eight loads in flight per thread (a 256-wide tile in one round)
                        float4 v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int i = i0 + u * WS_GT, r = i / k4, k = (i - r * k4) * 4;
                            v[u] = (i < total && r0 + r < row_hi && k < K0) ? ldg4(prm.x + (long long)(r0 + r) * prm.ldx + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int i = i0 + u * WS_GT, r = i / k4, k = (i - r * k4) * 4;
                            if (i >= total) continue;
                            uint32_t h0, l0, h1, l1;
                            ws_split2(v[u].x, v[u].y, h0, l0); ws_split2(v[u].z, v[u].w, h1, l1);
                            *reinterpret_cast<uint2*>(xh + r * XLD + k) = make_uint2(h0, h1);
                            *reinterpret_cast<uint2*>(xl + r * XLD + k) = make_uint2(l0, l1);
                        }
                    }
                } else {
                    for (int i = gtid; i < WS_RT * (K0p / 2); i += WS_GT) {
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
                ws_group_sync(g);
                WS_STAMP();                         // per tile: input rows staged
            }

            for (int li = 0; li < prm.n_layers; ++li) {
                const WsLayer& L = prm.layer[li];
                const int N = L.N, SW = L.SW;
                const bool last = li + 1 == prm.n_layers;
                const bool exchange = !(last && L.ln_w == nullptr);
                if (!weights_ready) { ws_mbar_wait(&wbar, 0); weights_ready = true; WS_STAMP(); }      // (first layer only) weights landed
                // ---- MMA: the 8-feature tiles of this CTA's slice, round-robin over the group's 4 warps (two at a time for wide slices)
                const int ftiles = SW >> 3, kchunks = L.Kpad >> 6;
                const int tile_bytes = SW * 128;                                 // one (hi or lo) SW x 64 bf16 tile
                const uint32_t w_lane = ws_u32(wsm + L.woff) + (uint32_t)(b_row * 128);
                float* stg_e = stg + (e & 1) * prm.stg_floats;
                auto emit = [&](int ft, const float (&v4)[4]) {
                    // D fragment: rows g8 / g8 + 8, features f0 + 2 t4 (+1)
                    const int fl = 8 * ft + 2 * t4, f = (int)crank * SW + fl;
                    if (f >= N) return;
                    float v[4] = {v4[0], v4[1], v4[2], v4[3]};
                    if (L.bias != nullptr) {
                        const float b0 = __ldg(L.bias + f), b1 = f + 1 < N ? __ldg(L.bias + f + 1) : 0.f;
                        v[0] += b0; v[1] += b1; v[2] += b0; v[3] += b1;
                    }
                    if (exchange) {
                        *reinterpret_cast<float2*>(stg_e + g8 * SW + fl) = make_float2(v[0], v[1]);
                        *reinterpret_cast<float2*>(stg_e + (g8 + 8) * SW + fl) = make_float2(v[2], v[3]);
                        return;
                    }
                    // slice mode: finish and store this CTA's features of the tile's rows
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int row = r0 + g8 + 8 * h;
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
                            for (int c = 0; c < 2; ++c) {
                                if (f + c >= N) continue;
                                const float a = ws_finish(prm, L, row, f + c, v[2 * h + c]);
                                if (L.y != nullptr) L.y[yoff + c] = a;
                                if (L.y_hi != nullptr) {
                                    const __nv_bfloat16 hb = __float2bfloat16_rn(a);
                                    L.y_hi[yoff + c] = hb;
                                    L.y_lo[yoff + c] = __float2bfloat16_rn(a - __bfloat162float(hb));
                                }
                            }
                        }
                    }
                };
                if (ftiles >= 8) {
                    for (int ft = 2 * gw; ft < ftiles; ft += 8) {
                        if ((int)crank * SW + 8 * ft >= N) continue;             // (warp-uniform) a slice beyond the layer's features
                        float v[2][4];
                        ws_mma_tiles<2>(v, w_lane + (uint32_t)(ft * 1024), tile_bytes, kchunks, xh_addr, xl_addr, b_row, b_sel);
                        emit(ft, v[0]);
                        if (ft + 1 < ftiles) emit(ft + 1, v[1]);
                    }
                } else {
                    for (int ft = gw; ft < ftiles; ft += 4) {
                        if ((int)crank * SW + 8 * ft >= N) continue;
                        float v[1][4];
                        ws_mma_tiles<1>(v, w_lane + (uint32_t)(ft * 1024), tile_bytes, kchunks, xh_addr, xl_addr, b_row, b_sel);
                        emit(ft, v[0]);
                    }
                }
                WS_STAMP();                              // per layer: MMAs done (this warp)
                if (!exchange) continue;                 // (slice mode is the last layer of the chain)

                // ---- exchange: this CTA's [16][SW] slice -> slot `crank` of every CTA's y tile
                ws_fence_async();                        // the slice is the source of bulk copies
                ws_group_sync(g);
                const uint32_t slice_bytes = (uint32_t)(WS_RT * SW * 4);
                if (gw == 0 && lane < WS_CL) {
                    if (e > 0) ws_mbar_wait(freeb, (e - 1) & 1);             // all 8 CTAs are done reading the previous exchange's y tile
                    ws_bulk_to_peer(ws_mapa(ws_u32(ys) + crank * slice_bytes, lane), ws_u32(stg_e), slice_bytes, ws_mapa(ws_u32(full), lane));
                }
                if (gtid == 0) ws_mbar_expect(full, WS_CL * slice_bytes);
                WS_STAMP();                              // per exchange layer: slice sent
                // full rows: warp gw finishes rows gw, gw + 4, gw + 8, gw + 12; residual rows and LayerNorm vectors are fetched while the slices fly
                const bool has_ln = L.ln_w != nullptr;
                const bool has_res = L.residual != nullptr;
                const bool pre_res = (L.flags & SBEV_DENSE_RES_PRE_LN) && has_res;
                const bool relu = (L.flags & SBEV_DENSE_RELU) != 0;
                const int Kn = last ? 0 : prm.layer[li + 1].Kpad;
                if (N <= 256) {
                    // two float4 per lane per row; the four rows of the warp in lock step
                    int yo[2];                            // offset of this lane's float4 i inside a row's slices: slice * 16 * SW + (n % SW)
                    bool in[2];
                    float4 lw[2], lb[2], q[4][2], v[4][2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int n = 4 * lane + 128 * i;
                        in[i] = n < N;
                        const int s = n / SW;
                        yo[i] = s * WS_RT * SW + (n - s * SW);
                        lw[i] = (has_ln && in[i]) ? ldg4(L.ln_w + n) : make_float4(1.f, 1.f, 1.f, 1.f);
                        lb[i] = (has_ln && in[i]) ? ldg4(L.ln_b + n) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int row = r0 + gw + 4 * jj;
                            q[jj][i] = (has_res && in[i] && row < row_hi) ? ws_ldcg4(L.residual + (long long)row * N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    ws_mbar_wait(full, e & 1);
                    WS_STAMP();                          // per exchange layer: all 8 slices here
                    float s[4], mean[4], rstd[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int r = gw + 4 * jj;
                        s[jj] = 0.f;
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            v[jj][i] = in[i] ? *reinterpret_cast<const float4*>(ys + yo[i] + r * SW) : make_float4(0.f, 0.f, 0.f, 0.f);
                            if (pre_res) { v[jj][i].x += q[jj][i].x; v[jj][i].y += q[jj][i].y; v[jj][i].z += q[jj][i].z; v[jj][i].w += q[jj][i].w; }
                            s[jj] += (v[jj][i].x + v[jj][i].y) + (v[jj][i].z + v[jj][i].w);
                        }
                        mean[jj] = 0.f; rstd[jj] = 1.f;
                    }
                    if (has_ln) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) s[jj] += __shfl_xor_sync(0xffffffffu, s[jj], o);
                        float ss[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            mean[jj] = s[jj] / (float)N;
                            ss[jj] = 0.f;
#pragma unroll
                            for (int i = 0; i < 2; ++i)
                                if (in[i]) {
                                    const float a = v[jj][i].x - mean[jj], b = v[jj][i].y - mean[jj], c = v[jj][i].z - mean[jj], d = v[jj][i].w - mean[jj];
                                    ss[jj] += (a * a + b * b) + (c * c + d * d);
                                }
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) ss[jj] += __shfl_xor_sync(0xffffffffu, ss[jj], o);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) rstd[jj] = rsqrtf(ss[jj] / (float)N + 1e-5f);
                    }
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int r = gw + 4 * jj, row = r0 + r;
                        const bool live = row < row_hi;
                        const bool owner = live && (r & (WS_CL - 1)) == (int)crank;
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            if (!in[i]) continue;
                            const int n = 4 * lane + 128 * i;
                            float4 o = v[jj][i];
                            if (has_ln) {
                                o.x = (o.x - mean[jj]) * rstd[jj] * lw[i].x + lb[i].x; o.y = (o.y - mean[jj]) * rstd[jj] * lw[i].y + lb[i].y;
                                o.z = (o.z - mean[jj]) * rstd[jj] * lw[i].z + lb[i].z; o.w = (o.w - mean[jj]) * rstd[jj] * lw[i].w + lb[i].w;
                            }
                            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                            if (has_res && !pre_res) { o.x += q[jj][i].x; o.y += q[jj][i].y; o.z += q[jj][i].z; o.w += q[jj][i].w; }
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
                        if (!last)
                            for (int k = N + 4 * lane; k < Kn; k += 128) {           // zero beyond N up to the next layer's padded K
                                *reinterpret_cast<uint2*>(xh + r * XLD + k) = make_uint2(0u, 0u);
                                *reinterpret_cast<uint2*>(xl + r * XLD + k) = make_uint2(0u, 0u);
                            }
                    }
                } else {
                    // wide rows (256 < N <= 512): one row at a time, up to four float4 per lane
                    ws_mbar_wait(full, e & 1);
                    const int per = (N + 127) >> 7;
                    for (int jj = 0; jj < 4; ++jj) {
                        const int r = gw + 4 * jj, row = r0 + r;
                        const bool live = row < row_hi;
                        float4 v[4], q[4];
                        float s = 0.f;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int n = 4 * lane + 128 * i;
                            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            q[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (i < per && n < N) {
                                const int sl = n / SW;
                                v[i] = *reinterpret_cast<const float4*>(ys + sl * WS_RT * SW + r * SW + (n - sl * SW));
                                if (has_res && live) q[i] = ws_ldcg4(L.residual + (long long)row * N + n);
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
                                    const float4 gm = ldg4(L.ln_w + n), b = ldg4(L.ln_b + n);
                                    o.x = (o.x - mean) * rstd * gm.x + b.x; o.y = (o.y - mean) * rstd * gm.y + b.y;
                                    o.z = (o.z - mean) * rstd * gm.z + b.z; o.w = (o.w - mean) * rstd * gm.w + b.w;
                                }
                                if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                                if (has_res && !pre_res) { o.x += q[i].x; o.y += q[i].y; o.z += q[i].z; o.w += q[i].w; }
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
                            for (int k = N + 4 * lane; k < Kn; k += 128) {
                                *reinterpret_cast<uint2*>(xh + r * XLD + k) = make_uint2(0u, 0u);
                                *reinterpret_cast<uint2*>(xl + r * XLD + k) = make_uint2(0u, 0u);
                            }
                    }
                }
                ws_group_sync(g);                        // the next layer's operand rows are complete; nobody reads this y tile any more
                WS_STAMP();                              // per exchange layer: rows finished
                // release: tell all 8 CTAs that this CTA's y tile may be overwritten (skipped after the group's very last exchange)
                const bool more = !(last && row_lo + WS_RT * (tile + prm.groups) >= row_hi);
                if (more && gtid < WS_CL) ws_mbar_arrive_remote(ws_mapa(ws_u32(freeb), gtid));
                ++e;
            }
        }
    }
    WS_STAMP();                                                                            // work done
    // no CTA may exit while a peer could still copy into, or arrive on, its shared memory
    ws_cluster_arrive();
    ws_cluster_wait();
    WS_STAMP();                                                                            // exit
#undef WS_STAMP
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

static unsigned long long* g_ws_dbg = nullptr;         // sbev_dense_chain_ws_debug

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
    int kmax = 64, swx = 0;
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
            swx = swx > sw[i] ? swx : sw[i];
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
        SBEV_REQUIRE(layers[n_layers - 1].ln_w != nullptr, SBEV_ERR_UNSUPPORTED, "%s: with the reduce prologue the last layer must end in LayerNorm", who);
        SBEV_REQUIRE((red->ln_w == nullptr) == (red->ln_b == nullptr), SBEV_ERR_INVALID, "%s: ln_w/ln_b go together", who);
        SBEV_REQUIRE(((reinterpret_cast<uintptr_t>(red->partial) | reinterpret_cast<uintptr_t>(red->bias) | reinterpret_cast<uintptr_t>(red->residual) |
                       reinterpret_cast<uintptr_t>(red->ln_w) | reinterpret_cast<uintptr_t>(red->ln_b) | reinterpret_cast<uintptr_t>(red->x_out)) & 15) == 0,
                     SBEV_ERR_INVALID, "%s: reduce operands must be 16-byte aligned", who);
        swx = swx > 8 ? swx : 8;               // the y tile doubles as the reduce scratch: 4 x 256 floats <= 8 x 16 x SW
        prm.red_partial = red->partial; prm.red_nsplit = red->nsplit; prm.red_bias = red->bias; prm.red_res = red->residual;
        prm.red_ln_w = red->ln_w; prm.red_ln_b = red->ln_b; prm.red_out = red->x_out;
    } else {
        SBEV_REQUIRE(ldx >= layers[0].K, SBEV_ERR_INVALID, "%s: ldx < K", who);
        prm.x_vec4 = ((layers[0].K & 3) == 0 && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) ? 1 : 0;
    }
    prm.x = x; prm.ldx = ldx; prm.M = M; prm.n_layers = n_layers;
    prm.aux_proposal = refine_proposal; prm.aux_time_diff = refine_time_diff; prm.aux_Q = refine_Q > 0 ? refine_Q : 1; prm.aux_T = refine_T;
    prm.blob = reinterpret_cast<const uint8_t*>(blob); prm.blob_stride = blob_stride_bytes; prm.blob_bytes = (int)bytes;
    prm.xld = kmax + 8;
    prm.dbg = g_ws_dbg;
    prm.ys_floats = WS_CL * WS_RT * swx;
    prm.stg_floats = WS_RT * swx;
    prm.group_bytes = (2 * WS_RT * prm.xld * 2 + (prm.ys_floats + 2 * prm.stg_floats) * 4 + 127) & ~127;
    // row groups per CTA: two when their buffers fit next to the weights (and the caller does not force one)
    const size_t smem_cap = 226 * 1024;               // 227 KB opt-in maximum minus the kernel's static shared memory (barriers)
    auto smem_for = [&](int groups) { return (size_t)((bytes + 1023) & ~1023ll) + (size_t)groups * prm.group_bytes + 1024; };
    int groups = get_option(OPT_DENSE_WS_GROUPS);
    if (groups != 1 && groups != 2) groups = 2;
    if (smem_for(groups) > smem_cap) groups = 1;
    SBEV_REQUIRE(smem_for(groups) <= smem_cap, SBEV_ERR_UNSUPPORTED, "%s: the chain's weights (%lld bytes per CTA) do not fit in shared memory", who, bytes);
    {
        cudaError_t attr_err = cudaSuccess;
        SBEV_PER_DEVICE_ONCE(attr_err = cudaFuncSetAttribute(dense_chain_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        if (attr_err != cudaSuccess) { cudaGetLastError(); set_error("%s(shared-memory opt-in): %s", who, cudaGetErrorString(attr_err)); return SBEV_ERR_CUDA; }
    }
    // clusters: at most what can be resident at once (2 per GPC on a B200), a whole number of row tiles each
    static std::mutex mu;
    static std::unordered_map<int, int> resident;
    int max_clusters = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(WS_THREADS); cfg.stream = (cudaStream_t)stream;
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
    const int tiles = (M + WS_RT - 1) / WS_RT;
    const int tiles_per_cluster = (tiles + max_clusters - 1) / max_clusters;
    if (tiles_per_cluster < 2) groups = 1;            // one tile per cluster: a second group would have nothing to do
    prm.groups = groups;
    prm.rows_per_cluster = tiles_per_cluster * WS_RT;
    const int clusters = (M + prm.rows_per_cluster - 1) / prm.rows_per_cluster;
    cfg.dynamicSmemBytes = smem_for(groups);
    cfg.gridDim = dim3(WS_CL * clusters);
    cfg.numAttrs = get_option(OPT_PDL) ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, dense_chain_ws_kernel, prm);
    if (e != cudaSuccess) { set_error("%s(cluster launch): %s", who, cudaGetErrorString(e)); return SBEV_ERR_CUDA; }
    return check_launch(who);
}

}  // namespace sbev

using namespace sbev;

extern "C" int sbev_dense_chain_ws_debug(unsigned long long* stamps) {
    g_ws_dbg = stamps;
    return SBEV_OK;
}

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
