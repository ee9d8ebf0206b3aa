// bf16 "TN" GEMM on the 5th-generation tensor cores (tcgen05 / UMMA, sm_100a only):
//     C[M,N] (fp32) = sum_s A_s[M,K] . B_s[N,K]^T  (+ bias)
// Used for the two dense contractions of AdaptiveMixing
// (/root/reference/models/sparsebev_transformer.py:358 parameter_generator, :378 out_proj), which the
// reference runs as fp32 cuBLAS SGEMMs.  fp32-grade accuracy is kept with the bf16x3 split
//     a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi        (|error| <= ~3 * 2^-18 |a.b|)
// expressed as three operand-pair "segments" that accumulate into the SAME TMEM accumulator, so the
// split costs tensor-core time only -- no extra passes over C.
//
// Structure (persistent CTAs, 192 threads, one CTA per SM):
//   warp 0   TMA producer: cp.async.bulk.tensor 2-D loads of 128x64 (A) and BNx64 (B) bf16 boxes, 128-byte swizzled,
//            into a shared-memory ring, completion on mbarriers.  In bf16x3 mode one stage holds {A_hi, B_hi, A_lo, B_lo}.
//   warp 1   allocates TMEM (2 accumulators x BN fp32 columns x 128 lanes); ONE thread issues
//            tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16): hi.hi, hi.lo, lo.hi per k-block, then
//            tcgen05.commit's the stage's "empty" barrier; a commit per tile signals the epilogue.
//   warps 2-5  epilogue: tcgen05.ld 32x32b (each warp its own TMEM lane quarter) -> +bias -> global, overlapping the
//            next tile's MMAs thanks to the second accumulator.
// All mbarrier waits are bounded: a protocol bug traps instead of hanging the GPU.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <mutex>
#include <unordered_map>

namespace sbev {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;                 // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_MAX_SEG = 3;

// ------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && spins > (1u << 26)) __trap();      // ~seconds: protocol bug, fail loudly instead of hanging
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// ---- CTA-pair (cta_group::2) flavours: the two CTAs of a cluster act as one 256-row MMA
__device__ __forceinline__ uint32_t gemm_cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void gemm_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared-memory object of this CTA) as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cluster_addr_of(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
// TMA load into THIS CTA's shared memory whose completion bytes are counted on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// commit: arrive (once all previously issued MMAs of the pair completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle, densely packed 128 B rows:
// canonical layout ((8,n),2):((8,SBO),1) in 16 B units -> LBO = 1 (unused), SBO = 1024 B (8 rows),
// descriptor version 1 (Blackwell), layout type 2 (SWIZZLE_128B).  Tile base is 1024 B aligned.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);      // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                            // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                            // version = 1
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}

// Instruction descriptor, kind::f16: D=F32 (bits 4-5 = 1), A=B=BF16 (bits 7-9, 10-12 = 1), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28.
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------- kernel
//   * persistent CTAs (grid = min(#tiles, #SMs)) walk a static tile schedule, M-tile fastest so the CTAs that run
//     concurrently share the same B (weight) tiles in L2;
//   * TWO TMEM accumulators: the epilogue of tile i drains accumulator i&1 while the MMA warp already fills the
//     other one for tile i+1 (tmem_full / tmem_empty mbarriers);
//   * in bf16x3 mode one pipeline stage holds {A_hi, A_lo, B_hi, B_lo} for a k-block and the MMA warp issues the
//     three products hi.hi, hi.lo, lo.hi from them -> every operand tile is fetched once, not 1.5x.
//   * epilogue: TMEM -> registers (+bias) -> 128-byte-swizzled shared staging -> TMA tensor stores of 32x32 fp32
//     boxes (full 128 B row segments, rows beyond M clipped by the tensor map), double-buffered per warp.
struct GemmMapsV2 {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    CUtensorMap c;          // fp32 [split_k][M][N], box {32 cols, 32 rows, 1}, 128-byte swizzle (epilogue TMA stores)
    CUtensorMap c_hi, c_lo; // OUT_SPLIT: bf16 [M][N] each, box {64 cols, 32 rows}, 128-byte swizzle
    int hints = 0;          // L2 cache hints (option "gemm_l2_hints"): 1 = OUT_SPLIT stores evict_first, 2 = B (weight) tile loads evict_last
};

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1), "l"(policy) : "memory");
}

// ARES (A-resident, for small K): the CTA is pinned to ONE M-tile, loads that tile's whole A operand (all k-blocks, hi and
// lo) into shared memory once and then only streams B through the ring while it walks its N-tiles -- a third less
// operand traffic per MMA for the parameter-generation GEMM (K = 256), whose A is the same 900 x 256 query matrix for
// all 256 N-tiles.  Requires X3, split_k == 1, gridDim.x % m_tiles == 0; ARES_KB = number of resident k-blocks.
// OUT_SPLIT: the epilogue stores C as a bf16 (hi, lo) pair (C ~= hi + lo) instead of fp32 -- the layout the mma.sync mix
// kernel consumes by TMA, so the 118 MB parameter tensor is never re-converted.
// PAIR: the kernel is launched in clusters of two CTAs that act as ONE tcgen05 "CTA pair" (cta_group::2): a work unit is a
// 256 x BN output tile; CTA r of the pair owns rows [128 r, 128 r + 128) of it (its own TMEM accumulator lanes), loads
// its 128 rows of A and HALF of the B tile (BN/2 rows) -- the tensor cores of both SMs read both halves -- so each SM
// pulls a third less operand data through L2 per MMA.  Only the leader (rank 0) issues MMAs; TMA completions of both CTAs
// count on the leader's "full" barrier, tcgen05.commit multicasts the "empty" / "accumulator full" arrivals to both CTAs,
// and both epilogues report "accumulator drained" to the leader.
template <int BN, int STAGES, bool X3, int ARES_KB = 0, bool OUT_SPLIT = false, bool PAIR = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_persistent_kernel(const __grid_constant__ GemmMapsV2 maps, int total_kb, int split_k, int m_tiles, int n_tiles, int num_tiles,
                               const float* __restrict__ bias, float* __restrict__ C, int M, int N) {
    constexpr bool ARES = ARES_KB > 0;
    constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static_assert(!(PAIR && ARES_KB > 0), "the A-resident schedule has no CTA-pair form");
    constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * GEMM_BK * 2;      // PAIR: this CTA's half of the B tile
    constexpr int STAGE_BYTES = ARES ? 2 * B_BYTES : (X3 ? 2 : 1) * (A_BYTES + B_BYTES);
    constexpr int ARES_BYTES = ARES_KB * 2 * A_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem_al = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ares = smem_al;                                  // [ARES_KB][A_hi | A_lo] (only when ARES)
    uint8_t* smem = smem_al + ARES_BYTES;                     // operand ring
    uint8_t* stg_base = smem + STAGES * STAGE_BYTES;          // epilogue staging: 4 warps x 2 buffers x [32 rows][128 B], swizzled
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2], ares_bar;
    // ARES tile schedule: M-tile fixed per CTA, N-tiles strided by the number of CTAs that share the M-tile
    const int ares_m = ARES ? (int)(blockIdx.x % m_tiles) : 0;
    const int ares_n0 = ARES ? (int)(blockIdx.x / m_tiles) : 0;
    const int ares_nstep = ARES ? (int)(gridDim.x / m_tiles) : 1;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? gemm_cta_rank() : 0u;                       // 0 = leader of the pair
    const int unit_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;     // work-unit stream of this CTA (pair)
    const int unit_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int m_units = PAIR ? (m_tiles + 1) / 2 : m_tiles;                  // 256-row units per N-tile column
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], PAIR ? 8 : 4); }
        mbar_init(&ares_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) { if (PAIR) tmem_alloc_pair(&tmem_slot, 2 * BN); else tmem_alloc(&tmem_slot, 2 * BN); }
    tc_fence_before();
    __syncthreads();
    if (PAIR) gemm_cluster_sync();          // both CTAs' barriers exist before any remote arrive / peer-counted TMA
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_wait();                 // barriers + TMEM are set up; from here on global memory is read (TMA) and written (epilogue)
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a_hi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b_hi) : "memory");
            if (X3) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a_lo) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b_lo) : "memory");
            }
            int it = 0;
            if (ARES) {
                mbar_expect_tx(&ares_bar, ARES_BYTES);
                for (int kb = 0; kb < ARES_KB; ++kb) {
                    tma_load_2d(ares + kb * 2 * A_BYTES, &maps.a_hi, &ares_bar, kb * GEMM_BK, ares_m * GEMM_BM);
                    tma_load_2d(ares + kb * 2 * A_BYTES + A_BYTES, &maps.a_lo, &ares_bar, kb * GEMM_BK, ares_m * GEMM_BM);
                }
                for (int nt = ares_n0; nt < n_tiles; nt += ares_nstep)
                    for (int kb = 0; kb < total_kb; ++kb, ++it) {
                        const int stage = it % STAGES;
                        mbar_wait(&empty_bar[stage], ((it / STAGES) & 1) ^ 1);
                        uint8_t* st = smem + stage * STAGE_BYTES;
                        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                        tma_load_2d(st, &maps.b_hi, &full_bar[stage], kb * GEMM_BK, nt * BN);
                        tma_load_2d(st + B_BYTES, &maps.b_lo, &full_bar[stage], kb * GEMM_BK, nt * BN);
                    }
            } else
            for (int tile = unit_id; tile < num_tiles; tile += unit_step) {
                const int m0 = ((tile % m_units) * (PAIR ? 2 : 1) + (int)rank) * GEMM_BM;
                const int rest = tile / m_units;
                const int n0 = (rest % n_tiles) * BN + (PAIR ? (int)rank * (BN / 2) : 0);      // PAIR: this CTA's half of the B rows
                const int z = rest / n_tiles;                       // split-K slice: k-blocks [z*total/split, (z+1)*total/split)
                const int kb0 = (int)((long long)z * total_kb / split_k);
                const int kb_cnt = (int)((long long)(z + 1) * total_kb / split_k) - kb0;
                for (int kb = 0; kb < kb_cnt; ++kb, ++it) {
                    const int stage = it % STAGES;
                    mbar_wait(&empty_bar[stage], ((it / STAGES) & 1) ^ 1);
                    uint8_t* st = smem + stage * STAGE_BYTES;
                    const int kc = (kb0 + kb) * GEMM_BK;
                    if (PAIR) {
                        // the leader's barrier counts the bytes of BOTH CTAs' loads of this stage
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                        const uint32_t lbar = cluster_addr_of(&full_bar[stage], 0);
                        tma_load_2d_pair(st, &maps.a_hi, lbar, kc, m0);
                        tma_load_2d_pair(st + A_BYTES, &maps.b_hi, lbar, kc, n0);
                        if (X3) {
                            tma_load_2d_pair(st + A_BYTES + B_BYTES, &maps.a_lo, lbar, kc, m0);
                            tma_load_2d_pair(st + 2 * A_BYTES + B_BYTES, &maps.b_lo, lbar, kc, n0);
                        }
                    } else {
                        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                        tma_load_2d(st, &maps.a_hi, &full_bar[stage], kc, m0);
                        if (maps.hints & 2) tma_load_2d_hint(st + A_BYTES, &maps.b_hi, &full_bar[stage], kc, n0, l2_policy_evict_last());
                        else tma_load_2d(st + A_BYTES, &maps.b_hi, &full_bar[stage], kc, n0);
                        if (X3) {
                            tma_load_2d(st + A_BYTES + B_BYTES, &maps.a_lo, &full_bar[stage], kc, m0);
                            if (maps.hints & 2) tma_load_2d_hint(st + 2 * A_BYTES + B_BYTES, &maps.b_lo, &full_bar[stage], kc, n0, l2_policy_evict_last());
                            else tma_load_2d(st + 2 * A_BYTES + B_BYTES, &maps.b_lo, &full_bar[stage], kc, n0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_f32(PAIR ? 2 * GEMM_BM : GEMM_BM, BN);
            int it = 0, lt = 0;
            if (ARES) { mbar_wait(&ares_bar, 0); tc_fence_after(); }
            const int my_tiles = ARES ? (n_tiles - ares_n0 + ares_nstep - 1) / ares_nstep
                                      : (num_tiles - unit_id + unit_step - 1) / unit_step;
            for (; lt < my_tiles; ++lt) {
                const int as = lt & 1;
                mbar_wait(&tmem_empty_bar[as], ((lt >> 1) & 1) ^ 1);        // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(as * BN);
                int kb_cnt = total_kb;
                if (!ARES) {
                    const int z = ((unit_id + lt * unit_step) / m_units) / n_tiles;
                    kb_cnt = (int)((long long)(z + 1) * total_kb / split_k) - (int)((long long)z * total_kb / split_k);
                }
                for (int kb = 0; kb < kb_cnt; ++kb, ++it) {
                    const int stage = it % STAGES;
                    mbar_wait(&full_bar[stage], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
                    uint64_t a_hi, a_lo, b_hi, b_lo;
                    if (ARES) {
                        const uint32_t ar = smem_u32(ares + kb * 2 * A_BYTES);
                        a_hi = umma_desc_k_sw128(ar); a_lo = umma_desc_k_sw128(ar + A_BYTES);
                        b_hi = umma_desc_k_sw128(st); b_lo = umma_desc_k_sw128(st + B_BYTES);
                    } else {
                        a_hi = umma_desc_k_sw128(st); b_hi = umma_desc_k_sw128(st + A_BYTES);
                        a_lo = umma_desc_k_sw128(st + A_BYTES + B_BYTES); b_lo = umma_desc_k_sw128(st + 2 * A_BYTES + B_BYTES);
                    }
                    auto mma = [&](uint64_t a, uint64_t b, uint32_t acc) {
                        if (PAIR) umma_bf16_pair(tacc, a, b, idesc, acc); else umma_bf16(tacc, a, b, idesc, acc);
                    };
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) mma(a_hi + 2 * k, b_hi + 2 * k, (kb > 0 || k > 0) ? 1u : 0u);
                    if (X3) {
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) mma(a_hi + 2 * k, b_lo + 2 * k, 1u);
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) mma(a_lo + 2 * k, b_hi + 2 * k, 1u);
                    }
                    if (PAIR) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
                }
                if (PAIR) umma_commit_pair(&tmem_full_bar[as]); else umma_commit(&tmem_full_bar[as]);
            }
        }
    } else {
        const int quarter = warp & 3;
        int lt = 0, chunk_ctr = 0;
        const int my_tiles = ARES ? (n_tiles - ares_n0 + ares_nstep - 1) / ares_nstep
                                  : (num_tiles - unit_id + unit_step - 1) / unit_step;
        const uint32_t leader_empty0 = PAIR ? cluster_addr_of(&tmem_empty_bar[0], 0) : 0u;
        const uint32_t leader_empty1 = PAIR ? cluster_addr_of(&tmem_empty_bar[1], 0) : 0u;
        for (; lt < my_tiles; ++lt) {
            const int tile = ARES ? 0 : unit_id + lt * unit_step;
            const int m0 = ARES ? ares_m * GEMM_BM : ((tile % m_units) * (PAIR ? 2 : 1) + (int)rank) * GEMM_BM;
            const int rest = tile / m_units;
            const int n0 = ARES ? (ares_n0 + lt * ares_nstep) * BN : (rest % n_tiles) * BN;
            const int z = ARES ? 0 : rest / n_tiles;
            const int as = lt & 1;
            mbar_wait(&tmem_full_bar[as], (lt >> 1) & 1);
            tc_fence_after();
            const bool add_bias = (bias != nullptr) && (z == 0);
            // staging per epilogue warp: 2 x 4 KB (fp32: alternating boxes; bf16 split: one hi / lo pair), or -- CTA-pair bf16
            // split, whose smaller operand stages leave the room -- two hi / lo pairs so a store drains while the next box is built
            constexpr int STG_PER_WARP = (OUT_SPLIT && PAIR && STAGES <= 2) ? 4 : 2;      // (the 3-stage pair form spends that room on the ring)
            uint8_t* stg_warp = stg_base + (warp - 2) * STG_PER_WARP * 4096;
            uint8_t* my_stg = stg_warp;
            if constexpr (OUT_SPLIT) {
#pragma unroll 1
                for (int c = 0; c < BN / 64; ++c) {
                    // the chunk's 64 bias values first: their L2 round trip overlaps the TMEM load below (profile: the first FADD on a
                    // bias value was 22 % of this kernel's stall samples when the loads sat inside the conversion loop)
                    float4 bq[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) bq[j] = add_bias ? ldg4(bias + n0 + c * 64 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float v[64];
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + c * 64), v);
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + c * 64 + 32), v + 32);
                    // bias + bf16 (hi, lo) split into registers FIRST: the TMA stores of the previous chunk drain the staging
                    // buffers meanwhile (profile: the wait right after the TMEM load was 21 % of this kernel's stall samples)
                    uint32_t hw[8][4], lw[8][4];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {                      // 8 chunks of 8 bf16 (16 B) per 128-byte row
                        const float4 b0 = bq[2 * j], b1 = bq[2 * j + 1];
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float a = v[8 * j + 2 * e] + bb[2 * e], b = v[8 * j + 2 * e + 1] + bb[2 * e + 1];
                            const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                            const float2 hf = __bfloat1622float2(h2);
                            const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - hf.x, b - hf.y);
                            hw[j][e] = *reinterpret_cast<const uint32_t*>(&h2);
                            lw[j][e] = *reinterpret_cast<const uint32_t*>(&l2);
                        }
                    }
                    // the stores that last read this hi / lo buffer pair must have drained it
                    if (STG_PER_WARP == 4) {
                        my_stg = stg_warp + (c & 1) * 8192;          // BN / 64 is even: the alternation carries over from tile to tile
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    } else {
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t off = lane * 128 + ((j ^ (lane & 7)) << 4);
                        *reinterpret_cast<uint4*>(my_stg + off) = make_uint4(hw[j][0], hw[j][1], hw[j][2], hw[j][3]);
                        *reinterpret_cast<uint4*>(my_stg + 4096 + off) = make_uint4(lw[j][0], lw[j][1], lw[j][2], lw[j][3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        if (maps.hints & 1) {        // the parameter tensor is written once and read once, much later: do not let it evict hotter lines
                            const uint64_t pol = l2_policy_evict_first();
                            tma_store_2d_hint(&maps.c_hi, smem_u32(my_stg), n0 + c * 64, m0 + quarter * 32, pol);
                            tma_store_2d_hint(&maps.c_lo, smem_u32(my_stg + 4096), n0 + c * 64, m0 + quarter * 32, pol);
                        } else {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(&maps.c_hi), "r"(smem_u32(my_stg)), "r"(n0 + c * 64), "r"(m0 + quarter * 32) : "memory");
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(&maps.c_lo), "r"(smem_u32(my_stg + 4096)), "r"(n0 + c * 64), "r"(m0 + quarter * 32) : "memory");
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            } else {
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + c * 32), v);
                uint8_t* buf = my_stg + (chunk_ctr & 1) * 4096;
                // the TMA store that last read this buffer (two chunks ago) must have finished reading shared memory
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    if (add_bias) { const float4 bv = ldg4(bias + n0 + c * 32 + i); o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w; }
                    // row = lane (128 B per row), 16-byte chunk j = i/4 stored at j ^ (row & 7): the layout the swizzled tensor map expects
                    *reinterpret_cast<float4*>(buf + lane * 128 + ((((i >> 2) ^ (lane & 7))) << 4)) = o;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                 ::"l"(&maps.c), "r"(smem_u32(buf)), "r"(n0 + c * 32), "r"(m0 + quarter * 32), "r"(z) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                ++chunk_ctr;
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(as ? leader_empty1 : leader_empty0) : "memory");
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[as])) : "memory");
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // all output tiles written before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) gemm_cluster_sync();          // no CTA frees TMEM / exits while its peer may still read its shared memory or signal it
    if (warp == 1) { tc_fence_after(); if (PAIR) tmem_dealloc_pair(tmem_base, 2 * BN); else tmem_dealloc(tmem_base, 2 * BN); }
}

// ------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// 2-D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128 B swizzle, zero OOB fill.
// Encoding a tensor map costs ~1-2 us of host time; the maps are pure functions of (pointer, shape, box), so they are
// memoised (weights never move; activation buffers are recycled by the caller's allocator).
struct MapKey {
    const void* ptr; long long rows, cols; int box_rows;
    bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && box_rows == o.box_rows; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        h ^= std::hash<long long>()(k.rows * 1000003ll + k.cols) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        return h ^ (size_t)k.box_rows;
    }
};

int make_bf16_map(CUtensorMap* out, const void* ptr, long long rows, long long cols, int box_rows) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ptr, rows, cols, box_rows};
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return SBEV_OK; }
    }
    EncodeTiledFn enc = get_encode_fn();
    SBEV_REQUIRE(enc != nullptr, SBEV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SBEV_REQUIRE(r == CUDA_SUCCESS, SBEV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for [%lld,%lld] box %d", (int)r, rows, cols, box_rows);
    {
        std::lock_guard<std::mutex> lk(mu);
        if (cache.size() > 4096) cache.clear();
        cache.emplace(key, *out);
    }
    return SBEV_OK;
}

// General 2-D bf16 map: [rows][cols] row-major, box {box_cols, box_rows}, swizzle_bytes in {0, 64, 128}.
int make_bf16_map_ex(CUtensorMap* out, const void* ptr, long long rows, long long cols, int box_rows, int box_cols, int swizzle_bytes) {
    EncodeTiledFn enc = get_encode_fn();
    SBEV_REQUIRE(enc != nullptr, SBEV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SBEV_REQUIRE(r == CUDA_SUCCESS, SBEV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for [%lld,%lld] box %dx%d", (int)r, rows, cols, box_rows, box_cols);
    return SBEV_OK;
}

// bf16 [M][N] output, box = {64 cols, 32 rows}, 128-byte swizzle on the shared-memory side (TMA stores).
static int make_bf16_store_map(CUtensorMap* out, const void* ptr, long long N, long long M) {
    EncodeTiledFn enc = get_encode_fn();
    SBEV_REQUIRE(enc != nullptr, SBEV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t gstride[1] = {(cuuint64_t)N * 2};
    cuuint32_t box[2] = {64, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SBEV_REQUIRE(r == CUDA_SUCCESS, SBEV_ERR_CUDA, "cuTensorMapEncodeTiled(C bf16) failed (%d)", (int)r);
    return SBEV_OK;
}

// fp32 [Z][M][N] output, box = {32 cols, 32 rows, 1}, 128-byte swizzle on the shared-memory side (TMA stores).
static int make_f32_store_map(CUtensorMap* out, const void* ptr, long long N, long long M, long long Z) {
    EncodeTiledFn enc = get_encode_fn();
    SBEV_REQUIRE(enc != nullptr, SBEV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)Z};
    cuuint64_t gstride[2] = {(cuuint64_t)N * 4, (cuuint64_t)N * M * 4};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SBEV_REQUIRE(r == CUDA_SUCCESS, SBEV_ERR_CUDA, "cuTensorMapEncodeTiled(C) failed (%d)", (int)r);
    return SBEV_OK;
}

}  // namespace sbev

using namespace sbev;

// Launch in clusters of two CTAs (one tcgen05 CTA pair per cluster), with the PDL attribute when enabled.
template <typename... P, typename... A>
static void launch_pair(void (*kernel)(P...), int clusters, size_t smem, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent schedule: no more clusters than can be resident at once (a GPC with a disabled SM leaves a TPC unpaired)
    static std::mutex mu;
    static std::unordered_map<const void*, int> resident;
    int max_clusters = 0;
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = resident.find((const void*)kernel);
        if (it == resident.end()) {
            cudaLaunchConfig_t probe = cfg;
            probe.gridDim = dim3(2 * 74);
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kernel, &probe) != cudaSuccess || n <= 0) { cudaGetLastError(); n = clusters; }
            it = resident.emplace((const void*)kernel, n).first;
        }
        max_clusters = it->second;
    }
    if (clusters > max_clusters) cfg.gridDim = dim3(2 * max_clusters);
    cfg.numAttrs = get_option(OPT_PDL) ? 2 : 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}
constexpr size_t GEMM_PAIR_SMEM = (size_t)3 * 2 * (GEMM_BM * GEMM_BK * 2 + 128 * GEMM_BK * 2) + 4 * 2 * 4096 + 1024;   // 3 stages x {A, B/2} x {hi, lo}
constexpr size_t GEMM_PAIR_SPLIT_SMEM = (size_t)2 * 2 * (GEMM_BM * GEMM_BK * 2 + 128 * GEMM_BK * 2) + 4 * 4 * 4096 + 1024;   // 2 stages + double-buffered (hi, lo) staging

extern "C" int sbev_gemm_bf16_tn(const uint16_t* const* A, const uint16_t* const* B, int nseg,
                                 const float* bias, int M, int N, int K, int split_k, float* C, void* stream) {
    SBEV_REQUIRE(A && B && C, SBEV_ERR_INVALID, "sbev_gemm_bf16_tn: null pointer");
    SBEV_REQUIRE(nseg >= 1 && nseg <= GEMM_MAX_SEG, SBEV_ERR_INVALID, "sbev_gemm_bf16_tn: nseg must be 1..3");
    SBEV_REQUIRE(M > 0 && N > 0 && K > 0 && split_k >= 1, SBEV_ERR_INVALID, "sbev_gemm_bf16_tn: bad sizes");
    SBEV_REQUIRE(K % GEMM_BK == 0, SBEV_ERR_UNSUPPORTED, "sbev_gemm_bf16_tn: K must be a multiple of 64 (got %d)", K);
    SBEV_REQUIRE(N % 128 == 0, SBEV_ERR_UNSUPPORTED, "sbev_gemm_bf16_tn: N must be a multiple of 128 (got %d)", N);
    SBEV_REQUIRE(split_k <= K / GEMM_BK, SBEV_ERR_UNSUPPORTED, "sbev_gemm_bf16_tn: split_k must not exceed K/64");
    SBEV_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
                 SBEV_ERR_INVALID, "sbev_gemm_bf16_tn: C / bias must be 16-byte aligned");
    for (int sg = 0; sg < nseg; ++sg) {
        SBEV_REQUIRE(A[sg] && B[sg], SBEV_ERR_INVALID, "sbev_gemm_bf16_tn: null operand in segment %d", sg);
        SBEV_REQUIRE((reinterpret_cast<uintptr_t>(A[sg]) & 15) == 0 && (reinterpret_cast<uintptr_t>(B[sg]) & 15) == 0,
                     SBEV_ERR_INVALID, "sbev_gemm_bf16_tn: operands must be 16-byte aligned");
    }
    const bool x3_pattern = nseg == 3 && A[0] == A[1] && B[0] == B[2];
    if (nseg == 1 || x3_pattern) {
        // ---- v2: persistent, double-buffered TMEM, shared operand tiles for the three bf16x3 products
        const int num_sms = device_num_sms();
        const int m_tiles_pre = (M + GEMM_BM - 1) / GEMM_BM;
        // A-resident variant: small K, many N-tiles per M-tile (the parameter-generation GEMM)
        const bool ares = get_option(OPT_GEMM_IMPL) == 1 && x3_pattern && split_k == 1 && K == 256 && m_tiles_pre <= num_sms &&
                          (N / 128) >= 2 * (num_sms / m_tiles_pre);
        const bool wide = !ares && (N % 256 == 0);
        const int BNv = wide ? 256 : 128;
        // (a single row tile -- a query shard of <= 128 rows -- would leave half of every 256-row pair unit empty: single CTAs)
        if (get_option(OPT_GEMM_IMPL) != 3 && get_option(OPT_GEMM_IMPL) != 1 && x3_pattern && wide && num_sms >= 2 && m_tiles_pre > 1) {
            // ---- CTA-pair (cta_group::2) schedule: 256 x 256 units, each CTA loads its A rows and half of the B tile
            GemmMapsV2 mp;
            int rc = make_bf16_map(&mp.a_hi, A[0], M, K, GEMM_BM);   if (rc) return rc;
            rc = make_bf16_map(&mp.b_hi, B[0], N, K, 128);           if (rc) return rc;
            rc = make_bf16_map(&mp.a_lo, A[2], M, K, GEMM_BM);       if (rc) return rc;
            rc = make_bf16_map(&mp.b_lo, B[1], N, K, 128);           if (rc) return rc;
            rc = make_f32_store_map(&mp.c, C, N, M, split_k);        if (rc) return rc;
            const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM, n_tiles = N / 256;
            const int num_units = ((m_tiles + 1) / 2) * n_tiles * split_k;
            const int clusters = num_units < num_sms / 2 ? num_units : num_sms / 2;
            SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(gemm_bf16_tn_persistent_kernel<256, 3, true, 0, false, true>,
                                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_PAIR_SMEM));
            launch_pair(gemm_bf16_tn_persistent_kernel<256, 3, true, 0, false, true>, clusters, GEMM_PAIR_SMEM, (cudaStream_t)stream,
                        mp, K / GEMM_BK, split_k, m_tiles, n_tiles, num_units, bias, C, M, N);
            return check_launch("sbev_gemm_bf16_tn(pair)");
        }
        GemmMapsV2 mp;
        int rc = make_bf16_map(&mp.a_hi, A[0], M, K, GEMM_BM);            if (rc) return rc;
        rc = make_bf16_map(&mp.b_hi, B[0], N, K, BNv);                    if (rc) return rc;
        rc = make_bf16_map(&mp.a_lo, x3_pattern ? A[2] : A[0], M, K, GEMM_BM); if (rc) return rc;
        rc = make_bf16_map(&mp.b_lo, x3_pattern ? B[1] : B[0], N, K, BNv);     if (rc) return rc;
        rc = make_f32_store_map(&mp.c, C, N, M, split_k);                 if (rc) return rc;
        const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM, n_tiles = N / BNv;
        const int num_tiles = m_tiles * n_tiles * split_k;
        const int grid = ares ? (num_sms / m_tiles) * m_tiles : (num_tiles < num_sms ? num_tiles : num_sms);
        const int kbs = K / GEMM_BK;
        cudaStream_t st = (cudaStream_t)stream;
#define SBEV_GEMM_V2(BNN, STG, XX)                                                                                               \
        do {                                                                                                                     \
            constexpr size_t smem_v2 = (size_t)STG * (XX ? 2 : 1) * (GEMM_BM * GEMM_BK * 2 + BNN * GEMM_BK * 2) + 4 * 2 * 4096 + 1024;          \
            SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(gemm_bf16_tn_persistent_kernel<BNN, STG, XX>,                         \
                                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v2));       \
            launch_pdl(gemm_bf16_tn_persistent_kernel<BNN, STG, XX>, dim3(grid), dim3(GEMM_THREADS), smem_v2, st, mp, kbs, split_k, m_tiles, n_tiles, num_tiles, bias, C, M, N); \
        } while (0)
        if (ares) {
            constexpr size_t smem_ar = (size_t)4 * 2 * (GEMM_BM * GEMM_BK * 2) + (size_t)2 * 2 * (128 * GEMM_BK * 2) + 4 * 2 * 4096 + 1024;
            SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(gemm_bf16_tn_persistent_kernel<128, 2, true, 4>,
                                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ar));
            launch_pdl(gemm_bf16_tn_persistent_kernel<128, 2, true, 4>, dim3(grid), dim3(GEMM_THREADS), smem_ar, st, mp, kbs, 1, m_tiles, n_tiles, num_tiles, bias, C, M, N);
        } else if (x3_pattern) { if (wide) SBEV_GEMM_V2(256, 2, true); else SBEV_GEMM_V2(128, 3, true); }
        else            { if (wide) SBEV_GEMM_V2(256, 4, false); else SBEV_GEMM_V2(128, 6, false); }
#undef SBEV_GEMM_V2
        return check_launch("sbev_gemm_bf16_tn(v2)");
    }
    SBEV_REQUIRE(false, SBEV_ERR_UNSUPPORTED, "sbev_gemm_bf16_tn: nseg must be 1, or 3 with the (A0,B0),(A0,B1),(A1,B0) bf16x3 pattern");
    return check_launch("sbev_gemm_bf16_tn");
}

// Same GEMM (bf16x3 pattern required), but C leaves as a bf16 (hi, lo) pair: C ~= C_hi + C_lo, [M][N] each.
extern "C" int sbev_gemm_bf16_tn_split(const uint16_t* A_hi, const uint16_t* A_lo, const uint16_t* B_hi, const uint16_t* B_lo,
                                       const float* bias, int M, int N, int K, uint16_t* C_hi, uint16_t* C_lo, void* stream) {
    SBEV_REQUIRE(A_hi && A_lo && B_hi && B_lo && C_hi && C_lo, SBEV_ERR_INVALID, "sbev_gemm_bf16_tn_split: null pointer");
    SBEV_REQUIRE(M > 0 && N > 0 && K > 0, SBEV_ERR_INVALID, "sbev_gemm_bf16_tn_split: bad sizes");
    SBEV_REQUIRE(K % GEMM_BK == 0 && N % 256 == 0, SBEV_ERR_UNSUPPORTED, "sbev_gemm_bf16_tn_split: needs K %% 64 == 0 and N %% 256 == 0");
    const void* ptrs[6] = {A_hi, A_lo, B_hi, B_lo, C_hi, C_lo};
    for (int i = 0; i < 6; ++i) SBEV_REQUIRE((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0, SBEV_ERR_INVALID, "sbev_gemm_bf16_tn_split: operands must be 16-byte aligned");
    SBEV_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0, SBEV_ERR_INVALID, "sbev_gemm_bf16_tn_split: bias must be 16-byte aligned");
    // CTA pairs only when forced.  gemm_impl 2: two 64 KB stages + double-buffered staging -- with K = 256 (four k-blocks per unit) the
    // extra cross-SM hop per pipeline stage costs more than the saved operand traffic (measured 59.7 us vs 55.3 us for the
    // parameter-generation GEMM); gemm_impl 5: THREE 64 KB stages + single staging (the single-CTA form has room for only two 96 KB
    // stages, so every k-block's load is exposed behind the MMA that frees its stage)
    const bool pair3 = get_option(OPT_GEMM_IMPL) == 5;
    const bool pair = get_option(OPT_GEMM_IMPL) == 2 || pair3;
    GemmMapsV2 mp;
    int rc = make_bf16_map(&mp.a_hi, A_hi, M, K, GEMM_BM);   if (rc) return rc;
    rc = make_bf16_map(&mp.a_lo, A_lo, M, K, GEMM_BM);       if (rc) return rc;
    rc = make_bf16_map(&mp.b_hi, B_hi, N, K, pair ? 128 : 256);   if (rc) return rc;
    rc = make_bf16_map(&mp.b_lo, B_lo, N, K, pair ? 128 : 256);   if (rc) return rc;
    rc = make_bf16_store_map(&mp.c_hi, C_hi, N, M);          if (rc) return rc;
    rc = make_bf16_store_map(&mp.c_lo, C_lo, N, M);          if (rc) return rc;
    mp.c = mp.c_hi;
    mp.hints = get_option(OPT_GEMM_L2_HINTS);
    const int num_sms = device_num_sms();
    const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM, n_tiles = N / 256;
    if (pair && num_sms >= 2) {
        const int num_units = ((m_tiles + 1) / 2) * n_tiles;
        const int clusters = num_units < num_sms / 2 ? num_units : num_sms / 2;
        if (pair3) {
            constexpr size_t smem3 = (size_t)3 * 2 * (GEMM_BM * GEMM_BK * 2 + 128 * GEMM_BK * 2) + 4 * 2 * 4096 + 1024;
            SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(gemm_bf16_tn_persistent_kernel<256, 3, true, 0, true, true>,
                                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            launch_pair(gemm_bf16_tn_persistent_kernel<256, 3, true, 0, true, true>, clusters, smem3, (cudaStream_t)stream,
                        mp, K / GEMM_BK, 1, m_tiles, n_tiles, num_units, bias, (float*)nullptr, M, N);
            return check_launch("sbev_gemm_bf16_tn_split(pair, 3 stages)");
        }
        SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(gemm_bf16_tn_persistent_kernel<256, 2, true, 0, true, true>,
                                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_PAIR_SPLIT_SMEM));
        launch_pair(gemm_bf16_tn_persistent_kernel<256, 2, true, 0, true, true>, clusters, GEMM_PAIR_SPLIT_SMEM, (cudaStream_t)stream,
                    mp, K / GEMM_BK, 1, m_tiles, n_tiles, num_units, bias, (float*)nullptr, M, N);
        return check_launch("sbev_gemm_bf16_tn_split(pair)");
    }
    if (get_option(OPT_GEMM_IMPL) == 4 && K == 4 * GEMM_BK && num_sms >= m_tiles) {
        // A-resident, 128-wide N tiles: a CTA keeps its M-tile's whole A operand (K = 256: 4 k-blocks x (hi, lo) = 128 KB) in
        // shared memory and streams only B (2 stages of 32 KB) -- a quarter less L2 -> SM traffic than re-fetching A per tile
        rc = make_bf16_map(&mp.b_hi, B_hi, N, K, 128);   if (rc) return rc;
        rc = make_bf16_map(&mp.b_lo, B_lo, N, K, 128);   if (rc) return rc;
        constexpr size_t smem_ares = (size_t)4 * 2 * (GEMM_BM * GEMM_BK * 2) + (size_t)2 * 2 * (128 * GEMM_BK * 2) + 4 * 2 * 4096 + 1024;
        SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(gemm_bf16_tn_persistent_kernel<128, 2, true, 4, true>,
                                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ares));
        const int grid_ares = (num_sms / m_tiles) * m_tiles;
        launch_pdl(gemm_bf16_tn_persistent_kernel<128, 2, true, 4, true>, dim3(grid_ares), dim3(GEMM_THREADS), smem_ares, (cudaStream_t)stream,
                   mp, K / GEMM_BK, 1, m_tiles, N / 128, m_tiles * (N / 128), bias, nullptr, M, N);
        return check_launch("sbev_gemm_bf16_tn_split(a-resident)");
    }
    const int num_tiles = m_tiles * n_tiles;
    const int grid = num_tiles < num_sms ? num_tiles : num_sms;
    constexpr size_t smem = (size_t)2 * 2 * (GEMM_BM * GEMM_BK * 2 + 256 * GEMM_BK * 2) + 4 * 2 * 4096 + 1024;
    SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(gemm_bf16_tn_persistent_kernel<256, 2, true, 0, true>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_pdl(gemm_bf16_tn_persistent_kernel<256, 2, true, 0, true>, dim3(grid), dim3(GEMM_THREADS), smem, (cudaStream_t)stream,
               mp, K / GEMM_BK, 1, m_tiles, n_tiles, num_tiles, bias, nullptr, M, N);
    return check_launch("sbev_gemm_bf16_tn_split");
}
