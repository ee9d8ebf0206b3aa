// Cross-GPU exchange for the query-/frame-sharded decoder: rows a rank has just produced are stored
// into the same offset of every peer's buffer over NVLink / NVSwitch (peer memory mapped into this process: CUDA IPC or
// torch symmetric memory), followed -- in the SAME kernel -- by an all-ranks barrier on flag words that live in the
// peers' memory.  No NCCL call, no host round trip; the launch is a plain kernel and therefore graph-capturable.
//
// The reference has no counterpart (its only multi-GPU mode is DDP, /root/reference/train.py:90-131); the data that moves
// is what the sharded form of its decoder layer needs between ranks: sample points + scale weights of all queries
// (models/sparsebev_transformer.py:279-300), and the refined boxes / query features every rank needs for the next layer's
// self-attention keys (models/sparsebev_transformer.py:93,169).
//
// Protocol (one flag array of SBEV_MAX_PEERS words per rank, all zero before the first call; every rank issues the same
// sequence of exchanges):
//   1. every CTA copies its share of the segments to all peers (16-byte stores), fences at system scope and bumps a local
//      arrival counter;
//   2. the last CTA to arrive takes epoch e = ++ctl.epoch, stores e into slot [my rank] of EVERY peer's flag array
//      (st.release.sys) and spins (ld.acquire.sys) until all slots of its OWN flag array have reached e: all peers' rows
//      have landed here, and all of them have finished reading what this rank will overwrite next;
//   3. kernel completion is the barrier for everything that follows on the stream (incl. programmatic dependents, whose
//      griddepcontrol.wait returns only when this grid has completed and flushed).
// A peer that never arrives (crashed rank) does not hang the GPU: the spin gives up after ~4 s, records it in ctl.status,
// and every later exchange returns immediately; the host reads ctl.status after synchronising.
#include "common.cuh"

namespace sbev {

constexpr int PEER_MAX_SEGS = 8;

struct PeerParams {
    const float* src[PEER_MAX_SEGS];                    // this rank's rows (inside its own buffer)
    float* dst[PEER_MAX_SEGS][SBEV_MAX_PEERS];          // the same rows inside every rank's buffer (entry [rank] unused)
    long long n[PEER_MAX_SEGS];                         // 4-byte words per segment
    int head[PEER_MAX_SEGS];                            // words before the first 16-byte boundary (src and every dst share the misalignment)
    int nseg, n_peers, rank;
    uint32_t* flags[SBEV_MAX_PEERS];                    // flags[w] = rank w's flag array [SBEV_MAX_PEERS]
    uint32_t* ctl;                                      // local: [0] epoch, [1] arrival counter, [2] status (1 = a wait timed out)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256)
peer_exchange_kernel(const __grid_constant__ PeerParams prm) {
    __shared__ int s_last;
    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x;
    const long long gtid = (long long)blockIdx.x * blockDim.x + tid, gstride = (long long)gridDim.x * blockDim.x;
    for (int s = 0; s < prm.nseg; ++s) {
        const float* src = prm.src[s];
        const long long n = prm.n[s], head = prm.head[s], n4 = (n - head) >> 2, tail0 = head + 4 * n4;
        for (long long i = gtid; i < n4; i += gstride) {                 // 16-byte body
            const float4 v = __ldcg(reinterpret_cast<const float4*>(src + head) + i);      // produced by the preceding kernel: bypass L1
#pragma unroll
            for (int w = 0; w < SBEV_MAX_PEERS; ++w)
                if (w < prm.n_peers && w != prm.rank) reinterpret_cast<float4*>(prm.dst[s][w] + head)[i] = v;
        }
        if (blockIdx.x == 0 && tid < 8) {                                // up to 3 leading + 3 trailing words
            const long long i = tid < 4 ? tid : tail0 + (tid - 4);
            if ((tid < 4 && i < head) || (tid >= 4 && i < n)) {
                const float v = __ldcg(src + i);
                for (int w = 0; w < prm.n_peers; ++w)
                    if (w != prm.rank) prm.dst[s][w][i] = v;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        // one system-scope fence per CTA: cumulative over the whole CTA's stores (ordered before it by the barrier above),
        // so they are performed at the peers before the arrival below -- and hence before the last CTA's flag stores
        __threadfence_system();
        s_last = atomicAdd(prm.ctl + 1, 1u) == gridDim.x - 1;
        __threadfence();                                 // acquire side of the arrival chain for the last CTA
    }
    __syncthreads();
    if (!s_last) return;
    // ---- last CTA: signal every peer (release), then wait for every peer (acquire)
    const uint32_t epoch = prm.ctl[0] + 1;
    const bool dead = prm.ctl[2] != 0;
    __syncthreads();
    if (tid == 0) { prm.ctl[0] = epoch; prm.ctl[1] = 0; }
    if (tid < prm.n_peers && tid != prm.rank) {
        st_release_sys(prm.flags[tid] + prm.rank, epoch);
        if (!dead) {
            const uint32_t* mine = prm.flags[prm.rank] + tid;
            const unsigned long long t0 = global_ns();
            while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
                if (global_ns() - t0 > 4000000000ull) { prm.ctl[2] = 1; break; }
            }
        }
    }
}

}  // namespace sbev

using namespace sbev;

extern "C" int sbev_peer_exchange(const sbev_peer_segment* segs, int nseg, int n_peers, int rank,
                                  uint32_t* const* flags, uint32_t* ctl, void* stream) {
    SBEV_REQUIRE(n_peers >= 1 && n_peers <= SBEV_MAX_PEERS && rank >= 0 && rank < n_peers, SBEV_ERR_INVALID,
                 "sbev_peer_exchange: rank %d / %d peers outside [1,%d]", rank, n_peers, SBEV_MAX_PEERS);
    SBEV_REQUIRE(nseg >= 0 && nseg <= PEER_MAX_SEGS && (nseg == 0 || segs != nullptr), SBEV_ERR_INVALID,
                 "sbev_peer_exchange: between 0 and %d segments (got %d)", PEER_MAX_SEGS, nseg);
    SBEV_REQUIRE(flags != nullptr && ctl != nullptr, SBEV_ERR_INVALID, "sbev_peer_exchange: null flag / control pointer");
    PeerParams prm{};
    prm.nseg = nseg; prm.n_peers = n_peers; prm.rank = rank; prm.ctl = ctl;
    long long total = 0;
    for (int w = 0; w < n_peers; ++w) {
        SBEV_REQUIRE(flags[w] != nullptr && (reinterpret_cast<uintptr_t>(flags[w]) & 3) == 0, SBEV_ERR_INVALID, "sbev_peer_exchange: flags[%d] null or misaligned", w);
        prm.flags[w] = flags[w];
    }
    for (int s = 0; s < nseg; ++s) {
        SBEV_REQUIRE(segs[s].bytes >= 0 && (segs[s].bytes & 3) == 0, SBEV_ERR_INVALID, "sbev_peer_exchange: segment %d size %lld is not a multiple of 4 bytes", s, (long long)segs[s].bytes);
        if (segs[s].bytes == 0) continue;
        const uintptr_t sa = reinterpret_cast<uintptr_t>(segs[s].src);
        SBEV_REQUIRE(sa != 0 && (sa & 3) == 0, SBEV_ERR_INVALID, "sbev_peer_exchange: segment %d source null or not 4-byte aligned", s);
        prm.src[s] = reinterpret_cast<const float*>(segs[s].src);
        prm.n[s] = segs[s].bytes / 4;
        const long long head = (long long)((16 - (sa & 15)) & 15) / 4;
        prm.head[s] = (int)(head < prm.n[s] ? head : prm.n[s]);
        for (int w = 0; w < n_peers; ++w) {
            if (w == rank) continue;
            const uintptr_t da = reinterpret_cast<uintptr_t>(segs[s].dst[w]);
            SBEV_REQUIRE(da != 0 && (da & 15) == (sa & 15), SBEV_ERR_INVALID,
                         "sbev_peer_exchange: segment %d destination %d null or not congruent to the source modulo 16 bytes", s, w);
            prm.dst[s][w] = reinterpret_cast<float*>(segs[s].dst[w]);
        }
        total += prm.n[s];
    }
    long long grid = (total + 4095) / 4096;               // 4 x 16 bytes per thread
    if (grid < 1) grid = 1;
    if (grid > 64) grid = 64;
    launch_pdl(peer_exchange_kernel, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, prm);
    return check_launch("sbev_peer_exchange");
}
