/* sparsebev_b200 -- C ABI of the B200-native SparseBEV decoder hot path.
 *
 * Plain C: raw DEVICE pointers (fp32 unless stated), sizes, and the CUDA stream to enqueue on.
 * No torch / ATen types.  Every function
 *   - only ENQUEUES work on `stream` (a cudaStream_t passed as void*; NULL = legacy stream 0),
 *     never synchronises the host, never allocates or frees device memory;
 *   - is re-entrant and thread-safe (no global mutable state except the last-error string,
 *     which is thread-local);
 *   - returns SBEV_OK or a negative SBEV_ERR_* code; sbev_last_error() describes the failure.
 * The caller owns and allocates every buffer (as in the reference, where the C++ host
 * functions allocate the outputs with at::zeros and hand raw pointers to the launchers:
 * /root/reference/models/csrc/msmv_sampling/msmv_sampling.cpp:136-148).
 *
 * Each entry point cites the reference interface it replaces (paths under /root/reference).
 */
#ifndef SPARSEBEV_B200_H_
#define SPARSEBEV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBEV_OK               0
#define SBEV_ERR_INVALID     -1   /* bad argument (null pointer, non-positive size, P > 32, ...) */
#define SBEV_ERR_UNSUPPORTED -2   /* shape outside what the kernels implement */
#define SBEV_ERR_CUDA        -3   /* CUDA runtime reported an error at launch */

#define SBEV_MAX_LEVELS 5          /* c2..c6, reference: MSMVSamplingC23456 */
#define SBEV_MAX_POINTS 32         /* reference: MAX_POINT, msmv_sampling.cpp:3 / :125 */
#define SBEV_MAX_PEERS 8           /* GPUs of one NVSwitch node that sbev_sampling4d_scatter_fwd stores to */

/* version / diagnostics */
int         sbev_abi_version(void);
const char* sbev_last_error(void);
/* Kernel-variant selectors for A/B measurements (process-wide; defaults from the environment):
 *   "gemm_impl"      0 = default: the fp32-output GEMM (sbev_gemm_bf16_tn, bf16x3, N % 256 == 0) runs as CTA pairs -- clusters of 2,
 *                    tcgen05 cta_group::2, 256 x 256 units, each CTA loads its A rows and half of the B tile (a third less
 *                    L2->SM operand traffic: 49 -> 39.5 us for the out-projection) -- while the bf16-pair-output GEMM
 *                    (sbev_gemm_bf16_tn_split, K = 256) stays on single CTAs (pairs measured slower there);
 *                    2 = CTA pairs for both; 3 = single CTAs for both; 1 = single CTAs + A-resident schedule for
 *                    small-K / many-N GEMMs (measured slower)
 *                    5 = as 0, but the bf16-pair-output GEMM runs as CTA pairs with a THREE-stage operand ring (64 KB stages, single
 *                    epilogue staging) instead of single CTAs with two 96 KB stages
 *   "mix_impl"       0 = mma.sync bf16x3 (default), 1 = fp32 FFMA
 *   "sasa_impl"      0 = mma.sync bf16x3 (default), 1 = fp32 FFMA
 *   "dense_impl"     0 = mma.sync bf16x3 chain with TMA-streamed weights (default), 1 = fp32 FFMA chain
 *   "dense_cluster"  0 = every CTA streams its own weight tiles (default); 2 / 4 / 8 = that many CTAs (row groups) form a
 *                    cluster and share every weight tile by TMA multicast (8 measured slower: lock step)
 *   "dense_nsplit"   0 = off; 2 / 4 = N-split cluster chain: that many CTAs form a cluster that owns 16 / 32 rows and SPLIT
 *                    every layer's output features (each streams 1/2 / 1/4 of the weights), exchanging the layer outputs
 *                    through distributed shared memory; chains it cannot express (an inner layer wider than 512 ...) take
 *                    the "dense_impl" 0 kernel
 *   "sasa_kq"        key splits per CTA of the attention core: 0 = automatic (8 when the launch is below one wave, i.e. a query shard;
 *                    else 4), 4 / 8 = forced.  The result of a query depends on the split (partial softmax merge order): <= 1e-6 relative
 *   "dense_pack"     1 = chains stream the pre-tiled weight copy (sbev_dense_layer.W_pack) with one bulk copy per stage when the caller
 *                    provides it (default); 0 = always the tensor-map path
 *   "dense_vec4"     1 = 16-byte vectorised row epilogue / operand staging in the chain kernels (default), 0 = scalar
 *   "dense_fuse_points" 1 = sbev_dense_chain_points_fwd computes the sample points in the chain's epilogue, 0 = it runs the
 *                    chain and then sample_points_kernel (default: measured 4 us faster per layer -- the fused epilogue
 *                    delays the parameter GEMM that waits on the same launch, the separate kernel overlaps with it)
 *   "pdl"            1 = hot-path kernels are launched with programmatic stream serialization (default): each kernel runs its
 *                    global-memory-free prologue while its predecessor drains, then griddepcontrol.wait; 0 = plain launches
 *   "legacy_rotation" 0 = v1.0.0 box convention (default); 1 = the sample-point rotation of checkpoints whose `version` is 'v0.17.1'
 *                    (models/utils.py:66-71: rotation_3d_in_axis turns the other way; toggled like the reference's global
 *                    VERSION.name, val.py:128-129) -- affects sbev_sample_points_fwd and sbev_dense_chain_points_fwd
 *   "dense_ws"       which chains the host mirror (ops.dense_chain / dense_chain_reduce) sends to the weights-stationary cluster kernel
 *                    (sbev_dense_chain_ws_*): 0 = none, 1 = every chain it can express (default), 2 = only chains of <= 256 rows
 *   "dense_ws_groups" independent 16-row groups per CTA of that kernel: 0 = automatic (two when their buffers fit; default), 1, 2
 *   "gather_variant" 0 = 16 lanes/point, all levels in flight; 1 = 16 lanes/point, two levels at a time, 3 CTAs/SM;
 *                    2 = 8 lanes/point x 8 channels, two levels at a time; 3 = 2 + the next level pair's lines are prefetched
 *                    into L2; 4 = 2 + level blocks in which no point of the warp has a live tap are skipped by a warp-uniform
 *                    branch (~30 % of the taps on a real camera rig fall outside every view); 5 = 4 at 3 CTAs/SM;
 *                    6 = 4 with ONE level (8 loads per lane) at a time, 64 registers, 4 CTAs/SM (default: 54.0 -> 44.5 us at r50-T8) */
int         sbev_set_option(const char* name, int value);
/* current value of an option (the environment / built-in default until sbev_set_option overrides it); -1 = unknown name */
int         sbev_get_option(const char* name);

/* ---------------------------------------------------------------------------------------------
 * msmv_sampling forward.
 * Replaces: ms_deform_attn_cuda_c2345_forward / _c23456_forward (msmv_sampling.cpp:98-210) and the
 * launchers ms_deformable_im2col_cuda_c2345/_c23456 (msmv_sampling_forward.cu:269-333), generalised
 * to any 1 <= L <= 5.
 *   feats  HOST array of L DEVICE pointers; feats[l] = [Bp, N, H_l, W_l, C] channel-last, contiguous
 *   hw     HOST array of 2*L ints: H_0, W_0, H_1, W_1, ...
 *   loc    [Bp, Q, P, 3]  (u, v in normalised image coords, view index / (N-1))
 *   w      [Bp, Q, P, L]  scale weights
 *   out    [Bp, Q, C, P]  every element is written (no pre-zeroing needed)
 * Semantics are the reference CUDA kernel's: view = round(z*(N-1)); per level bilinear with
 * align_corners=True and zero padding; sum over levels of tap*weight.  A view index outside
 * [0, N) (undefined behaviour in the reference) contributes zero here.
 */
int sbev_msmv_fwd(const float* const* feats, const int* hw, int L,
                  const float* loc, const float* w,
                  int Bp, int N, int C, int Q, int P,
                  float* out, void* stream);

/* msmv_sampling backward.
 * Replaces: ms_deform_attn_cuda_c2345_backward / _c23456_backward (msmv_sampling.cpp:212-360) and
 * ms_deformable_col2im_cuda_* (msmv_sampling_backward.cu:363-448).
 *   grad_out    [Bp, Q, C, P]
 *   grad_feats  HOST array of L DEVICE pointers, same shapes as feats   (zero-filled here, then accumulated)
 *   grad_loc    [Bp, Q, P, 3]   (component 2, the view coordinate, gets exactly 0 as in the reference)
 *   grad_w      [Bp, Q, P, L]
 */
int sbev_msmv_bwd(const float* grad_out, const float* const* feats, const int* hw, int L,
                  const float* loc, const float* w,
                  int Bp, int N, int C, int Q, int P,
                  float* const* grad_feats, float* grad_loc, float* grad_w, void* stream);

/* Deterministic msmv_sampling backward: same contract as sbev_msmv_bwd, but grad_feats is produced
 * WITHOUT floating-point atomics -- the reference's scatter (one atomicAdd per corner and channel,
 * msmv_sampling_backward.cu:29-224) is inverted into a per-pixel segmented reduction that sums the contributions of a
 * pixel in ascending (point, corner) order, so repeated runs are bit-identical.  C must be 64.
 *   workspace        DEVICE scratch owned by the caller, at least sbev_msmv_bwd_det_workspace(...) bytes, 4-byte aligned;
 *                    contents are undefined afterwards.  grad_feats need NOT be zeroed (every pixel row is written once).
 * sbev_msmv_bwd_det_workspace returns the byte count, or -1 for invalid sizes.
 */
long long sbev_msmv_bwd_det_workspace(const int* hw, int L, int Bp, int N, int Q, int P);
int sbev_msmv_bwd_det(const float* grad_out, const float* const* feats, const int* hw, int L,
                      const float* loc, const float* w,
                      int Bp, int N, int C, int Q, int P,
                      float* const* grad_feats, float* grad_loc, float* grad_w,
                      void* workspace, long long workspace_bytes, void* stream);

/* The integer sample indices the forward kernel derives from `loc` -- for bit-exact index parity
 * tests (same device code path as sbev_msmv_fwd: msmv_sampling_forward.cu:110,123-126,33-36).
 *   view [Bp,Q,P] int32;  y0, x0, inside [Bp,Q,P,L] int32
 */
int sbev_msmv_indices(const int* hw, int L, const float* loc, int Bp, int N, int Q, int P,
                      int32_t* view, int32_t* y0, int32_t* x0, int32_t* inside, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused adaptive spatio-temporal sampling: motion warp -> projection to T x N views -> first-valid
 * view pick -> multi-scale gather, written straight in the layout AdaptiveMixing consumes.
 * Replaces: sampling_4d (models/sparsebev_sampling.py:27-130) + the warp in
 * SparseBEVSampling.inner_forward (models/sparsebev_transformer.py:286-295) + msmv_sampling.
 *   feats         HOST array of L DEVICE pointers.  Element (b,t,g,view,y,x,c) of level l lives at
 *                 feats[l] + (b*T+t)*stride_bt[l] + g*stride_g[l] + view*stride_v[l] + (y*W+x)*stride_px[l] + c
 *                 (strides in floats, HOST arrays of L int64).  This covers both the reference's
 *                 regrouped [B*T*G,N,H,W,C] layout and an un-regrouped NHWC [B,T*N,H,W,G*C] map.
 *   points        [B, Q, G*P, 3]   lidar-frame sample points BEFORE the motion warp (metres)
 *   velocity      [B, Q, 2]        query_bbox[..., 8:10]
 *   time_diff     [B, T]
 *   lidar2img     [B, T*N, 4, 4]   row-major
 *   scale_w       [B, Q, G, P, L]  softmaxed scale weights (not expanded over T)
 *   out           [B, Q, G, T*P, C]   point index = t*P + p
 *   loc_out       optional (may be NULL) [B*T*G, Q, P, 3]  the (u, v, view/(N-1)) handed to the op
 * The weight row used for (t,g) is that of group ((t*G+g)/T)%G -- the reference's (b,t,g)/(b,g,t)
 * flattening mismatch (sparsebev_sampling.py:112-119) is reproduced on purpose.
 */
int sbev_sampling4d_fwd(const float* const* feats, const int* hw, int L,
                        const int64_t* stride_bt, const int64_t* stride_g,
                        const int64_t* stride_v, const int64_t* stride_px,
                        const float* points, const float* velocity, const float* time_diff,
                        const float* lidar2img, const float* scale_w,
                        int B, int T, int G, int N, int C, int Q, int P,
                        float image_h, float image_w, float eps,
                        float* out, float* loc_out, void* stream);

/* Frame-window form of sbev_sampling4d_fwd, for the frame-sharded decoder (partitioning B: every GPU keeps
 * the feature maps of the frames its backbone produced and samples only those).  `feats` hold the Tl frames
 * [t0, t0+Tl) -- element (b, t, ...) at (b*Tl + (t - t0))*stride_bt[l] + ... -- while time_diff, lidar2img and the
 * (t,g)->weight-group pairing keep indexing all T frames.  out [B, Q, G, Tl*P, C] (point index (t-t0)*P + p),
 * loc_out optional [B*Tl*G, Q, P, 3].  The velocity of query (b,q) is read at velocity + (b*Q+q)*ld_vel: 2 for a packed
 * [B,Q,2] tensor, 10 to read it in place from query_bbox + 8 (no slice copy).  t0 = 0, Tl = T, ld_vel = 2 is exactly
 * sbev_sampling4d_fwd.  Every sample is computed
 * independently of the window, so the union of the windows is bit-identical to the unsharded call.
 */
int sbev_sampling4d_window_fwd(const float* const* feats, const int* hw, int L,
                               const int64_t* stride_bt, const int64_t* stride_g,
                               const int64_t* stride_v, const int64_t* stride_px,
                               const float* points, const float* velocity, int ld_vel, const float* time_diff,
                               const float* lidar2img, const float* scale_w,
                               int B, int T, int t0, int Tl, int G, int N, int C, int Q, int P,
                               float image_h, float image_w, float eps,
                               float* out, float* loc_out, void* stream);

/* Fused gather + all-gather for the frame-sharded decoder: the window's rows are stored straight into n_out FULL-SIZE
 * buffers [B, Q, G, T*P, C] (point index t*P + p), `outs` = HOST array of n_out DEVICE pointers -- this GPU's buffer
 * and the peer GPUs' buffers mapped into this process (CUDA IPC / symmetric memory; the stores travel over NVLink).
 * After every rank has run its window and a cross-GPU barrier, each buffer holds exactly what the unsharded
 * sbev_sampling4d_fwd writes -- no NCCL call and no re-layout copy in between.  n_out <= SBEV_MAX_PEERS.
 */
int sbev_sampling4d_scatter_fwd(const float* const* feats, const int* hw, int L,
                                const int64_t* stride_bt, const int64_t* stride_g,
                                const int64_t* stride_v, const int64_t* stride_px,
                                const float* points, const float* velocity, int ld_vel, const float* time_diff,
                                const float* lidar2img, const float* scale_w,
                                int B, int T, int t0, int Tl, int G, int N, int C, int Q, int P,
                                float image_h, float image_w, float eps,
                                float* const* outs, int n_out, float* loc_out, void* stream);

/* Owner form of the fused gather, for the query- AND frame-sharded decoder (B must be 1): this GPU holds the
 * feature maps of frames [t0, t0+Tl) and samples them for ALL Q queries, but query q is mixed by rank q / q_per_rank only,
 * so each 256 B row is stored exactly once -- into outs[q / q_per_rank], that rank's [q_per_rank, G, T*P, C] buffer (peer
 * memory over NVLink, or this GPU's own), at row (q % q_per_rank).  After every rank has run its window and a
 * sbev_peer_exchange barrier, rank r holds the rows sbev_sampling4d_fwd would have written for its queries: the all-to-all
 * of the sampled features (models/sparsebev_sampling.py:112-128 regroups them per query) is fused into the gather's stores.
 * `outs` = HOST array of n_ranks DEVICE pointers, n_ranks * q_per_rank >= Q, n_ranks <= SBEV_MAX_PEERS. */
int sbev_sampling4d_owner_fwd(const float* const* feats, const int* hw, int L,
                              const int64_t* stride_bt, const int64_t* stride_g,
                              const int64_t* stride_v, const int64_t* stride_px,
                              const float* points, const float* velocity, int ld_vel, const float* time_diff,
                              const float* lidar2img, const float* scale_w,
                              int T, int t0, int Tl, int G, int N, int C, int Q, int P,
                              float image_h, float image_w, float eps,
                              float* const* outs, int n_ranks, int q_per_rank, void* stream);

/* Cross-GPU exchange + barrier in ONE kernel (no counterpart in the reference, whose only multi-GPU mode is DDP; this is
 * what the sharded decoder layer moves between ranks -- sample points / scale weights, refined boxes, query features).
 * Segment s: `bytes` (multiple of 4) are copied from `src` (this rank's freshly produced rows, inside its own buffer) to
 * dst[w] for every rank w != rank -- the same rows inside rank w's buffer (same address modulo 16), mapped into this process
 * (CUDA IPC / torch symmetric memory; stores travel over NVLink).  Then an all-ranks barrier: flags[w] is rank w's array of SBEV_MAX_PEERS
 * uint32 words (zero before the first call); the kernel stores its epoch into word [rank] of every peer's array
 * (st.release.sys) and waits until all words of its own array reached it (ld.acquire.sys).  When the kernel completes, every
 * peer's segments of the same exchange have landed in this rank's buffers, and every peer has completed all stream work it
 * enqueued before ITS call.  nseg = 0 is a pure barrier (used after sbev_sampling4d_owner_fwd).
 *   ctl   DEVICE, 3 uint32 owned by this rank, zero-initialised: epoch, arrival counter, status.  status becomes 1 when a
 *         wait gave up after ~4 s (a peer died); later calls then return without waiting.  Every rank must issue the same
 *         sequence of exchanges.  Plain kernel launch: capturable in a CUDA graph, never synchronises the host. */
typedef struct sbev_peer_segment {
    const void* src;
    void* dst[SBEV_MAX_PEERS];
    int64_t bytes;
} sbev_peer_segment;
int sbev_peer_exchange(const sbev_peer_segment* segs, int nseg, int n_peers, int rank,
                       uint32_t* const* flags, uint32_t* ctl, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Small dense building block: y = epilogue(x @ W^T), row-major fp32, with the weight given
 * PRE-TRANSPOSED as Wt[K][ldw] (= W^T, zero-padded to ldw >= N, ldw % 4 == 0; the host mirror caches
 * this copy per parameter).  epilogue: + bias -> (+ residual if SBEV_DENSE_RES_PRE_LN) -> LayerNorm over
 * N (eps 1e-5, if ln_w/ln_b given) -> ReLU (if SBEV_DENSE_RELU) -> (+ residual otherwise).
 * Replaces the nn.Linear / nn.LayerNorm / ReLU chains of position_encoder, attention in/out
 * projections, FFN, cls/reg branches, gen_tau, sampling_offset, scale_weights
 * (models/sparsebev_transformer.py:113-144,166-176,203,262-263) and norm1/norm3.
 *   x [M,K] with row stride ldx >= K floats, bias [N] or NULL, ln_w/ln_b [N] or both NULL,
 *   residual [M,N] or NULL, y [M,N]
 */
#define SBEV_DENSE_RELU        1
#define SBEV_DENSE_RES_PRE_LN  2
#define SBEV_DENSE_REFINE      4   /* chain only: box refinement epilogue, see sbev_dense_chain_fwd */
#define SBEV_DENSE_WIDE_CTA    8   /* chain only, a HINT on the FIRST layer: 16 rows per CTA instead of 8 (half as many CTAs stream the
                                      weights) -- for two chains that run side by side on two streams (cls || reg) so that both fit on the
                                      SMs at once; honoured by the streaming tensor-core chain when every K, N <= 256, ignored otherwise */
int sbev_dense_fwd(const float* x, int ldx, const float* Wt, int ldw, const float* bias,
                   const float* ln_w, const float* ln_b, const float* residual,
                   int M, int K, int N, int flags, float* y, void* stream);

/* A CHAIN of 1..6 such layers in ONE kernel: a CTA owns 8 full rows through all layers (intermediate
 * activations stay in shared memory), weights of all layers stream through a TMA / cp.async.bulk + mbarrier ring;
 * the matmuls run on the tensor cores (mma.sync, bf16x3 split, fp32 accumulate) when W_hi/W_lo are given.
 * Layer i consumes layer i-1's output (K_i == N_{i-1}); every layer may additionally store its result
 * (y != NULL, row stride ldy); the last layer must.  With SBEV_DENSE_REFINE on the last layer the epilogue is
 * refine_bbox + velocity rescale (models/sparsebev_transformer.py:155-160,179-183): columns 0..2 become
 * sigmoid(v + inverse_sigmoid(refine_proposal[row, n])), columns >= 8 are divided by time_diff[b,1] (if T > 1).
 * Used for: position_encoder (2 layers), in_proj+gen_tau (1, concatenated), out_proj+norm1 -> sampling heads (2),
 * FFN+norm3 (2), cls_branch (3), reg_branch+refine (3).
 */
typedef struct sbev_dense_layer {
    const float* Wt;        /* [K][ldw] pre-transposed weight, zero padded */
    int ldw, K, N;
    const float* bias;      /* [N] or NULL */
    const float* ln_w;      /* [N] or NULL (with ln_b) */
    const float* ln_b;
    const float* residual;  /* [M][N] or NULL */
    int flags;              /* SBEV_DENSE_* */
    float* y;               /* [M][ldy] or NULL */
    int ldy;
    const uint16_t* W_hi;   /* tensor-core path: bf16 (hi, lo) split of W in the nn.Linear layout [N][Kpad], zero padded */
    const uint16_t* W_lo;   /*   (Kpad = K rounded up to 64); NULL selects the fp32 FFMA path, which needs Wt instead   */
    int Kpad;
    uint16_t* y_hi;         /* optional: bf16 (hi, lo) split of the stored output, [M][ldy] each (feeds the tensor-core   */
    uint16_t* y_lo;         /*   kernels that follow: attention core, parameter GEMM); both or neither                    */
    const uint16_t* W_pack; /* optional: the same (hi, lo) weights PRE-TILED in streaming order -- [ceil(N/128)][Kpad/64][hi|lo][128 rows][64 k]
                               bf16, rows zero-padded to a multiple of 128, each 128-byte row stored with its 16-byte chunks XOR-swizzled
                               by (row & 7) (the layout ldmatrix reads, = TMA's 128-byte swizzle).  Every pipeline stage is then ONE contiguous
                               32 KB cp.async.bulk instead of two tensor-map box loads of 128 separate 128-byte rows.  128-byte aligned. */
} sbev_dense_layer;
int sbev_dense_chain_fwd(const float* x, int ldx, int M, int n_layers, const sbev_dense_layer* layers,
                         const float* refine_proposal, const float* refine_time_diff, int refine_Q, int refine_T,
                         void* stream);

/* sbev_dense_chain_fwd whose LAST layer (a plain Linear: no LayerNorm / ReLU / residual, output y required) holds, per row,
 * GP*3 sampling offsets starting at column off_col and GP*L scale logits starting at column log_col; its epilogue also
 * turns them into the sample points and the per-level softmax weights of sbev_sample_points_fwd (same arithmetic, same
 * bits): points [M][GP][3], scale_w [M][GP][L].  Fuses SparseBEVSampling's two Linear heads, make_sample_points and the
 * softmax (models/sparsebev_transformer.py:279-283,298-299; models/sparsebev_sampling.py:8-24) into the launch that also
 * applies the attention out-projection + norm1.  Kernel variants without the fused epilogue run the chain, then
 * sbev_sample_points_fwd. */
int sbev_dense_chain_points_fwd(const float* x, int ldx, int M, int n_layers, const sbev_dense_layer* layers,
                                const float* query_bbox, const float* pc_range, int GP, int L, int off_col, int log_col,
                                float* points, float* scale_w, void* stream);

/* The same chain fed by the split-K partials of the preceding GEMM: its input rows are
 *     x_out[row] = LayerNorm(sum_z partial[z][row] + bias + residual[row])        (ln_w/ln_b NULL: no LayerNorm)
 * with partial [nsplit][M][K0], K0 = layers[0].K -- i.e. sbev_reduce_ln_fwd fused into the chain's prologue (one launch and
 * one round trip of the [M, K0] activations less).  x_out [M][K0] is also stored (it is the residual of a later layer:
 * AdaptiveMixing's out_proj + norm2 feeding the FFN, models/sparsebev_transformer.py:171-175).  Falls back to the two
 * separate kernels when the tensor-core chain cannot take the layers or K0 is not 128 / 256.
 */
int sbev_dense_chain_reduce_fwd(const float* partial, int nsplit, const float* bias, const float* residual,
                                const float* ln_w, const float* ln_b, float* x_out,
                                int M, int n_layers, const sbev_dense_layer* layers,
                                const float* refine_proposal, const float* refine_time_diff, int refine_Q, int refine_T,
                                void* stream);

/* WEIGHTS-STATIONARY form of the two chain entry points above (same layers, same epilogues, same results to fp32 round-off;
 * replaces the same reference lines): a cluster of 8 CTAs owns a block of rows and splits every layer's OUTPUT FEATURES 8 ways, each
 * CTA keeping its slice of ALL layers' weights resident in shared memory (one bulk copy, issued under the previous kernel's tail),
 * layer outputs exchanged through distributed shared memory (bulk shared->peer copies counted on the receiver's mbarrier).  No CTA streams a whole chain's weights, so the time scales with the
 * row count instead of being pinned at 9-20 us by per-SM ingest (csrc/dense_ws.cu).
 *   blob: the chain's weights pre-sliced per CTA rank c = 0..7, rank stride blob_stride_bytes (>= sbev_dense_chain_ws_blob_bytes,
 *   multiple of 128, 128-byte aligned).  Rank c, layer i (layers back to back): [Kpad_i / 64][hi | lo][SW_i rows][64 k] bf16 with
 *   SW_i = ceil(N_i / 8) rounded up to 8, row j = output feature c * SW_i + j (zero rows beyond N_i), every 128-byte row stored with its
 *   16-byte chunks XOR-swizzled by (j & 7).  Only K, Kpad, N, bias, ln_w, ln_b, residual, flags, y, ldy, y_hi, y_lo of the layers are used.
 * Limits: Kpad <= 512; a layer that feeds another one or ends in LayerNorm needs N <= 512, N % 4 == 0 and 16-byte aligned operands;
 * the blob plus one 16-row group's buffers must fit 226 KB of shared memory (SBEV_ERR_UNSUPPORTED otherwise -- callers then use the streaming chain);
 * the reduce prologue handles K0 == 256 and needs a last layer that ends in LayerNorm.
 */
long long sbev_dense_chain_ws_blob_bytes(int n_layers, const sbev_dense_layer* layers);
/* Diagnostics: while `stamps` (device memory, 64 x 8 bytes per CTA of the largest launch, zero-filled by the caller) is set, thread 0 of
 * every CTA records clock64() at the kernel's phase boundaries (entry, barriers up, predecessor done, cluster up, rows staged, weights
 * landed, then per layer: MMAs done / slice sent / slices received / rows finished, ..., work done, exit).  NULL switches it off. */
int sbev_dense_chain_ws_debug(unsigned long long* stamps);
int sbev_dense_chain_ws_fwd(const float* x, int ldx, int M, int n_layers, const sbev_dense_layer* layers,
                            const uint16_t* blob, long long blob_stride_bytes,
                            const float* refine_proposal, const float* refine_time_diff, int refine_Q, int refine_T,
                            void* stream);
int sbev_dense_chain_ws_reduce_fwd(const float* partial, int nsplit, const float* bias, const float* residual,
                                   const float* ln_w, const float* ln_b, float* x_out,
                                   int M, int n_layers, const sbev_dense_layer* layers,
                                   const uint16_t* blob, long long blob_stride_bytes,
                                   const float* refine_proposal, const float* refine_time_diff, int refine_Q, int refine_T,
                                   void* stream);

/* Sampling head epilogue: box decode + offset scaling + yaw rotation + softmax over levels.
 * Replaces make_sample_points (models/sparsebev_sampling.py:8-24), decode_bbox (models/bbox/utils.py:63-77),
 * rotation_3d_in_axis (models/utils.py:49-84) and the softmax at sparsebev_transformer.py:298-299.
 *   query_bbox [BQ,10]; offset rows of GP*3 floats (row stride ld_off); scale_logits rows of GP*L floats (row
 *   stride ld_log) -- both may be column blocks of one concatenated Linear output; pc_range HOST[6]
 *   points [BQ, GP, 3]; scale_w [BQ, GP, L]
 */
int sbev_sample_points_fwd(const float* query_bbox, const float* offset, int ld_off, const float* scale_logits, int ld_log,
                           const float* pc_range, int BQ, int GP, int L,
                           float* points, float* scale_w, void* stream);

/* Box refinement closing a decoder layer (models/sparsebev_transformer.py:155-160 refine_bbox, :179-183
 * velocity rescale): out[...,0:3] = sigmoid(delta[...,0:3] + inverse_sigmoid(proposal[...,0:3])),
 * out[...,3:] = delta[...,3:], and if T > 1 out[...,8:] /= (time_diff[b,1] < 1e-5 ? 1 : time_diff[b,1]).
 *   proposal, delta, out [B,Q,code_size]; time_diff [B,T]
 */
int sbev_refine_bbox_fwd(const float* proposal, const float* delta, const float* time_diff,
                         int B, int Q, int T, int code_size, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Scale-adaptive self-attention core: softmax(q k^T / sqrt(hd) - tau_h * dist(c_i, c_j)) v.
 * Replaces SparseBEVSelfAttention.inner_forward's mask build + nn.MultiheadAttention core
 * (models/sparsebev_transformer.py:210-248); the [B*8,Q,Q] mask is never materialised.
 *   qkv [B*Q rows, row stride ld_qkv >= 3*D] (in_proj output, q|k|v), query_bbox [B,Q,10] (centres are decoded
 *   in-kernel with pc_range, HOST float[6]), tau [B*Q rows, row stride ld_tau >= H] (may alias columns of the
 *   same matrix as qkv when in_proj and gen_tau are run as one concatenated Linear), dn_mask optional [Q,Q]
 *   uint8 (1 = blocked, query denoising), out [B,Q,D] (heads concatenated, before out_proj)
 */
int sbev_sasa_fwd(const float* qkv, int ld_qkv, const float* query_bbox, const float* tau, int ld_tau,
                  const uint8_t* dn_mask, const float* pc_range, int B, int Q, int H, int D, float* out, void* stream);

/* Same attention core, tensor-core fast path: the in_proj output pre-split ONCE into bf16 (hi, lo) (sbev_split_bf16
 * over the whole [B*Q, ld] matrix, ld % 8 == 0); tau stays fp32.  Every warp is an independent worker on
 * (16 queries, head, quarter of the keys) with its own cp.async double buffer; partials merged per CTA. */
int sbev_sasa_split_fwd(const uint16_t* qkv_hi, const uint16_t* qkv_lo, int ld, const float* query_bbox,
                        const float* tau, int ld_tau, const uint8_t* dn_mask, const float* pc_range,
                        int B, int Q, int H, int D, float* out, void* stream);

/* sbev_sasa_split_fwd for the QUERIES [q_begin, q_end) only (keys / values are always all Q rows): the query-sharded decoder
 * runs the attention of its own queries.  Only rows [q_begin, q_end) of out [B,Q,D] are written; they are bit-identical to
 * the same rows of the full call (a query's result does not depend on which other queries share its tile). */
int sbev_sasa_split_range_fwd(const uint16_t* qkv_hi, const uint16_t* qkv_lo, int ld, const float* query_bbox,
                              const float* tau, int ld_tau, const uint8_t* dn_mask, const float* pc_range,
                              int B, int Q, int H, int D, int q_begin, int q_end, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * AdaptiveMixing (models/sparsebev_transformer.py:351-381).
 * (1) dynamic-parameter generation and (3) output projection are GEMMs on the tcgen05 tensor cores
 *     (sbev_gemm_bf16_tn below); (2) is the per-(query,group) mix:
 *       h = relu(LN_{Pin x C}(x @ M));  y = relu(LN_{Pout x C}(S @ h))
 *   params [BQ, G*(C*C + Pout*Pin)] fp32 (bias already added), x [BQ,G,Pin,C],
 *   y_hi/y_lo [BQ, G*Pout*C] bf16 split of y (y ~= hi + lo) ready for the out_proj GEMM; y_f32 optional.
 */
int sbev_mix_fwd(const float* params, const float* x, int BQ, int G, int Pin, int Pout, int C,
                 uint16_t* y_hi, uint16_t* y_lo, float* y_f32, void* stream);

/* fp32 -> (hi, lo) bf16 split, elementwise: hi = bf16(x), lo = bf16(x - hi). lo may be NULL. */
int sbev_split_bf16(const float* x, int64_t n, uint16_t* hi, uint16_t* lo, void* stream);

/* C[M,N] (fp32) = sum_s A_s[M,K] . B_s[N,K]^T over `nseg` operand pairs (bf16, K-major), + bias.
 * tcgen05.mma kind::f16 with fp32 accumulators in TMEM, TMA-staged 128B-swizzled operand tiles.
 * nseg = 1: plain bf16 GEMM; nseg = 3 with (A_hi,B_hi),(A_hi,B_lo),(A_lo,B_hi): fp32-grade "bf16x3".
 *   split_k > 1: partial sums go to `C + z*M*N` for z in [0, split_k) (caller reduces); bias only in z=0; slice z covers
 *   k-blocks [z*(K/64)/split_k, (z+1)*(K/64)/split_k) -- split_k need not divide K/64.
 * Requires K % 64 == 0, N % 128 == 0; rows of A beyond M are treated as zero.
 */
int sbev_gemm_bf16_tn(const uint16_t* const* A, const uint16_t* const* B, int nseg,
                      const float* bias, int M, int N, int K, int split_k, float* C, void* stream);

/* Same GEMM in bf16x3 mode, but C leaves as a bf16 (hi, lo) pair (C ~= C_hi + C_lo, [M][N] each, N % 256 == 0): the form the
 * TMA-fed mix kernel consumes, so the dynamic-parameter tensor is written once and never re-converted. */
int sbev_gemm_bf16_tn_split(const uint16_t* A_hi, const uint16_t* A_lo, const uint16_t* B_hi, const uint16_t* B_lo,
                            const float* bias, int M, int N, int K, uint16_t* C_hi, uint16_t* C_lo, void* stream);

/* sbev_mix_fwd with the parameters given as that (hi, lo) pair; in_points must be 32 (M 64x64 and S 128x32 tiles are
 * pulled by swizzled 2-D TMA loads straight into ldmatrix-ready shared memory, double-buffered across items). */
int sbev_mix_presplit_fwd(const uint16_t* params_hi, const uint16_t* params_lo, const float* x, int BQ, int G, int Pin,
                          int Pout, int C, uint16_t* y_hi, uint16_t* y_lo, float* y_f32, void* stream);

/* out[M,N] = LN(sum_z partial[z] + bias + residual) : split-K reduction fused with the residual
 * and LayerNorm that follow mixing.out_proj (sparsebev_transformer.py:377-379 + norm2 at :171). */
int sbev_reduce_ln_fwd(const float* partial, int nsplit, const float* bias, const float* residual,
                       const float* ln_w, const float* ln_b, int M, int N, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backbone convolutions.
 * Replaces the cuDNN calls (F.conv2d -> norm -> activation) behind the reference's conv wrappers:
 * models/backbones/eva02/wrappers.py:76-120 (Conv2d), models/backbones/vovnet.py:117-154 (conv3x3 / conv1x1 =
 * Conv2d + BatchNorm2d + ReLU) and the mmdet ResNet / FPN blocks built by configs/r50_nuimg_704x256.py:31-45, as
 * called from models/sparsebev.py:46-59 (extract_img_feat, fp16 autocast there; bf16 operands + fp32 accumulate here).
 *
 * out[n,ho,wo,co] = act( scale[co] * sum_{kh,kw,ci} x[n, ho*stride+kh-pad, wo*stride+kw-pad, ci] * w[co,kh,kw,ci] + shift[co]
 *                        + residual[n, ho*res_H/Ho, wo*res_W/Wo, co] )
 *   x        NHWC bf16 [Nimg][H][W][Cin], Cin % 64 == 0          w  bf16 [Cout][KH][KW][Cin], Cout % 32 == 0
 *   scale    fp32 [Cout] or NULL (= 1): folded BatchNorm gamma / sqrt(var + eps);  shift fp32 [Cout]: folded beta / conv bias
 *   residual NHWC bf16 [Nimg][res_H][res_W][Cout] or NULL; res_H x res_W == Ho x Wo is the bottleneck identity, a smaller
 *            map is read nearest-neighbour (FPN top-down: lateral + upsampled coarser level)
 *   out      NHWC [Nimg][Ho][Wo][Cout], bf16 (out_f32 == 0) or fp32 (out_f32 != 0); stride 1 or 2, zero padding.
 * Implicit GEMM on tcgen05 (4-D TMA boxes of 8 x 16 output pixels x 64 channels per tap, weights by 2-D TMA, fp32
 * accumulators in TMEM).  All operands 16-byte aligned; caller allocates everything. */
int sbev_conv2d_nhwc_fwd(const uint16_t* x, int Nimg, int H, int W, int Cin,
                         const uint16_t* w, int Cout, int KH, int KW, int stride, int pad,
                         const float* scale, const float* shift,
                         const uint16_t* residual, int res_H, int res_W, int relu,
                         void* out, int out_f32, void* stream);

/* ResNet stem: 7x7 stride-2 pad-3 conv of the NCHW fp32 image [Nimg][3][H][W] (what models/sparsebev.py:61-100 hands the
 * backbone) + folded BN + ReLU -> NHWC bf16 [Nimg][Ho][Wo][64].  w fp32 [7][7][3][64]. */
int sbev_stem_conv_fwd(const float* img, int Nimg, int H, int W, const float* w, const float* scale, const float* shift,
                       uint16_t* out, void* stream);

/* Same with a ksize x ksize kernel, ksize = 7 (pad 3) or 3 (pad 1: VoVNet's stem_1, models/backbones/vovnet.py:289); w fp32 [ksize][ksize][3][64]. */
int sbev_stem_conv_k_fwd(const float* img, int Nimg, int H, int W, const float* w, int ksize, const float* scale, const float* shift,
                         uint16_t* out, void* stream);

/* 3x3 stride-2 pad-1 max pool, NHWC bf16, C % 8 == 0 -> [Nimg][(H-1)/2+1][(W-1)/2+1][C]. */
int sbev_maxpool3x3s2_nhwc_fwd(const uint16_t* x, int Nimg, int H, int W, int C, uint16_t* out, void* stream);

/* 3x3 stride-2 max pool with explicit padding (0 | 1) and torch's ceil_mode: VoVNet's OSA stages open with
 * nn.MaxPool2d(kernel_size=3, stride=2, ceil_mode=True) (models/backbones/vovnet.py:228).  out [Nimg][Ho][Wo][C], Ho / Wo by
 * torch.nn.MaxPool2d's rule (window positions beyond the input are ignored). */
int sbev_maxpool3x3s2_ex_nhwc_fwd(const uint16_t* x, int Nimg, int H, int W, int C, int pad, int ceil_mode, uint16_t* out, void* stream);

/* Effective squeeze-excitation of a VoVNet OSA block (models/backbones/vovnet.py:157-178,213-218), NHWC bf16:
 *   out = x * hsigmoid(fc(mean_{h,w} x)) (+ identity),   hsigmoid(v) = relu6(v + 3) / 6,   fc = 1x1 conv [C,C] + bias (fp32).
 * Three launches (deterministic two-stage mean, gate mat-vec, scale); workspace = sbev_ese_workspace_floats(...) fp32, caller-owned. */
long long sbev_ese_workspace_floats(int Nimg, int H, int W, int C);
int sbev_ese_nhwc_fwd(const uint16_t* x, int Nimg, int H, int W, int C, const float* fc_weight, const float* fc_bias,
                      const uint16_t* identity, float* workspace, uint16_t* out, void* stream);

/* out[n,ho,wo,:] = x[n,2ho,2wo,:], fp32 NHWC, C % 4 == 0 (mmdet FPN extra level: F.max_pool2d(x, 1, stride=2)). */
int sbev_subsample2_nhwc_fwd(const float* x, int Nimg, int H, int W, int C, float* out, void* stream);

/* fp32 -> bf16 (round to nearest even), elementwise. */
int sbev_cast_bf16(const float* x, int64_t n, uint16_t* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* SPARSEBEV_B200_H_ */
