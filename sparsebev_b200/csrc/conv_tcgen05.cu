// Backbone convolutions for sm_100a: conv (+ folded BatchNorm / bias) (+ residual) (+ ReLU) as an implicit GEMM on
// tcgen05, activations NHWC bf16, fp32 accumulation in TMEM, output NHWC bf16 (next conv's operand) or fp32 (the FPN
// levels, which the fused gather consumes in place: [B, T*N, H, W, 256] is its 'nhwc' layout).
//
// Replaces the cuDNN calls behind the reference's conv wrappers:
//   /root/reference/models/backbones/eva02/wrappers.py:76-120  (Conv2d = F.conv2d + norm + activation),
//   /root/reference/models/backbones/vovnet.py:117-154         (conv3x3 / conv1x1 = Conv2d + BatchNorm2d + ReLU),
//   and the mmdet ResNet / FPN blocks of configs/r50_nuimg_704x256.py:31-45 (third party, same conv + BN + ReLU pattern),
//   called from models/sparsebev.py:46-59 (extract_img_feat).
//
// Implicit GEMM without im2col: the M dimension of a tile is an 8 x 16 patch of output pixels of ONE image, the K loop
// walks (kh, kw, 64-channel block).  For each step the producer issues ONE 4-D TMA load of the box
// {64 channels, 16 w, 8 h, 1 image} from the input at the tap's offset -- it lands in shared memory as 128 rows of 128
// bytes with the 128-byte swizzle, i.e. exactly the K-major operand tile tcgen05.mma wants -- and one 2-D load of the
// weight tile [BN][64] (weights are stored [Cout][KH][KW][Cin]).  Zero padding is TMA's out-of-bounds fill (coordinates
// may be negative).  Strided convolutions use one tensor map per input parity class: input pixel s*o + d (d = k - pad)
// is element o + floor(d / s) of the parity-(d mod s) sub-grid, which is a unit-stride box again.
// Pipeline as in gemm_tcgen05.cu: warp 0 TMA producer, warp 1 single-thread MMA issue into one of two TMEM accumulators,
// warps 2-5 epilogue (tcgen05.ld -> scale / shift -> + residual (optionally nearest-upsampled: the FPN top-down path)
// -> ReLU -> 16-byte stores), persistent CTAs.  All mbarrier waits are bounded (trap, never hang).
//
// Also here: the 7x7/2 stem (3 input channels: not tensor-core shaped; direct fp32 convolution + BN + ReLU from the NCHW
// image the reference feeds the backbone) and the 3x3/2 max pool, both writing NHWC bf16.
#include "tcgen05.cuh"
#include <mutex>

namespace sbev {

constexpr int CONV_TW = 16, CONV_TH = 8;          // output patch of a tile: 8 rows x 16 columns = 128 pixels = 128 TMEM lanes
constexpr int CONV_THREADS = 192;
constexpr int CONV_A_BYTES = 128 * 64 * 2;

struct ConvParams {
    int Nimg, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad;
    int tiles_w, tiles_h, m_tiles, n_tiles, cin_blocks;
    const float* scale; const float* shift;
    const __nv_bfloat16* residual; int res_H, res_W;
    int relu, out_f32;
    void* out;
};
struct ConvMaps { CUtensorMap x[4]; CUtensorMap w; };

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <int BN, int STAGES>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ ConvParams prm, const __grid_constant__ ConvMaps maps) {
    constexpr int B_BYTES = BN * 64 * 2;
    constexpr int STAGE_BYTES = CONV_A_BYTES + B_BYTES;
    extern __shared__ uint8_t conv_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(conv_smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = prm.m_tiles * prm.n_tiles;
    const int k_steps = prm.KH * prm.KW * prm.cin_blocks;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x[0]) : "memory");
            int it = 0;
            const int s = prm.stride;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mt = tile % prm.m_tiles, n0 = (tile / prm.m_tiles) * BN;
                const int tw = mt % prm.tiles_w, th = (mt / prm.tiles_w) % prm.tiles_h, img = mt / (prm.tiles_w * prm.tiles_h);
                for (int kh = 0; kh < prm.KH; ++kh) {
                    const int dh = kh - prm.pad, qh = floor_div(dh, s), ph = dh - qh * s;
                    for (int kw = 0; kw < prm.KW; ++kw) {
                        const int dw = kw - prm.pad, qw = floor_div(dw, s), pw = dw - qw * s;
                        const CUtensorMap* xm = &maps.x[ph * s + pw];
                        for (int cb = 0; cb < prm.cin_blocks; ++cb, ++it) {
                            const int stage = it % STAGES;
                            mbar_wait(&empty_bar[stage], ((it / STAGES) & 1) ^ 1);
                            uint8_t* st = smem + stage * STAGE_BYTES;
                            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                            tma_load_4d(st, xm, &full_bar[stage], cb * 64, tw * CONV_TW + qw, th * CONV_TH + qh, img);
                            tma_load_2d(st + CONV_A_BYTES, &maps.w, &full_bar[stage], ((kh * prm.KW + kw) * prm.cin_blocks + cb) * 64, n0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_f32(128, BN);
            int it = 0, lt = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
                const int as = lt & 1;
                mbar_wait(&tmem_empty_bar[as], ((lt >> 1) & 1) ^ 1);        // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(as * BN);
                for (int ks = 0; ks < k_steps; ++ks, ++it) {
                    const int stage = it % STAGES;
                    mbar_wait(&full_bar[stage], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t a = umma_desc_k_sw128(st), b = umma_desc_k_sw128(st + CONV_A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tacc, a + 2 * k, b + 2 * k, idesc, (ks > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                }
                umma_commit(&tmem_full_bar[as]);
            }
        }
    } else {
        const int quarter = warp & 3;
        int lt = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
            const int mt = tile % prm.m_tiles, n0 = (tile / prm.m_tiles) * BN;
            const int tw = mt % prm.tiles_w, th = (mt / prm.tiles_w) % prm.tiles_h, img = mt / (prm.tiles_w * prm.tiles_h);
            const int pix = quarter * 32 + lane;                           // TMEM lane = pixel of the patch (row-major 8 x 16)
            const int ho = th * CONV_TH + pix / CONV_TW, wo = tw * CONV_TW + pix % CONV_TW;
            const bool live = ho < prm.Ho && wo < prm.Wo;
            const long long opix = ((long long)img * prm.Ho + ho) * prm.Wo + wo;
            long long rpix = 0;
            if (prm.residual != nullptr) {
                const int rh = prm.res_H == prm.Ho ? ho : (int)((long long)ho * prm.res_H / prm.Ho);
                const int rw = prm.res_W == prm.Wo ? wo : (int)((long long)wo * prm.res_W / prm.Wo);
                rpix = ((long long)img * prm.res_H + rh) * prm.res_W + rw;
            }
            const int as = lt & 1;
            mbar_wait(&tmem_full_bar[as], (lt >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + c * 32), v);
                const int n = n0 + c * 32;
                if (live && n < prm.Cout) {                                // (Cout % 32 == 0: a chunk is all in or all out)
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 sh = ldg4(prm.shift + n + i);
                    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (prm.scale != nullptr) sc = ldg4(prm.scale + n + i);
                    v[i] = v[i] * sc.x + sh.x; v[i + 1] = v[i + 1] * sc.y + sh.y; v[i + 2] = v[i + 2] * sc.z + sh.z; v[i + 3] = v[i + 3] * sc.w + sh.w;
                }
                if (prm.residual != nullptr) {
                    const uint4* rp = reinterpret_cast<const uint4*>(prm.residual + rpix * prm.Cout + n);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 r = __ldg(rp + j);
                        const uint32_t rw4[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rw4[e]));
                            v[8 * j + 2 * e] += f.x; v[8 * j + 2 * e + 1] += f.y;
                        }
                    }
                }
                if (prm.relu) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                if (prm.out_f32) {
                    float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(prm.out) + opix * prm.Cout + n);
#pragma unroll
                    for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                } else {
                    uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(prm.out) + opix * prm.Cout + n);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t w4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
                            w4[e] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                        op[j] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                    }
                }
                }
                __syncwarp();                                              // reconverge before the next (warp-aligned) tcgen05.ld
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[as])) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * BN); }
}

// ---------------------------------------------------------------------------------------------------------------------
// Stem: 7x7 stride-2 pad-3 convolution of the 3-channel NCHW fp32 image + folded BN + ReLU -> NHWC bf16 [N][Ho][Wo][64].
// One CTA = 8 x 16 output pixels x 64 channels; the input patch (21 x 37 x 3) and all weights ([147][64] fp32, 37 KB)
// sit in shared memory; a thread owns one pixel and 32 channels (weights are warp-uniform LDS.128 broadcasts).
// KS = 7 (ResNet stem, pad 3) or 3 (VoVNet stem_1, models/backbones/vovnet.py:289, pad 1); stride 2.
constexpr int STEM_TH = 8, STEM_TW = 16;
template <int KS>
__global__ void __launch_bounds__(256)
stem_conv_kernel(const float* __restrict__ img, const float* __restrict__ w /* [KS][KS][3][64] */, const float* __restrict__ scale,
                 const float* __restrict__ shift, int Nimg, int H, int W, int Ho, int Wo, __nv_bfloat16* __restrict__ out) {
    constexpr int STEM_PH = STEM_TH * 2 + KS - 2, STEM_PW = STEM_TW * 2 + KS - 2, TAPS = KS * KS * 3;
    extern __shared__ float stem_smem[];
    float* ws = stem_smem;                       // [TAPS][64]
    float* xs = stem_smem + TAPS * 64;           // [3][STEM_PH][STEM_PW]
    const int tiles_w = (Wo + STEM_TW - 1) / STEM_TW, tiles_h = (Ho + STEM_TH - 1) / STEM_TH;
    const int tw = blockIdx.x % tiles_w, th = (blockIdx.x / tiles_w) % tiles_h, n = blockIdx.x / (tiles_w * tiles_h);
    pdl_wait();
    pdl_trigger();
    for (int i = threadIdx.x; i < TAPS * 64 / 4; i += 256) reinterpret_cast<float4*>(ws)[i] = ldg4(w + 4 * i);
    const int h_in0 = th * STEM_TH * 2 - KS / 2, w_in0 = tw * STEM_TW * 2 - KS / 2;
    for (int i = threadIdx.x; i < 3 * STEM_PH * STEM_PW; i += 256) {
        const int c = i / (STEM_PH * STEM_PW), r = (i / STEM_PW) % STEM_PH, q = i % STEM_PW;
        const int hh = h_in0 + r, ww = w_in0 + q;
        xs[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(img + (((long long)n * 3 + c) * H + hh) * W + ww) : 0.f;
    }
    __syncthreads();
    const int pix = threadIdx.x & 127, half = threadIdx.x >> 7;      // warps 0-3: channels 0..31, warps 4-7: channels 32..63
    const int pr = pix / STEM_TW, pc = pix % STEM_TW;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    for (int kh = 0; kh < KS; ++kh)
        for (int kw = 0; kw < KS; ++kw)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float x = xs[(c * STEM_PH + pr * 2 + kh) * STEM_PW + pc * 2 + kw];
                const float4* wp = reinterpret_cast<const float4*>(ws + ((kh * KS + kw) * 3 + c) * 64 + half * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 wv = wp[j];
                    acc[4 * j] = fmaf(x, wv.x, acc[4 * j]); acc[4 * j + 1] = fmaf(x, wv.y, acc[4 * j + 1]);
                    acc[4 * j + 2] = fmaf(x, wv.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(x, wv.w, acc[4 * j + 3]);
                }
            }
    const int ho = th * STEM_TH + pr, wo = tw * STEM_TW + pc;
    if (ho >= Ho || wo >= Wo) return;
    uint4* op = reinterpret_cast<uint4*>(out + (((long long)n * Ho + ho) * Wo + wo) * 64 + half * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t w4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int ch = half * 32 + 8 * j + 2 * e;
            const float a = fmaxf(acc[8 * j + 2 * e] * __ldg(scale + ch) + __ldg(shift + ch), 0.f);
            const float b = fmaxf(acc[8 * j + 2 * e + 1] * __ldg(scale + ch + 1) + __ldg(shift + ch + 1), 0.f);
            const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
            w4[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        op[j] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
}

// 3x3 stride-2 pad-1 max pool, NHWC bf16 (C % 8 == 0): a thread owns 8 channels of one output pixel.
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x, int Nimg, int H, int W, int C, int Ho, int Wo, __nv_bfloat16* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int c8 = C / 8;
    const long long total = (long long)Nimg * Ho * Wo * c8;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int cg = (int)(i % c8);
        const long long p = i / c8;
        const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
        float m[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = -3.0e38f;
        for (int dh = -1; dh <= 1; ++dh)
            for (int dw = -1; dw <= 1; ++dw) {
                const int hh = 2 * ho + dh, ww = 2 * wo + dw;
                if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
                const uint4 r = __ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + hh) * W + ww) * C + cg * 8));
                const uint32_t r4[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r4[e]));
                    m[2 * e] = fmaxf(m[2 * e], f.x); m[2 * e + 1] = fmaxf(m[2 * e + 1], f.y);
                }
            }
        uint32_t w4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(m[2 * e], m[2 * e + 1]);
            w4[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(out + p * C + cg * 8) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
}

// out[n][ho][wo][:] = x[n][2 ho][2 wo][:] (fp32 NHWC, C % 4 == 0): mmdet FPN's extra level, F.max_pool2d(x, 1, stride=2)
__global__ void __launch_bounds__(256)
subsample2_kernel(const float* __restrict__ x, int Nimg, int H, int W, int C, int Ho, int Wo, float* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int c4 = C / 4;
    const long long total = (long long)Nimg * Ho * Wo * c4;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int cg = (int)(i % c4);
        const long long p = i / c4;
        const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
        reinterpret_cast<float4*>(out)[i] = ldg4(x + (((long long)n * H + 2 * ho) * W + 2 * wo) * C + cg * 4);
    }
}

// fp32 -> bf16 cast (weights / test inputs), 16-byte aligned, n % 4 == 0 handled by the tail loop
__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ x, long long n, __nv_bfloat16* __restrict__ y) {
    pdl_wait();
    pdl_trigger();
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) y[i] = __float2bfloat16_rn(__ldg(x + i));
}

// ------------------------------------------------------------------------------- host side
typedef CUresult (*ConvEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static ConvEncodeFn conv_encode_fn() {
    static ConvEncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<ConvEncodeFn>(p);
    });
    return fn;
}

// Parity-(ph, pw) sub-grid of an NHWC bf16 tensor as a 4-D map {C, ceil((W-pw)/s), ceil((H-ph)/s), N}, box {64, 16, 8, 1}.
static int make_act_map(CUtensorMap* out, const void* x, int N, int H, int W, int C, int s, int ph, int pw) {
    ConvEncodeFn enc = conv_encode_fn();
    SBEV_REQUIRE(enc != nullptr, SBEV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const char* base = reinterpret_cast<const char*>(x) + ((long long)ph * W + pw) * C * 2;
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)((W - pw + s - 1) / s), (cuuint64_t)((H - ph + s - 1) / s), (cuuint64_t)N};
    cuuint64_t gstride[3] = {(cuuint64_t)s * C * 2, (cuuint64_t)s * W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)CONV_TW, (cuuint32_t)CONV_TH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SBEV_REQUIRE(r == CUDA_SUCCESS, SBEV_ERR_CUDA, "cuTensorMapEncodeTiled(conv input %dx%dx%dx%d, stride %d) failed (%d)", N, H, W, C, s, (int)r);
    return SBEV_OK;
}

template <int BN, int STAGES>
static int launch_conv(const ConvParams& prm, const ConvMaps& maps, cudaStream_t st) {
    constexpr size_t smem = (size_t)STAGES * (CONV_A_BYTES + BN * 128) + 1024;
    SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(conv_igemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = device_num_sms();
    const int tiles = prm.m_tiles * prm.n_tiles;
    launch_pdl(conv_igemm_kernel<BN, STAGES>, dim3(tiles < sms ? tiles : sms), dim3(CONV_THREADS), smem, st, prm, maps);
    return check_launch("sbev_conv2d_nhwc_fwd");
}

}  // namespace sbev

using namespace sbev;

extern "C" int sbev_conv2d_nhwc_fwd(const uint16_t* x, int Nimg, int H, int W, int Cin,
                                    const uint16_t* w, int Cout, int KH, int KW, int stride, int pad,
                                    const float* scale, const float* shift,
                                    const uint16_t* residual, int res_H, int res_W, int relu,
                                    void* out, int out_f32, void* stream) {
    SBEV_REQUIRE(x && w && shift && out, SBEV_ERR_INVALID, "sbev_conv2d_nhwc_fwd: null pointer");
    SBEV_REQUIRE(Nimg >= 0 && H > 0 && W > 0, SBEV_ERR_INVALID, "sbev_conv2d_nhwc_fwd: bad sizes");
    SBEV_REQUIRE(Cin % 64 == 0 && Cin > 0, SBEV_ERR_UNSUPPORTED, "sbev_conv2d_nhwc_fwd: Cin must be a multiple of 64 (got %d; the 3-channel stem has its own entry point)", Cin);
    SBEV_REQUIRE(Cout % 32 == 0 && Cout > 0, SBEV_ERR_UNSUPPORTED, "sbev_conv2d_nhwc_fwd: Cout must be a multiple of 32 (got %d)", Cout);
    SBEV_REQUIRE(KH >= 1 && KH <= 7 && KW >= 1 && KW <= 7 && (stride == 1 || stride == 2) && pad >= 0 && pad < KH + 8, SBEV_ERR_UNSUPPORTED,
                 "sbev_conv2d_nhwc_fwd: kernel %dx%d stride %d pad %d not supported", KH, KW, stride, pad);
    const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
    SBEV_REQUIRE(Ho > 0 && Wo > 0, SBEV_ERR_INVALID, "sbev_conv2d_nhwc_fwd: empty output");
    const void* al[6] = {x, w, out, residual, scale, shift};
    for (int i = 0; i < 6; ++i) SBEV_REQUIRE((reinterpret_cast<uintptr_t>(al[i]) & 15) == 0, SBEV_ERR_INVALID, "sbev_conv2d_nhwc_fwd: operands must be 16-byte aligned");
    if (residual) SBEV_REQUIRE(res_H > 0 && res_W > 0 && res_H <= Ho && res_W <= Wo, SBEV_ERR_INVALID, "sbev_conv2d_nhwc_fwd: bad residual size");
    if (Nimg == 0) return SBEV_OK;
    ConvParams prm;
    prm.Nimg = Nimg; prm.H = H; prm.W = W; prm.Cin = Cin; prm.Ho = Ho; prm.Wo = Wo; prm.Cout = Cout;
    prm.KH = KH; prm.KW = KW; prm.stride = stride; prm.pad = pad;
    prm.tiles_w = (Wo + CONV_TW - 1) / CONV_TW; prm.tiles_h = (Ho + CONV_TH - 1) / CONV_TH;
    prm.m_tiles = Nimg * prm.tiles_w * prm.tiles_h;
    prm.cin_blocks = Cin / 64;
    prm.scale = scale; prm.shift = shift;
    prm.residual = reinterpret_cast<const __nv_bfloat16*>(residual); prm.res_H = res_H; prm.res_W = res_W;
    prm.relu = relu; prm.out_f32 = out_f32; prm.out = out;
    ConvMaps maps;
    for (int ph = 0; ph < stride; ++ph)
        for (int pw = 0; pw < stride; ++pw) {
            if (ph >= H || pw >= W) { maps.x[ph * stride + pw] = maps.x[0]; continue; }
            int rc = make_act_map(&maps.x[ph * stride + pw], x, Nimg, H, W, Cin, stride, ph, pw);
            if (rc) return rc;
        }
    int rc;
    const int bn = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
    prm.n_tiles = (Cout + bn - 1) / bn;
    rc = make_bf16_map(&maps.w, w, Cout, (long long)KH * KW * Cin, bn);
    if (rc) return rc;
    if (bn == 256) return launch_conv<256, 4>(prm, maps, (cudaStream_t)stream);
    if (bn == 128) return launch_conv<128, 6>(prm, maps, (cudaStream_t)stream);
    return launch_conv<64, 8>(prm, maps, (cudaStream_t)stream);
}

extern "C" int sbev_stem_conv_k_fwd(const float* img, int Nimg, int H, int W, const float* w, int ksize, const float* scale, const float* shift,
                                    uint16_t* out, void* stream);

extern "C" int sbev_stem_conv_fwd(const float* img, int Nimg, int H, int W, const float* w, const float* scale, const float* shift,
                                  uint16_t* out, void* stream) {
    SBEV_REQUIRE(img && w && scale && shift && out, SBEV_ERR_INVALID, "sbev_stem_conv_fwd: null pointer");
    SBEV_REQUIRE(Nimg >= 0 && H > 0 && W > 0, SBEV_ERR_INVALID, "sbev_stem_conv_fwd: bad sizes");
    SBEV_REQUIRE(((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, SBEV_ERR_INVALID, "sbev_stem_conv_fwd: w / out must be 16-byte aligned");
    if (Nimg == 0) return SBEV_OK;
    return sbev_stem_conv_k_fwd(img, Nimg, H, W, w, 7, scale, shift, out, stream);
}

extern "C" int sbev_stem_conv_k_fwd(const float* img, int Nimg, int H, int W, const float* w, int ksize, const float* scale, const float* shift,
                                    uint16_t* out, void* stream) {
    SBEV_REQUIRE(img && w && scale && shift && out, SBEV_ERR_INVALID, "sbev_stem_conv_k_fwd: null pointer");
    SBEV_REQUIRE(ksize == 7 || ksize == 3, SBEV_ERR_UNSUPPORTED, "sbev_stem_conv_k_fwd: kernel size 3 or 7 (got %d)", ksize);
    SBEV_REQUIRE(Nimg >= 0 && H > 0 && W > 0, SBEV_ERR_INVALID, "sbev_stem_conv_k_fwd: bad sizes");
    SBEV_REQUIRE(((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, SBEV_ERR_INVALID, "sbev_stem_conv_k_fwd: operands must be 16-byte aligned");
    if (Nimg == 0) return SBEV_OK;
    const int pad = ksize / 2;
    const int Ho = (H + 2 * pad - ksize) / 2 + 1, Wo = (W + 2 * pad - ksize) / 2 + 1;
    const size_t smem = (size_t)(ksize * ksize * 3 * 64 + 3 * (STEM_TH * 2 + ksize - 2) * (STEM_TW * 2 + ksize - 2)) * sizeof(float);
    const int tiles = Nimg * ((Wo + STEM_TW - 1) / STEM_TW) * ((Ho + STEM_TH - 1) / STEM_TH);
    if (ksize == 7) {
        SBEV_PER_DEVICE_ONCE(cudaFuncSetAttribute(stem_conv_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        launch_pdl(stem_conv_kernel<7>, dim3(tiles), dim3(256), smem, (cudaStream_t)stream, img, w, scale, shift, Nimg, H, W, Ho, Wo,
                   reinterpret_cast<__nv_bfloat16*>(out));
    } else {
        launch_pdl(stem_conv_kernel<3>, dim3(tiles), dim3(256), smem, (cudaStream_t)stream, img, w, scale, shift, Nimg, H, W, Ho, Wo,
                   reinterpret_cast<__nv_bfloat16*>(out));
    }
    return check_launch("sbev_stem_conv_fwd");
}

extern "C" int sbev_maxpool3x3s2_nhwc_fwd(const uint16_t* x, int Nimg, int H, int W, int C, uint16_t* out, void* stream) {
    SBEV_REQUIRE(x && out, SBEV_ERR_INVALID, "sbev_maxpool3x3s2_nhwc_fwd: null pointer");
    SBEV_REQUIRE(C % 8 == 0 && C > 0 && H > 0 && W > 0 && Nimg >= 0, SBEV_ERR_UNSUPPORTED, "sbev_maxpool3x3s2_nhwc_fwd: C must be a multiple of 8");
    SBEV_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, SBEV_ERR_INVALID, "sbev_maxpool3x3s2_nhwc_fwd: operands must be 16-byte aligned");
    if (Nimg == 0) return SBEV_OK;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = (long long)Nimg * Ho * Wo * (C / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_pdl(maxpool3x3s2_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(x), Nimg, H, W, C, Ho, Wo,
               reinterpret_cast<__nv_bfloat16*>(out));
    return check_launch("sbev_maxpool3x3s2_nhwc_fwd");
}

extern "C" int sbev_cast_bf16(const float* x, int64_t n, uint16_t* y, void* stream) {
    SBEV_REQUIRE(x && y && n >= 0, SBEV_ERR_INVALID, "sbev_cast_bf16: bad arguments");
    if (n == 0) return SBEV_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_pdl(cast_bf16_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, (long long)n, reinterpret_cast<__nv_bfloat16*>(y));
    return check_launch("sbev_cast_bf16");
}

extern "C" int sbev_subsample2_nhwc_fwd(const float* x, int Nimg, int H, int W, int C, float* out, void* stream) {
    SBEV_REQUIRE(x && out, SBEV_ERR_INVALID, "sbev_subsample2_nhwc_fwd: null pointer");
    SBEV_REQUIRE(C % 4 == 0 && C > 0 && H > 0 && W > 0 && Nimg >= 0, SBEV_ERR_UNSUPPORTED, "sbev_subsample2_nhwc_fwd: C must be a multiple of 4");
    SBEV_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, SBEV_ERR_INVALID, "sbev_subsample2_nhwc_fwd: operands must be 16-byte aligned");
    if (Nimg == 0) return SBEV_OK;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long total = (long long)Nimg * Ho * Wo * (C / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_pdl(subsample2_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, Nimg, H, W, C, Ho, Wo, out);
    return check_launch("sbev_subsample2_nhwc_fwd");
}
