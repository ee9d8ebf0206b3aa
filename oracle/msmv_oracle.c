/* CPU oracle for the msmv_sampling op with the reference *CUDA kernel's* semantics.
 * TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline) -- never linked by the product.
 *
 * Restates, as scalar fp32 C:
 *   forward   /root/reference/models/csrc/msmv_sampling/msmv_sampling_forward.cu:27-73 (bilinear tap),
 *             :75-164 (c2345) and :166-267 (c23456), generalised to L levels;
 *   backward  /root/reference/models/csrc/msmv_sampling/msmv_sampling_backward.cu:29-105, :108-224
 *             (atomicAdd replaced by plain sequential accumulation; grad wrt the view coord is 0);
 *   indices   the integers the kernel derives from `loc`: view = round(z*(N-1)),
 *             y0 = floor(v*(H-1)), x0 = floor(u*(W-1)), and the whole-tap guard (:110,:123-126).
 * Layouts: feats[l] = [B',N,H_l,W_l,C] channel-last fp32; loc [B',Q,P,3]; w [B',Q,P,L];
 *          out / grad_out [B',Q,C,P].
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/c_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static float tap(const float* base, int H, int W, int C, float y, float x, int c)
{
    int y0 = (int)floorf(y), x0 = (int)floorf(x);
    int y1 = y0 + 1, x1 = x0 + 1;
    float ly = y - y0, lx = x - x0, hy = 1 - ly, hx = 1 - lx;
    float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
    if (y0 >= 0 && x0 >= 0)          v1 = base[((int64_t)y0 * W + x0) * C + c];
    if (y0 >= 0 && x1 <= W - 1)      v2 = base[((int64_t)y0 * W + x1) * C + c];
    if (y1 <= H - 1 && x0 >= 0)      v3 = base[((int64_t)y1 * W + x0) * C + c];
    if (y1 <= H - 1 && x1 <= W - 1)  v4 = base[((int64_t)y1 * W + x1) * C + c];
    return (hy * hx) * v1 + (hy * lx) * v2 + (ly * hx) * v3 + (ly * lx) * v4;
}

int msmv_oracle_fwd(const float* const* feats, const int* hw, int L, const float* loc, const float* w,
                    int Bp, int N, int C, int Q, int P, float* out)
{
    for (int b = 0; b < Bp; ++b)
      for (int q = 0; q < Q; ++q)
        for (int c = 0; c < C; ++c)
          for (int p = 0; p < P; ++p) {
            const float* lp = loc + (((int64_t)b * Q + q) * P + p) * 3;
            const float* wp = w + (((int64_t)b * Q + q) * P + p) * L;
            int view = (int)roundf(lp[2] * (N - 1));
            float acc = 0;
            for (int l = 0; l < L; ++l) {
                int H = hw[2 * l], W = hw[2 * l + 1];
                float y = lp[1] * (H - 1), x = lp[0] * (W - 1);
                if (y > -1 && x > -1 && y < H && x < W) {
                    const float* base = feats[l] + ((int64_t)b * N + view) * H * W * C;
                    acc += tap(base, H, W, C, y, x, c) * wp[l];
                }
            }
            out[(((int64_t)b * Q + q) * C + c) * P + p] = acc;
          }
    return 0;
}

/* view [B',Q,P]; y0,x0,inside [B',Q,P,L] */
int msmv_oracle_indices(const int* hw, int L, const float* loc, int Bp, int N, int Q, int P,
                        int32_t* view, int32_t* y0, int32_t* x0, int32_t* inside)
{
    for (int64_t i = 0; i < (int64_t)Bp * Q * P; ++i) {
        const float* lp = loc + i * 3;
        view[i] = (int)roundf(lp[2] * (N - 1));
        for (int l = 0; l < L; ++l) {
            int H = hw[2 * l], W = hw[2 * l + 1];
            float y = lp[1] * (H - 1), x = lp[0] * (W - 1);
            y0[i * L + l] = (int)floorf(y);
            x0[i * L + l] = (int)floorf(x);
            inside[i * L + l] = (y > -1 && x > -1 && y < H && x < W);
        }
    }
    return 0;
}

int msmv_oracle_bwd(const float* grad_out, const float* const* feats, const int* hw, int L,
                    const float* loc, const float* w, int Bp, int N, int C, int Q, int P,
                    float* const* grad_feats, float* grad_loc, float* grad_w)
{
    for (int l = 0; l < L; ++l)
        memset(grad_feats[l], 0, sizeof(float) * (size_t)Bp * N * hw[2 * l] * hw[2 * l + 1] * C);
    memset(grad_loc, 0, sizeof(float) * (size_t)Bp * Q * P * 3);
    memset(grad_w, 0, sizeof(float) * (size_t)Bp * Q * P * L);
    for (int b = 0; b < Bp; ++b)
      for (int q = 0; q < Q; ++q)
        for (int c = 0; c < C; ++c)
          for (int p = 0; p < P; ++p) {
            int64_t pi = ((int64_t)b * Q + q) * P + p;
            const float* lp = loc + pi * 3;
            float g = grad_out[(((int64_t)b * Q + q) * C + c) * P + p];
            int view = (int)roundf(lp[2] * (N - 1));
            for (int l = 0; l < L; ++l) {
                int H = hw[2 * l], W = hw[2 * l + 1];
                float y = lp[1] * (H - 1), x = lp[0] * (W - 1);
                if (!(y > -1 && x > -1 && y < H && x < W)) continue;
                int64_t off = ((int64_t)b * N + view) * H * W * C;
                const float* base = feats[l] + off;
                float* gbase = grad_feats[l] + off;
                float aw = w[pi * L + l], gv = g * aw;
                int y0 = (int)floorf(y), x0 = (int)floorf(x), y1 = y0 + 1, x1 = x0 + 1;
                float ly = y - y0, lx = x - x0, hy = 1 - ly, hx = 1 - lx;
                float v1 = 0, v2 = 0, v3 = 0, v4 = 0, gy = 0, gx = 0;
                if (y0 >= 0 && x0 >= 0) { int64_t o = ((int64_t)y0 * W + x0) * C + c; v1 = base[o];
                    gy -= hx * v1; gx -= hy * v1; gbase[o] += (hy * hx) * gv; }
                if (y0 >= 0 && x1 <= W - 1) { int64_t o = ((int64_t)y0 * W + x1) * C + c; v2 = base[o];
                    gy -= lx * v2; gx += hy * v2; gbase[o] += (hy * lx) * gv; }
                if (y1 <= H - 1 && x0 >= 0) { int64_t o = ((int64_t)y1 * W + x0) * C + c; v3 = base[o];
                    gy += hx * v3; gx -= ly * v3; gbase[o] += (ly * hx) * gv; }
                if (y1 <= H - 1 && x1 <= W - 1) { int64_t o = ((int64_t)y1 * W + x1) * C + c; v4 = base[o];
                    gy += lx * v4; gx += ly * v4; gbase[o] += (ly * lx) * gv; }
                float val = (hy * hx) * v1 + (hy * lx) * v2 + (ly * hx) * v3 + (ly * lx) * v4;
                grad_w[pi * L + l] += g * val;
                grad_loc[pi * 3 + 0] += (W - 1) * gx * gv;
                grad_loc[pi * 3 + 1] += (H - 1) * gy * gv;
            }
          }
    return 0;
}
