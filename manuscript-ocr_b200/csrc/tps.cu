// TPS rectification of the recogniser batch: thin-plate-spline grid generation + bilinear grid_sample in one kernel.
//
// NOT a reference behaviour ("parity unpinned"): the reference's TRBAModel has no transformation stage -- no TPS / STN,
// no grid_sample anywhere (recognizers/_trba/model/model.py:338-393).  BASELINE.json's north_star names "the
// TPS rectification grid_sample", so the operation is specified here as the TPS-STN of the TRBA literature (Baek et
// al. 2019, "What is wrong with scene text recognition model comparisons?", the T of T-R-B-A; GridGenerator +
// F.grid_sample(padding_mode="border", align_corners=True) in its public implementation) and checked against
// torch.nn.functional.grid_sample in tests/ within a stated float32 tolerance.
//
//   T      = inv_delta_C (F+3, F+3) @ [C' (F, 2); 0 (3, 2)]            per batch item   (C' = predicted fiducial points)
//   P'     = P_hat (n, F+3) @ T                                         n = out_h * out_w  sampling positions in [-1, 1]
//   out    = grid_sample(input, P', bilinear, border padding, align_corners = True)
//
// inv_delta_C and P_hat depend only on (F, out_h, out_w); the host computes them once (float64, stored float32, P_hat
// TRANSPOSED to (F+3, n) so that a warp's pixels read consecutive floats); the two small matrix products are
// accumulated in float64 (their terms cancel), the bilinear interpolation in float32 like torch's.  One CTA = one batch item x 256 output
// pixels; T lives in shared memory.  HBM roofline: C*in_h*in_w*4 bytes read + C*out_h*out_w*4 written per item (P_hat,
// 94 KB for 32x100, stays in L2).
#include "ms_internal.cuh"

namespace {

constexpr int kTpsThreads = 256;
constexpr int kTpsMaxK = 64;  // F + 3

__global__ void __launch_bounds__(kTpsThreads) tps_rectify_kernel(const float *__restrict__ input,
                                                                   const float *__restrict__ c_prime,
                                                                   const float *__restrict__ inv_delta_c,
                                                                   const float *__restrict__ p_hat_t, int n_fid, int chans,
                                                                   int in_h, int in_w, int out_h, int out_w,
                                                                   float *__restrict__ out)
{
    ms_pdl_wait();
    __shared__ double s_t[kTpsMaxK][2];
    const int b = blockIdx.y, K = n_fid + 3, n = out_h * out_w;
    // T = inv_delta_C[:, :F] @ C'   (the three appended rows of zeros contribute nothing)
    for (int e = threadIdx.x; e < K * 2; e += kTpsThreads) {
        const int k = e >> 1, d = e & 1;
        double acc = 0.0;  // float64 accumulation: the TPS sums cancel (terms of ~10 for a result in [-1, 1])
        for (int j = 0; j < n_fid; j++)
            acc += (double)inv_delta_c[k * K + j] * (double)c_prime[((size_t)b * n_fid + j) * 2 + d];
        s_t[k][d] = acc;
    }
    __syncthreads();
    const int pix = blockIdx.x * kTpsThreads + threadIdx.x;
    if (pix >= n) return;
    double gxd = 0.0, gyd = 0.0;
    for (int k = 0; k < K; k++) {
        const double p = (double)p_hat_t[(size_t)k * n + pix];
        gxd += p * s_t[k][0];
        gyd += p * s_t[k][1];
    }
    const float gx = (float)gxd, gy = (float)gyd;
    // grid_sample, align_corners = True: -1 -> pixel 0, +1 -> pixel size - 1; border padding clamps the coordinate
    float ix = (gx + 1.f) * 0.5f * (float)(in_w - 1), iy = (gy + 1.f) * 0.5f * (float)(in_h - 1);
    ix = fminf(fmaxf(ix, 0.f), (float)(in_w - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(in_h - 1));
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const bool vx1 = x1 <= in_w - 1, vy1 = y1 <= in_h - 1;  // (x0, y0) is always inside after the clamp
    const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
    const size_t plane_in = (size_t)in_h * in_w;
    for (int c = 0; c < chans; c++) {
        const float *src = input + ((size_t)b * chans + c) * plane_in;
        float v = src[(size_t)y0 * in_w + x0] * w00;
        if (vx1) v += src[(size_t)y0 * in_w + x1] * w10;
        if (vy1) v += src[(size_t)y1 * in_w + x0] * w01;
        if (vx1 && vy1) v += src[(size_t)y1 * in_w + x1] * w11;
        __stcs(out + ((size_t)b * chans + c) * n + pix, v);
    }
}

}  // namespace

int msk_tps_rectify(ms_ctx *ctx, const float *input, const float *c_prime, const float *inv_delta_c, const float *p_hat_t,
                    int batch, int n_fid, int chans, int in_h, int in_w, int out_h, int out_w, float *out, cudaStream_t st)
{
    if (batch <= 0) return MS_OK;
    if (!input || !c_prime || !inv_delta_c || !p_hat_t || !out || n_fid < 1 || n_fid + 3 > kTpsMaxK || chans < 1 ||
        in_h < 1 || in_w < 1 || out_h < 1 || out_w < 1 || batch > 65535) {
        ms_set_error("tps_rectify: bad arguments (at most %d fiducial points, 65535 items per call)", kTpsMaxK - 3);
        return MS_ERR_INVALID;
    }
    const int n = out_h * out_w;
    dim3 grid((n + kTpsThreads - 1) / kTpsThreads, batch);
    ms_launch(tps_rectify_kernel, grid, kTpsThreads, 0, st, input, c_prime, inv_delta_c, p_hat_t, n_fid, chans, in_h, in_w, out_h,
                                                    out_w, out);
    MS_LAUNCH_CHECK(ctx);
    return MS_OK;
}
