// Crop -> resize-and-pad -> normalise -> CHW float32 batch, written directly as the TRBA input.
// Replaces Pipeline._extract_word_image (reference _pipeline.py:204-221), ResizeAndPadA.apply and
// get_val_transform (recognizers/_trba/data/transforms.py:62-120, 185-193) and the per-image
// `.to(device)` + torch.stack of TRBA.predict (recognizers/_trba/__init__.py:264-288, 382-390).
//
// cv2.resize arithmetic restated bit-exactly for uint8 x 3 channels (OpenCV imgproc/resize.cpp,
// validated against cv2 4.13 through the oracle):
//   INTER_LINEAR  (neither side shrinks): 11-bit fixed-point coefficients, vertical pass
//                 ((b0*(r0>>4))>>16 + (b1*(r1>>4))>>16 + 2) >> 2
//   INTER_AREA    (a side shrinks): integer-ratio fast path (2x2: (s+2)>>2, else round(sum*(1/area)))
//                 or the float decimation-table path, accumulated in OpenCV's order
//   same size:    copy
// then paste at (x0=0, y0=(ih-new_h)//2) on a 255 canvas, (v - 127.5) * (1/127.5) in f32, HWC -> CHW.
//
// HBM roofline: per crop 3*w*h bytes gathered + 3*ih*iw*4 bytes written (49 152 B at 32x128): the
// kernel is write dominated.  One CTA per crop (persistent grid); the canvas is assembled in shared
// memory and streamed out with 16-byte coalesced stores.
#include "ms_internal.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxEntries = 10;  // cached decimation-table entries per axis (longer tables are recomputed)

struct AreaAxis {
    // OpenCV computeResizeAreaTab for one destination index
    int s_first;        // source index of the first entry
    int n;              // number of entries
    float a_first, a_mid, a_last;
    bool has_first, has_last;
    int s_mid0, s_mid1;  // full-weight source range [s_mid0, s_mid1)
};

__device__ __forceinline__ AreaAxis area_axis(int d, double scale, int ssize)
{
    AreaAxis t;
    double f1 = d * scale, f2 = f1 + scale;
    double cell = fmin(scale, (double)ssize - f1);
    int s1 = (int)ceil(f1), s2 = (int)floor(f2);
    s2 = min(s2, ssize - 1);
    s1 = min(s1, s2);
    t.has_first = (s1 - f1) > 1e-3;
    t.a_first = (float)((s1 - f1) / cell);
    t.s_mid0 = s1;
    t.s_mid1 = s2;
    t.a_mid = (float)(1.0 / cell);
    t.has_last = (f2 - s2) > 1e-3;
    t.a_last = (float)(fmin(fmin(f2 - s2, 1.0), cell) / cell);
    t.s_first = t.has_first ? s1 - 1 : s1;
    t.n = (t.has_first ? 1 : 0) + (s2 > s1 ? s2 - s1 : 0) + (t.has_last ? 1 : 0);
    return t;
}

__device__ __forceinline__ float axis_weight(const AreaAxis &t, int e, int &src)
{
    // e-th entry in table order
    if (t.has_first) {
        if (e == 0) {
            src = t.s_mid0 - 1;
            return t.a_first;
        }
        e--;
    }
    int mid = t.s_mid1 - t.s_mid0;
    if (mid < 0) mid = 0;
    if (e < mid) {
        src = t.s_mid0 + e;
        return t.a_mid;
    }
    src = t.s_mid1;
    return t.a_last;
}

__device__ __forceinline__ int cv_round(float v) { return __float2int_rn(v); }
__device__ __forceinline__ unsigned char sat_u8(int v) { return (unsigned char)min(max(v, 0), 255); }

struct Plan {
    int w, h, nw, nh, x0, y0, interp;  // interp: 0 copy, 1 linear, 2 area-fast, 3 area-general
    double scale_x, scale_y;
    int isx, isy;
};

__device__ __forceinline__ Plan make_plan(int w, int h, int ih, int iw)
{
    Plan p;
    p.w = w;
    p.h = h;
    // transforms.py:91-95
    double s1 = (double)ih / (double)max(h, 1), s2 = (double)iw / (double)max(w, 1);
    double sc = fmin(s1, s2);
    p.nw = max(1, (int)rint(w * sc));  // python round(): half to even
    p.nh = max(1, (int)rint(h * sc));
    p.nw = min(p.nw, iw);
    p.nh = min(p.nh, ih);
    int shrink = (p.nh < h || p.nw < w);  // transforms.py:80-83
    p.x0 = 0;
    p.y0 = (ih - p.nh) / 2;
    p.y0 = max(0, min(p.y0, ih - p.nh));
    p.scale_x = 1.0 / ((double)p.nw / (double)w);
    p.scale_y = 1.0 / ((double)p.nh / (double)h);
    p.isx = (int)rint(p.scale_x);
    p.isy = (int)rint(p.scale_y);
    if (p.nw == w && p.nh == h)
        p.interp = 0;
    else if (!shrink)
        p.interp = 1;
    else {
        bool fast = fabs(p.scale_x - p.isx) < 2.220446049250313e-16 && fabs(p.scale_y - p.isy) < 2.220446049250313e-16;
        p.interp = fast ? 2 : 3;
    }
    return p;
}

template <bool kWriteF32, bool kWriteU8>
__global__ void __launch_bounds__(kThreads) crop_resize_pad_kernel(const uint8_t *__restrict__ pages, int n_pages,
                                                                   int img_h, int img_w,
                                                                   const int32_t *__restrict__ crops,
                                                                   const int32_t *__restrict__ n_crops_dev,
                                                                   int64_t crops_cap, int ih, int iw,
                                                                   float *__restrict__ batch,
                                                                   uint8_t *__restrict__ canvas_out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    float *lut = reinterpret_cast<float *>(smem);          // 256 floats
    unsigned char *canvas = smem + 256 * sizeof(float);      // ih*iw*3 bytes (padded to 16)
    const int canvas_bytes = ih * iw * 3;
    {
        const float inv = 1.0f / 127.5f;
        for (int v = threadIdx.x; v < 256; v += kThreads) lut[v] = ((float)v - 127.5f) * inv;
    }
    int64_t n_crops = *n_crops_dev;
    if (n_crops > crops_cap) n_crops = crops_cap;
    const size_t page_bytes = (size_t)img_h * img_w * 3;
    const size_t stride = (size_t)img_w * 3;

    for (int64_t ci = blockIdx.x; ci < n_crops; ci += gridDim.x) {
        __syncthreads();  // previous canvas fully consumed
        const int32_t *cr = crops + ci * 5;
        const int page = cr[0], x1 = cr[1], y1 = cr[2], x2 = cr[3], y2 = cr[4];
        const int w = x2 - x1, h = y2 - y1;
        bool ok = page >= 0 && page < n_pages && w > 0 && h > 0 && x1 >= 0 && y1 >= 0 && x2 <= img_w && y2 <= img_h;
        // 255 canvas (transforms.py:100)
        {
            uint32_t *c4 = reinterpret_cast<uint32_t *>(canvas);
            for (int i = threadIdx.x; i < (canvas_bytes + 3) / 4; i += kThreads) c4[i] = 0xFFFFFFFFu;
        }
        __syncthreads();
        if (ok) {
            const Plan p = make_plan(w, h, ih, iw);
            const uint8_t *src = pages + (size_t)page * page_bytes + (size_t)y1 * stride + (size_t)x1 * 3;
            const int npx = p.nw * p.nh;
            for (int t = threadIdx.x; t < npx; t += kThreads) {
                const int dy = t / p.nw, dx = t - dy * p.nw;
                unsigned char o0, o1, o2;
                if (p.interp == 0) {
                    const uint8_t *s = src + (size_t)dy * stride + dx * 3;
                    o0 = s[0];
                    o1 = s[1];
                    o2 = s[2];
                } else if (p.interp == 1) {
                    float fx = (float)((dx + 0.5) * p.scale_x - 0.5);
                    int sx = (int)floorf(fx);
                    fx -= sx;
                    if (sx < 0) {
                        fx = 0;
                        sx = 0;
                    }
                    bool edge = sx + 1 >= w;
                    if (edge) {
                        fx = 0;
                        sx = w - 1;
                    }
                    const int a0 = (short)cv_round((1.f - fx) * 2048.f), a1 = (short)cv_round(fx * 2048.f);
                    float fy = (float)((dy + 0.5) * p.scale_y - 0.5);
                    int sy = (int)floorf(fy);
                    fy -= sy;
                    const int b0 = (short)cv_round((1.f - fy) * 2048.f), b1 = (short)cv_round(fy * 2048.f);
                    const int ya = min(max(sy, 0), h - 1), yb = min(max(sy + 1, 0), h - 1);
                    const uint8_t *S0 = src + (size_t)ya * stride + sx * 3;
                    const uint8_t *S1 = src + (size_t)yb * stride + sx * 3;
                    unsigned char o[3];
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        int r0, r1;
                        if (edge) {
                            r0 = S0[c] * 2048;
                            r1 = S1[c] * 2048;
                        } else {
                            r0 = S0[c] * a0 + S0[c + 3] * a1;
                            r1 = S1[c] * a0 + S1[c + 3] * a1;
                        }
                        o[c] = (unsigned char)((((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2);
                    }
                    o0 = o[0];
                    o1 = o[1];
                    o2 = o[2];
                } else if (p.interp == 2) {
                    int sum[3] = {0, 0, 0};
                    for (int yy = 0; yy < p.isy; yy++) {
                        const uint8_t *s = src + (size_t)(dy * p.isy + yy) * stride + (size_t)dx * p.isx * 3;
                        for (int xx = 0; xx < p.isx; xx++) {
                            sum[0] += s[xx * 3];
                            sum[1] += s[xx * 3 + 1];
                            sum[2] += s[xx * 3 + 2];
                        }
                    }
                    if (p.isx == 2 && p.isy == 2) {
                        o0 = (unsigned char)((sum[0] + 2) >> 2);
                        o1 = (unsigned char)((sum[1] + 2) >> 2);
                        o2 = (unsigned char)((sum[2] + 2) >> 2);
                    } else {
                        const float inv = 1.f / (float)(p.isx * p.isy);
                        o0 = sat_u8(cv_round((float)sum[0] * inv));
                        o1 = sat_u8(cv_round((float)sum[1] * inv));
                        o2 = sat_u8(cv_round((float)sum[2] * inv));
                    }
                } else {
                    const AreaAxis tx = area_axis(dx, p.scale_x, w);
                    const AreaAxis ty = area_axis(dy, p.scale_y, h);
                    float wx[kMaxEntries];
                    const bool cached = tx.n <= kMaxEntries;
                    if (cached) {
#pragma unroll
                        for (int e = 0; e < kMaxEntries; e++) {
                            int sxi;
                            wx[e] = e < tx.n ? axis_weight(tx, e, sxi) : 0.f;
                        }
                    }
                    float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f;
                    for (int j = 0; j < ty.n; j++) {
                        int sy;
                        const float beta = axis_weight(ty, j, sy);
                        const uint8_t *s = src + (size_t)sy * stride + (size_t)tx.s_first * 3;
                        float b0 = 0.f, b1 = 0.f, b2 = 0.f;
                        if (cached) {
#pragma unroll
                            for (int e = 0; e < kMaxEntries; e++) {
                                if (e < tx.n) {
                                    const float a = wx[e];
                                    b0 = b0 + (float)s[e * 3] * a;
                                    b1 = b1 + (float)s[e * 3 + 1] * a;
                                    b2 = b2 + (float)s[e * 3 + 2] * a;
                                }
                            }
                        } else {
                            for (int e = 0; e < tx.n; e++) {
                                int sxi;
                                const float a = axis_weight(tx, e, sxi);
                                b0 = b0 + (float)s[e * 3] * a;
                                b1 = b1 + (float)s[e * 3 + 1] * a;
                                b2 = b2 + (float)s[e * 3 + 2] * a;
                            }
                        }
                        if (j == 0) {
                            sum0 = beta * b0;
                            sum1 = beta * b1;
                            sum2 = beta * b2;
                        } else {
                            sum0 += beta * b0;
                            sum1 += beta * b1;
                            sum2 += beta * b2;
                        }
                    }
                    o0 = sat_u8(cv_round(sum0));
                    o1 = sat_u8(cv_round(sum1));
                    o2 = sat_u8(cv_round(sum2));
                }
                unsigned char *d = canvas + ((size_t)(p.y0 + dy) * iw + (p.x0 + dx)) * 3;
                d[0] = o0;
                d[1] = o1;
                d[2] = o2;
            }
        }
        __syncthreads();
        // stream the canvas out: normalised CHW f32 (16-byte stores) and / or the raw HWC bytes
        if (kWriteF32) {
            float *dst = batch + (size_t)ci * 3 * ih * iw;
            const int plane = ih * iw;
            if ((iw & 3) == 0) {
                const int n4 = 3 * plane / 4;
                for (int i = threadIdx.x; i < n4; i += kThreads) {
                    int e = i * 4;
                    int c = e / plane, rem = e - c * plane;  // 4 consecutive x of one (c, y)
                    const unsigned char *s = canvas + (size_t)rem * 3 + c;
                    float4 v = make_float4(lut[s[0]], lut[s[3]], lut[s[6]], lut[s[9]]);
                    __stcs(reinterpret_cast<float4 *>(dst) + i, v);
                }
            } else {
                for (int e = threadIdx.x; e < 3 * plane; e += kThreads) {
                    int c = e / plane, rem = e - c * plane;
                    dst[e] = lut[canvas[(size_t)rem * 3 + c]];
                }
            }
        }
        if (kWriteU8) {
            uint8_t *dst = canvas_out + (size_t)ci * canvas_bytes;
            if ((canvas_bytes & 3) == 0) {
                const uint32_t *c4 = reinterpret_cast<const uint32_t *>(canvas);
                uint32_t *d4 = reinterpret_cast<uint32_t *>(dst);
                for (int i = threadIdx.x; i < canvas_bytes / 4; i += kThreads) d4[i] = c4[i];
            } else {
                for (int i = threadIdx.x; i < canvas_bytes; i += kThreads) dst[i] = canvas[i];
            }
        }
    }
}

}  // namespace

int msk_crop(ms_ctx *ctx, const uint8_t *pages, int n_pages, int img_h, int img_w, const int32_t *crops,
             const int32_t *n_crops, int64_t crops_cap, int out_h, int out_w, float *batch_f32, uint8_t *canvas_u8,
             cudaStream_t st)
{
    if (crops_cap <= 0 || n_pages <= 0) return MS_OK;
    if (out_h <= 0 || out_w <= 0 || img_h <= 0 || img_w <= 0 || (!batch_f32 && !canvas_u8)) {
        ms_set_error("crop: bad arguments");
        return MS_ERR_INVALID;
    }
    size_t smem = 256 * sizeof(float) + (((size_t)out_h * out_w * 3 + 15) & ~size_t(15));
    if (smem > 200 * 1024) {
        ms_set_error("crop: canvas %dx%d does not fit in shared memory", out_h, out_w);
        return MS_ERR_INVALID;
    }
    int per_sm = (int)((200 * 1024) / smem);
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)ctx->num_sms * per_sm;
    if (grid > crops_cap) grid = crops_cap;
#define MS_CROP_LAUNCH(F32, U8)                                                                                      \
    do {                                                                                                             \
        auto kfn = crop_resize_pad_kernel<F32, U8>;                                                                  \
        if (smem > 48 * 1024) MS_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kfn<<<(int)grid, kThreads, smem, st>>>(pages, n_pages, img_h, img_w, crops, n_crops, crops_cap, out_h, out_w, \
                                              batch_f32, canvas_u8);                                                 \
    } while (0)
    if (batch_f32 && canvas_u8)
        MS_CROP_LAUNCH(true, true);
    else if (batch_f32)
        MS_CROP_LAUNCH(true, false);
    else
        MS_CROP_LAUNCH(false, true);
#undef MS_CROP_LAUNCH
    MS_LAUNCH_CHECK(ctx);
    return MS_OK;
}
