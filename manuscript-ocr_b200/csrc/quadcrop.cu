// Rectified crops of rotated word quads -> resize-and-pad -> normalise -> CHW float32 batch.
//
// NOT a reference behaviour: the reference crops the axis-aligned bounding rectangle of every polygon
// (_pipeline.py:204-221) and lists the rotated crop as future work (todo.md:1).  The operation is defined through the
// OpenCV calls a maintainer would write,
//     w, h  = round-half-even of the longer of each pair of opposite quad edges (float64); no patch if a side is < 2
//             or > 32767
//     Minv  = cv2.getPerspectiveTransform([[0,0],[w-1,0],[w-1,h-1],[0,h-1]], quad)
//     patch = cv2.warpPerspective(page, Minv, (w, h), INTER_LINEAR | WARP_INVERSE_MAP, borderMode, borderValue)
// followed by exactly the ResizeAndPadA + normalise of crop.cu, and restated bit for bit from OpenCV's
// imgproc/imgwarp.cpp + core/matrix_decomp.cpp (checked against cv2 4.13 via tests/golden/quad_warp.npz):
//   * the 8x8 system with float32 products in columns 6-7, LU with partial pivoting in float64 (eps = 100 ulp)
//   * destination coordinates evaluated in blocks of bw0 columns: X0 = M0*bx + M1*y + M2, then (X0 + M0*x1) * (32/W),
//     rounded half-to-even to 5 fractional bits
//   * 15-bit bilinear weights (32-ay)(32-ax)*32 ..., with the integer-position entry {32767, 0, 0, 1};
//     (sum + 2^14) >> 15; BORDER_CONSTANT substitutes the value per tap, BORDER_REPLICATE clamps the tap.
//
// Kernels: quad_plan_kernel (thread per quad: size, homography, resize plan) and quad_crop_kernel (CTA per quad:
// warp the patch into shared memory -- or evaluate it on the fly when it is larger than the stage -- then the same
// per-pixel resampling as crop_generic_kernel).  HBM traffic per quad: the touched page pixels + 3*ih*iw*4 bytes out.
#include "crop_common.cuh"

namespace {

constexpr int kQcThreads = 256;
constexpr int kQcPatchBytes = 40 * 1024;  // shared-memory patch stage (e.g. 260 x 52 pixels)
constexpr int kQcMaxTab = 384;            // AxisEnt entries (nw + nh of a 64 x 256 canvas is 320)

struct QuadPlan {
    double m[9];  // patch (x, y, 1) -> page (X, Y, W)
    int page, w, h, bw0;
    int ok, pad0, pad1, pad2;
};

// oracle quad_patch_size: edge lengths in float64 from the float32 vertices
__device__ __forceinline__ bool quad_patch_size(const float *q, int &w, int &h)
{
    double x[4], y[4];
#pragma unroll
    for (int v = 0; v < 4; v++) {
        x[v] = (double)q[2 * v];
        y[v] = (double)q[2 * v + 1];
    }
    auto edge = [&](int a, int b) {
        const double dx = x[b] - x[a], dy = y[b] - y[a];
        return sqrt(dx * dx + dy * dy);
    };
    const double e01 = edge(0, 1), e32 = edge(3, 2), e03 = edge(0, 3), e12 = edge(1, 2);
    const double ew = e32 > e01 ? e32 : e01, eh = e12 > e03 ? e12 : e03;  // Python max(a, b)
    w = h = 0;
    if (!(isfinite(e01) && isfinite(e32) && isfinite(e03) && isfinite(e12))) return false;
    if (ew > 32767.0 || eh > 32767.0) return false;  // beyond the 16-bit coordinates of the remap
    w = (int)rint(ew);
    h = (int)rint(eh);
    return w >= 2 && h >= 2;
}

// cv2.getPerspectiveTransform(rect(w,h), quad): imgwarp.cpp builds the system, matrix_decomp.cpp LUImpl solves it
__device__ bool perspective_from_rect(const float *q, int w, int h, double *m)
{
    const float sx[4] = {0.f, (float)(w - 1), (float)(w - 1), 0.f};
    const float sy[4] = {0.f, 0.f, (float)(h - 1), (float)(h - 1)};
    double a[8][8], b[8];
    for (int i = 0; i < 4; i++) {
        const float dx = q[2 * i], dy = q[2 * i + 1];
        a[i][0] = a[i + 4][3] = sx[i];
        a[i][1] = a[i + 4][4] = sy[i];
        a[i][2] = a[i + 4][5] = 1.0;
        a[i][3] = a[i][4] = a[i][5] = a[i + 4][0] = a[i + 4][1] = a[i + 4][2] = 0.0;
        a[i][6] = (double)(-sx[i] * dx);  // float32 products (Point2f arithmetic)
        a[i][7] = (double)(-sy[i] * dx);
        a[i + 4][6] = (double)(-sx[i] * dy);
        a[i + 4][7] = (double)(-sy[i] * dy);
        b[i] = dx;
        b[i + 4] = dy;
    }
    const double eps = 2.220446049250313e-16 * 100;
    for (int i = 0; i < 8; i++) {
        int k = i;
        for (int j = i + 1; j < 8; j++)
            if (fabs(a[j][i]) > fabs(a[k][i])) k = j;
        if (fabs(a[k][i]) < eps) return false;
        if (k != i) {
            for (int j = i; j < 8; j++) {
                const double t = a[i][j];
                a[i][j] = a[k][j];
                a[k][j] = t;
            }
            const double t = b[i];
            b[i] = b[k];
            b[k] = t;
        }
        const double d = -1 / a[i][i];
        for (int j = i + 1; j < 8; j++) {
            const double alpha = a[j][i] * d;
            for (int c = i + 1; c < 8; c++) a[j][c] += alpha * a[i][c];
            b[j] += alpha * b[i];
        }
    }
    for (int i = 7; i >= 0; i--) {
        double s = b[i];
        for (int c = i + 1; c < 8; c++) s -= a[i][c] * b[c];
        b[i] = s / a[i][i];
    }
    for (int i = 0; i < 8; i++) m[i] = b[i];
    m[8] = 1.0;
    return true;
}

__global__ void __launch_bounds__(128) quad_plan_kernel(const float *__restrict__ quads, int quad_stride,
                                                        const int32_t *__restrict__ page_of, int64_t n, int n_pages,
                                                        int min_text_size, int ih, int iw, QuadPlan *__restrict__ qplans,
                                                        Plan *__restrict__ plans, int32_t *__restrict__ sizes_out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float *q = quads + i * quad_stride;
        QuadPlan qp;
        Plan p;
        p.page = p.x1 = p.y1 = p.w = p.h = p.nw = p.nh = p.y0 = p.interp = p.isx = p.isy = 0;
        p.ok = p.staged = p.fast = p.pitch = p.stride = 0;
        p.src = nullptr;
        p.scale_x = p.scale_y = 1.0;
        qp.page = page_of ? page_of[i] : 0;
        qp.ok = qp.pad0 = qp.pad1 = qp.pad2 = 0;
        qp.bw0 = 1;
#pragma unroll
        for (int k = 0; k < 9; k++) qp.m[k] = 0.0;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = q[k];
        bool ok = qp.page >= 0 && qp.page < n_pages && quad_patch_size(v, qp.w, qp.h);
        ok = ok && qp.w >= min_text_size && qp.h >= min_text_size;  // _pipeline.py:130-133 applied to the patch
        ok = ok && perspective_from_rect(v, qp.w, qp.h, qp.m);
        if (ok) {
            // WarpPerspectiveInvoker's block width for a (w, h) destination
            const int bh0 = min(16, qp.h);
            qp.bw0 = min(1024 / bh0, qp.w);
            qp.ok = 1;
            p.w = qp.w;
            p.h = qp.h;
            p.ok = 1;
            plan_resize(qp.w, qp.h, ih, iw, p);
            p.staged = (size_t)qp.w * qp.h * 3 <= (size_t)kQcPatchBytes;  // the patch fits the shared-memory stage
        } else {
            qp.w = qp.h = 0;
        }
        qplans[i] = qp;
        plans[i] = p;
        if (sizes_out) {
            sizes_out[2 * i] = qp.w;
            sizes_out[2 * i + 1] = qp.h;
        }
    }
}

struct PageView {
    const uint8_t *px;  // (img_h, img_w, 3)
    int img_h, img_w, replicate, bval;
};

// remapBilinear's fixed-point interpolation at the 5-fractional-bit position (X, Y)
__device__ __forceinline__ void sample_px(const PageView &pg, int X, int Y, int &c0, int &c1, int &c2)
{
    const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);
    const int ax = X & 31, ay = Y & 31;
    int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32, w10 = ay * (32 - ax) * 32, w11 = ay * ax * 32;
    if ((ax | ay) == 0) {  // initInterTab2D: saturate_cast<short>(32768) = 32767, the missing unit goes to tap (1,1)
        w00 = 32767;
        w11 = 1;
    }
    const int H = pg.img_h, Wd = pg.img_w;
    int t0[3], t1[3], t2[3], t3[3];
    if (sx >= 0 && sy >= 0 && sx + 1 < Wd && sy + 1 < H) {  // all four taps inside: two 6-byte runs
        // a page is at most 32767 x 32767 x 3 bytes: the offset fits 32 bits
        const uint8_t *s = pg.px + ((uint32_t)sy * (uint32_t)Wd + (uint32_t)sx) * 3u;
        const uint8_t *r = s + (uint32_t)Wd * 3u;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            t0[c] = s[c];
            t1[c] = s[3 + c];
            t2[c] = r[c];
            t3[c] = r[3 + c];
        }
    } else {
        auto tap = [&](int yy, int xx, int *t) {
            const bool inside = yy >= 0 && yy < H && xx >= 0 && xx < Wd;
            if (inside || pg.replicate) {
                yy = min(max(yy, 0), H - 1);
                xx = min(max(xx, 0), Wd - 1);
                const uint8_t *s = pg.px + ((uint32_t)yy * (uint32_t)Wd + (uint32_t)xx) * 3u;
                t[0] = s[0];
                t[1] = s[1];
                t[2] = s[2];
            } else {
                t[0] = t[1] = t[2] = pg.bval;
            }
        };
        tap(sy, sx, t0);
        tap(sy, sx + 1, t1);
        tap(sy + 1, sx, t2);
        tap(sy + 1, sx + 1, t3);
    }
    auto mix = [&](int c) { return (t0[c] * w00 + t1[c] * w01 + t2[c] * w10 + t3[c] * w11 + (1 << 14)) >> 15; };
    // the weights are non-negative and sum to 2^15: the result is already within [0, 255]
    c0 = mix(0);
    c1 = mix(1);
    c2 = mix(2);
}

// WarpPerspectiveInvoker's destination -> source coordinates: (X0, Y0, W0) are the terms of the block start
// (bx, y), x1 the column inside the block
__device__ __forceinline__ void warp_xy(const QuadPlan &qp, double X0, double Y0, double W0, int x1, int &X, int &Y)
{
    double W = W0 + qp.m[6] * x1;
    W = W != 0.0 ? 32.0 / W : 0.0;
    const double fX = (X0 + qp.m[0] * x1) * W, fY = (Y0 + qp.m[3] * x1) * W;
    // saturate_cast<int>(std::max((double)INT_MIN, std::min((double)INT_MAX, v))): the conversion instruction already
    // saturates to [INT_MIN, INT_MAX] (round half to even, as cvRound); a NaN comes out of the min/max as INT_MAX
    X = fX != fX ? INT_MAX : __double2int_rn(fX);
    Y = fY != fY ? INT_MAX : __double2int_rn(fY);
}

__device__ __forceinline__ void block_terms(const QuadPlan &qp, int bx, int y, double &X0, double &Y0, double &W0)
{
    X0 = (qp.m[0] * bx + qp.m[1] * y) + qp.m[2];
    Y0 = (qp.m[3] * bx + qp.m[4] * y) + qp.m[5];
    W0 = (qp.m[6] * bx + qp.m[7] * y) + qp.m[8];
}

// one patch pixel at (x, y)
__device__ __forceinline__ void warp_px(const QuadPlan &qp, const PageView &pg, int x, int y, int &c0, int &c1, int &c2)
{
    const int bx = (x / qp.bw0) * qp.bw0;
    double X0, Y0, W0;
    block_terms(qp, bx, y, X0, Y0, W0);
    int X, Y;
    warp_xy(qp, X0, Y0, W0, x - bx, X, Y);
    sample_px(pg, X, Y, c0, c1, c2);
}

// rows [row0, row0 + step, ...) of the patch by one warp each, lanes along x inside every coordinate block: no
// integer division, the block terms are evaluated once per (row, block)
__device__ __forceinline__ void warp_rows(const QuadPlan &qp, const PageView &pg, unsigned char *patch, size_t pitch,
                                          int row0, int row_step, int lane)
{
    for (int y = row0; y < qp.h; y += row_step) {
        unsigned char *dst = patch + (size_t)y * pitch;
        for (int bx = 0; bx < qp.w; bx += qp.bw0) {
            double X0, Y0, W0;
            block_terms(qp, bx, y, X0, Y0, W0);
            const int bw = min(qp.bw0, qp.w - bx);
            for (int x1 = lane; x1 < bw; x1 += 32) {
                int X, Y, c0, c1, c2;
                warp_xy(qp, X0, Y0, W0, x1, X, Y);
                sample_px(pg, X, Y, c0, c1, c2);
                unsigned char *d = dst + (size_t)(bx + x1) * 3;
                d[0] = (unsigned char)c0;
                d[1] = (unsigned char)c1;
                d[2] = (unsigned char)c2;
            }
        }
    }
}

// patch too large for the stage: every tap is evaluated where it is used
struct WarpSrc {
    const QuadPlan *qp;
    PageView pg;
    __device__ __forceinline__ int at(int sy, int b) const
    {
        const int x = b / 3, c = b - 3 * x;
        int c0, c1, c2;
        warp_px(*qp, pg, x, sy, c0, c1, c2);
        return c == 0 ? c0 : (c == 1 ? c1 : c2);
    }
};

template <bool kWriteF32, bool kWriteU8, class Src>
__device__ __forceinline__ void resample_canvas(const Plan &p, const Src &src, const AxisEnt *s_tab, bool tab_ok,
                                                bool need_tab, int ih, int iw, float *dstf, uint8_t *dstu)
{
    const int plane = ih * iw;
    const float inv = 1.0f / 127.5f;
    const int nw = p.nw, nh = p.nh, npx = nw * nh;
    for (int t = threadIdx.x; t < npx; t += blockDim.x) {
        const int dy = t / nw, dx = t - dy * nw;
        AxisEnt ex = {}, ey = {};
        if (tab_ok) {
            ex = s_tab[dx];
            ey = s_tab[nw + dy];
        } else if (need_tab) {
            ex = p.interp == 3 ? area_entry(dx, p.scale_x, p.w) : linear_entry_x(dx, p.scale_x, p.w);
            ey = p.interp == 3 ? area_entry(dy, p.scale_y, p.h) : linear_entry_y(dy, p.scale_y, p.h);
        }
        unsigned char o0, o1, o2;
        resample_px(p, dx, dy, src, ex, ey, o0, o1, o2);
        const int at = (p.y0 + dy) * iw + dx;
        if (kWriteF32) {
            dstf[at] = ((float)o0 - 127.5f) * inv;
            dstf[plane + at] = ((float)o1 - 127.5f) * inv;
            dstf[2 * plane + at] = ((float)o2 - 127.5f) * inv;
        }
        if (kWriteU8) {
            dstu[(size_t)at * 3] = o0;
            dstu[(size_t)at * 3 + 1] = o1;
            dstu[(size_t)at * 3 + 2] = o2;
        }
    }
}

template <bool kWriteF32, bool kWriteU8>
__global__ void __launch_bounds__(kQcThreads, 4) quad_crop_kernel(const uint8_t *__restrict__ pages, int img_h,
                                                                  int img_w, const QuadPlan *__restrict__ qplans,
                                                                  const Plan *__restrict__ plans, int64_t n,
                                                                  int replicate, int bval, int ih, int iw,
                                                                  float *__restrict__ batch,
                                                                  uint8_t *__restrict__ canvas_out, int vec_ok)
{
    // [patch stage kQcPatchBytes + 16][tables: AxisEnt[kQcMaxTab] for resample_px, or the SoA tables of area4_strips]
    extern __shared__ __align__(16) unsigned char qc_smem[];
    unsigned char *qc_patch = qc_smem;
    AxisEnt *s_tab = reinterpret_cast<AxisEnt *>(qc_smem + kQcPatchBytes + 16);
    uint32_t *soa = reinterpret_cast<uint32_t *>(s_tab);
    __shared__ QuadPlan s_qp;
    const int plane = ih * iw, tab_n = area_tab_words(ih, iw);
    const bool soa_fits = (size_t)tab_n * sizeof(uint32_t) <= (size_t)kQcMaxTab * sizeof(AxisEnt);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t ci = blockIdx.x; ci < n; ci += gridDim.x) {
        __syncthreads();  // the previous quad's plan, patch and tables are no longer read
        if (threadIdx.x == 0) s_qp = qplans[ci];
        const Plan p = plans[ci];
        float *dstf = kWriteF32 ? batch + (size_t)ci * 3 * plane : nullptr;
        uint8_t *dstu = kWriteU8 ? canvas_out + (size_t)ci * 3 * plane : nullptr;
        write_padding<kWriteF32, kWriteU8>(p, ih, iw, dstf, dstu, vec_ok, threadIdx.x, blockDim.x);
        if (!p.ok) continue;  // uniform across the CTA
        // crop.cu's 4-tap INTER_AREA path applies to a staged patch shrunk by factors below 3
        const bool strips = p.staged && soa_fits && p.interp == 3 && p.scale_x < 2.999 && p.scale_y < 2.999;
        const bool need_tab = p.interp == 1 || p.interp == 3;
        const bool tab_ok = need_tab && p.nw + p.nh <= kQcMaxTab;
        bool x3_mine = true;
        if (strips) {
            x3_mine = build_tables(p, soa, tab_n, iw, threadIdx.x, kQcThreads);
        } else if (tab_ok) {
            for (int t = threadIdx.x; t < p.nw + p.nh; t += blockDim.x) {
                const bool isx = t < p.nw;
                const int d = isx ? t : t - p.nw;
                s_tab[t] = p.interp == 3 ? (isx ? area_entry(d, p.scale_x, p.w) : area_entry(d, p.scale_y, p.h))
                                         : (isx ? linear_entry_x(d, p.scale_x, p.w) : linear_entry_y(d, p.scale_y, p.h));
            }
        }
        const bool x3 = __syncthreads_and(x3_mine) != 0;  // also the barrier after the tables
        const QuadPlan &qp = s_qp;
        const PageView pg{pages + (size_t)qp.page * img_h * (size_t)img_w * 3, img_h, img_w, replicate, bval};
        if (p.staged) {
            const size_t pitch = (size_t)qp.w * 3;
            warp_rows(qp, pg, qc_patch, pitch, warp, kQcThreads / 32, lane);
            __syncthreads();
            bool redo = !strips;
            if (strips)
                redo = __syncthreads_or(
                    x3 ? area4_strips<kWriteF32, kWriteU8, kQcThreads, 3>(qc_smem, 0u, (uint32_t)pitch, 0u, 0u, soa, tab_n, ih,
                                                                          iw, p.nw, p.nh, p.y0, dstf, dstu, threadIdx.x)
                       : area4_strips<kWriteF32, kWriteU8, kQcThreads, 4>(qc_smem, 0u, (uint32_t)pitch, 0u, 0u, soa, tab_n, ih,
                                                                          iw, p.nw, p.nh, p.y0, dstf, dstu, threadIdx.x));
            if (redo) {
                if (strips) {  // an entry with more than 4 taps after all: per-pixel entries, no table
                    resample_canvas<kWriteF32, kWriteU8>(p, PitchedSrc{qc_patch, pitch}, s_tab, false, need_tab, ih, iw,
                                                         dstf, dstu);
                } else {
                    resample_canvas<kWriteF32, kWriteU8>(p, PitchedSrc{qc_patch, pitch}, s_tab, tab_ok, need_tab, ih,
                                                         iw, dstf, dstu);
                }
            }
        } else {
            const WarpSrc src{&qp, pg};
            resample_canvas<kWriteF32, kWriteU8>(p, src, s_tab, tab_ok, need_tab, ih, iw, dstf, dstu);
        }
    }
}

// the patch alone (row-major h x w x 3), for one planned quad
__global__ void __launch_bounds__(256) quad_warp_kernel(const uint8_t *__restrict__ page, int img_h, int img_w,
                                                        const QuadPlan *__restrict__ qplan, int replicate, int bval,
                                                        uint8_t *__restrict__ patch)
{
    __shared__ QuadPlan s_qp;
    if (threadIdx.x == 0) s_qp = *qplan;
    __syncthreads();
    const QuadPlan &qp = s_qp;
    if (!qp.ok) return;
    const PageView pg{page, img_h, img_w, replicate, bval};
    const int64_t npx = (int64_t)qp.w * qp.h;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < npx; t += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(t / qp.w), x = (int)(t - (int64_t)y * qp.w);
        int c0, c1, c2;
        warp_px(qp, pg, x, y, c0, c1, c2);
        patch[t * 3] = (unsigned char)c0;
        patch[t * 3 + 1] = (unsigned char)c1;
        patch[t * 3 + 2] = (unsigned char)c2;
    }
}

}  // namespace

size_t msk_quad_crop_scratch(int64_t n)
{
    const size_t k = (size_t)(n > 0 ? n : 0);
    return k * (sizeof(QuadPlan) + sizeof(Plan)) + 4096;
}

#define MS_TRY(expr)                 \
    do {                             \
        int _r = (expr);             \
        if (_r != MS_OK) return _r;  \
    } while (0)

static int quad_args_ok(int n_pages, int img_h, int img_w, int border_mode, int border_value, const char *who)
{
    if (n_pages <= 0 || img_h <= 0 || img_w <= 0 || img_h > 32767 || img_w > 32767) {
        ms_set_error("%s: pages must be 1..32767 pixels on a side (16-bit remap coordinates), got %dx%d", who, img_h, img_w);
        return MS_ERR_INVALID;
    }
    if (border_mode < 0 || border_mode > 1 || border_value < 0 || border_value > 255) {
        ms_set_error("%s: border_mode must be 0 (constant) or 1 (replicate), border_value 0..255", who);
        return MS_ERR_INVALID;
    }
    return MS_OK;
}

int msk_quad_crop(ms_ctx *ctx, const uint8_t *pages, int n_pages, int img_h, int img_w, const float *quads,
                  int quad_stride, const int32_t *page_of, int64_t n, int min_text_size, int border_mode,
                  int border_value, int out_h, int out_w, float *batch_f32, uint8_t *canvas_u8, int32_t *sizes_out,
                  ms_bump bump, cudaStream_t st)
{
    if (n <= 0) return MS_OK;
    MS_TRY(quad_args_ok(n_pages, img_h, img_w, border_mode, border_value, "quad_crop"));
    if (out_h <= 0 || out_w <= 0 || quad_stride < 8 || (!batch_f32 && !canvas_u8) ||
        (int64_t)out_h * out_w > (1 << 20)) {
        ms_set_error("quad_crop: bad arguments");
        return MS_ERR_INVALID;
    }
    QuadPlan *qplans = bump.take<QuadPlan>((size_t)n);
    Plan *plans = bump.take<Plan>((size_t)n);
    if (!qplans || !plans) {
        ms_set_error("quad_crop: scratch too small");
        return MS_ERR_CAPACITY;
    }
    int64_t pg = (n + 127) / 128;
    if (pg > (int64_t)ctx->num_sms * 16) pg = (int64_t)ctx->num_sms * 16;
    quad_plan_kernel<<<(int)pg, 128, 0, st>>>(quads, quad_stride, page_of, n, n_pages, min_text_size, out_h, out_w,
                                             qplans, plans, sizes_out);
    MS_LAUNCH_CHECK(ctx);
    int64_t grid = (int64_t)ctx->num_sms * 4;  // 40 KB stage + 12 KB tables, 64 registers: four CTAs per SM
    if (grid > n) grid = n;
    const int vec_ok = ((out_w & 3) == 0 && (reinterpret_cast<uintptr_t>(batch_f32) & 15) == 0) ? 1 : 0;
    const int smem = kQcPatchBytes + 16 + kQcMaxTab * (int)sizeof(AxisEnt);
    // cudaFuncSetAttribute is a synchronous driver call: once per context and kernel
#define MS_QC_LAUNCH(F32, U8)                                                                                          \
    do {                                                                                                               \
        auto kfn = quad_crop_kernel<F32, U8>;                                                                          \
        int &granted = ctx->smem_attr[4 + (F32 ? 1 : 0) + (U8 ? 2 : 0) - 1];                                           \
        if (smem > granted) {                                                                                          \
            MS_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                     \
            granted = smem;                                                                                            \
        }                                                                                                              \
        kfn<<<(int)grid, kQcThreads, smem, st>>>(pages, img_h, img_w, qplans, plans, n, border_mode, border_value,     \
                                                 out_h, out_w, batch_f32, canvas_u8, vec_ok);                         \
    } while (0)
    if (batch_f32 && canvas_u8)
        MS_QC_LAUNCH(true, true);
    else if (batch_f32)
        MS_QC_LAUNCH(true, false);
    else
        MS_QC_LAUNCH(false, true);
#undef MS_QC_LAUNCH
    MS_LAUNCH_CHECK(ctx);
    return MS_OK;
}

// One quad's patch.  *w / *h receive the patch size (0, 0: no patch); the pixels are produced only when they fit
// patch_cap bytes (device memory).  Synchronises `st` to read the size back.
int msk_quad_warp(ms_ctx *ctx, const uint8_t *page, int img_h, int img_w, const float *quad_dev, int border_mode,
                  int border_value, uint8_t *patch_dev, size_t patch_cap, int *w, int *h, ms_bump bump, cudaStream_t st)
{
    MS_TRY(quad_args_ok(1, img_h, img_w, border_mode, border_value, "quad_warp"));
    QuadPlan *qplan = bump.take<QuadPlan>(1);
    Plan *plan = bump.take<Plan>(1);
    int32_t *size = bump.take<int32_t>(2);
    if (!qplan || !plan || !size) {
        ms_set_error("quad_warp: scratch too small");
        return MS_ERR_CAPACITY;
    }
    quad_plan_kernel<<<1, 128, 0, st>>>(quad_dev, 8, nullptr, 1, 1, 0, 32, 128, qplan, plan, size);
    MS_LAUNCH_CHECK(ctx);
    int32_t hs[2] = {0, 0};
    MS_CUDA(cudaMemcpyAsync(hs, size, sizeof(hs), cudaMemcpyDeviceToHost, st));
    MS_CUDA(cudaStreamSynchronize(st));
    *w = hs[0];
    *h = hs[1];
    const size_t need = (size_t)hs[0] * hs[1] * 3;
    if (need == 0) return MS_OK;
    if (need > patch_cap) {
        ms_set_error("quad_warp: the %dx%d patch needs %zu bytes, capacity %zu", hs[0], hs[1], need, patch_cap);
        return MS_ERR_CAPACITY;
    }
    int64_t grid = ((int64_t)hs[0] * hs[1] + 255) / 256;
    if (grid > (int64_t)ctx->num_sms * 8) grid = (int64_t)ctx->num_sms * 8;
    quad_warp_kernel<<<(int)grid, 256, 0, st>>>(page, img_h, img_w, qplan, border_mode, border_value, patch_dev);
    MS_LAUNCH_CHECK(ctx);
    return MS_OK;
}
