/*
 * oracle.c -- CPU restatement of the manuscript-ocr detector->recognizer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under manuscript-ocr_b200/ may import, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and there only as the checker.
 *
 * Every function restates (does not copy) the algorithm of the reference
 * olegiy/manuscript-ocr v0.1.8; paths below are relative to the reference root.
 * Parity status: pinned.  tests/golden/make_golden.py runs the real reference
 * (numpy/numba/cv2) in the build container and commits its outputs; tests/test_oracle_*.py
 * compare this file against those vectors and against the known-answer cases of
 * tests/detectors/east/test_lanms.py:18-188.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math  (see oracle/build.py).
 * -ffp-contract=off matters: numba (no fastmath) and numpy never fuse a*b+c, so we must not.
 *
 * Tie rule (documented deviation): the reference sorts with numpy's default
 * *unstable* argsort (lanms.py:138,167) whose tie order is platform dependent.  This
 * oracle, and the CUDA path, break ties by original index (== numpy kind="stable").
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXV 20 /* lanms.py:34 -- clip buffer of 20 vertices */

/* ------------------------------------------------------------------------- */
/* polygon primitives: lanms.py:7-130 (all float64)                           */
/* ------------------------------------------------------------------------- */

/* lanms.py:7-14  shoelace, sequential accumulation from 0.0, |.|/2 */
double orc_polygon_area(const double *p, int n)
{
    double acc = 0.0;
    for (int i = 0; i < n; i++) {
        int j = (i + 1) % n;
        acc += p[2 * i] * p[2 * j + 1] - p[2 * j] * p[2 * i + 1];
    }
    return fabs(acc) / 2.0;
}

/* lanms.py:17-29  segment p1->p2 against the line through A,B; parallel -> p1 */
void orc_line_hit(const double *p1, const double *p2, const double *A, const double *B, double *out)
{
    double ex = p2[0] - p1[0], ey = p2[1] - p1[1];
    double lx = B[0] - A[0], ly = B[1] - A[1];
    double den = ex * ly - ey * lx;
    double ax = A[0] - p1[0], ay = A[1] - p1[1];
    if (den == 0) {
        out[0] = p1[0];
        out[1] = p1[1];
        return;
    }
    double t = (ax * ly - ay * lx) / den;
    out[0] = p1[0] + t * ex;
    out[1] = p1[1] + t * ey;
}

static inline int left_of(const double *A, const double *B, const double *P)
{
    /* lanms.py:40-45  inside <=> cross >= 0 */
    return (B[0] - A[0]) * (P[1] - A[1]) - (B[1] - A[1]) * (P[0] - A[0]) >= 0;
}

/* lanms.py:32-57  one Sutherland-Hodgman step; returns vertex count */
int orc_clip_halfplane(const double *sub, int n, const double *A, const double *B, double *out)
{
    int cnt = 0;
    for (int i = 0; i < n; i++) {
        const double *cur = sub + 2 * i;
        const double *prv = sub + 2 * ((i + n - 1) % n);
        int cin = left_of(A, B, cur);
        int pin = left_of(A, B, prv);
        if (cin) {
            if (!pin) {
                orc_line_hit(prv, cur, A, B, out + 2 * cnt);
                cnt++;
            }
            out[2 * cnt] = cur[0];
            out[2 * cnt + 1] = cur[1];
            cnt++;
        } else if (pin) {
            orc_line_hit(prv, cur, A, B, out + 2 * cnt);
            cnt++;
        }
    }
    return cnt;
}

/* lanms.py:60-77  clip poly1 by every directed edge of poly2, stop when empty */
int orc_polygon_intersection(const double *p1, int n1, const double *p2, int n2, double *out)
{
    double a[2 * ORC_MAXV], b[2 * ORC_MAXV];
    double *cur = a, *nxt = b;
    int n = n1;
    memcpy(cur, p1, sizeof(double) * 2 * n1);
    for (int i = 0; i < n2; i++) {
        const double *A = p2 + 2 * i;
        const double *B = p2 + 2 * ((i + 1) % n2);
        n = orc_clip_halfplane(cur, n, A, B, nxt);
        double *t = cur;
        cur = nxt;
        nxt = t;
        if (n == 0)
            break;
    }
    memcpy(out, cur, sizeof(double) * 2 * n);
    return n;
}

/* lanms.py:80-91 */
double orc_polygon_iou(const double *p1, const double *p2)
{
    double ip[2 * ORC_MAXV];
    int n = orc_polygon_intersection(p1, 4, p2, 4, ip);
    double ia = 0.0;
    if (n > 2)
        ia = orc_polygon_area(ip, n);
    double a1 = orc_polygon_area(p1, 4);
    double a2 = orc_polygon_area(p2, 4);
    double uni = a1 + a2 - ia;
    if (uni <= 0)
        return 0.0;
    return ia / uni;
}

/* lanms.py:94-96  strict > */
/* n pairs at once (tests compare 10^6 pairs with the device predicates) */
void orc_polygon_iou_batch(const double *p1, const double *p2, long long n, double *out)
{
    for (long long i = 0; i < n; i++)
        out[i] = orc_polygon_iou(p1 + 8 * i, p2 + 8 * i);
}

int orc_should_merge(const double *p1, const double *p2, double thr)
{
    return orc_polygon_iou(p1, p2) > thr;
}

/* lanms.py:99-130  best of 4 rotations x 2 directions, first strict minimum wins */
void orc_normalize_polygon(const double *ref, const double *poly, double *out)
{
    int best_dir = 0, best_start = 0;
    double best = 1e20;
    for (int s = 0; s < 4; s++) {
        double d = 0.0;
        for (int i = 0; i < 4; i++) {
            int k = (s + i) % 4;
            double dx = ref[2 * i] - poly[2 * k];
            double dy = ref[2 * i + 1] - poly[2 * k + 1];
            d += dx * dx + dy * dy;
        }
        if (d < best) {
            best = d;
            best_start = s;
            best_dir = 0;
        }
    }
    for (int s = 0; s < 4; s++) {
        double d = 0.0;
        for (int i = 0; i < 4; i++) {
            int k = ((s - i) % 4 + 4) % 4;
            double dx = ref[2 * i] - poly[2 * k];
            double dy = ref[2 * i + 1] - poly[2 * k + 1];
            d += dx * dx + dy * dy;
        }
        if (d < best) {
            best = d;
            best_start = s;
            best_dir = 1;
        }
    }
    for (int i = 0; i < 4; i++) {
        int k = best_dir == 0 ? (best_start + i) % 4 : ((best_start - i) % 4 + 4) % 4;
        out[2 * i] = poly[2 * k];
        out[2 * i + 1] = poly[2 * k + 1];
    }
}

/* ------------------------------------------------------------------------- */
/* stable index sorts (tie rule: original index ascending)                    */
/* ------------------------------------------------------------------------- */
typedef struct {
    double key;
    int64_t idx;
} orc_kv;

static int kv_cmp(const void *a, const void *b)
{
    const orc_kv *x = (const orc_kv *)a, *y = (const orc_kv *)b;
    /* numpy sort order: NaN last */
    int xn = isnan(x->key), yn = isnan(y->key);
    if (xn || yn) {
        if (xn != yn)
            return xn - yn;
    } else {
        if (x->key < y->key)
            return -1;
        if (x->key > y->key)
            return 1;
    }
    return (x->idx > y->idx) - (x->idx < y->idx);
}

static void stable_argsort(const double *keys, int64_t n, int64_t *order)
{
    orc_kv *kv = (orc_kv *)malloc(sizeof(orc_kv) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) {
        kv[i].key = keys[i];
        kv[i].idx = i;
    }
    qsort(kv, (size_t)n, sizeof(orc_kv), kv_cmp);
    for (int64_t i = 0; i < n; i++)
        order[i] = kv[i].idx;
    free(kv);
}

/* ------------------------------------------------------------------------- */
/* lanms.py:133-153  greedy NMS, descending score, subject = kept box         */
/* returns K; keep_idx[0..K) are indices into polys in kept (priority) order  */
/* ------------------------------------------------------------------------- */
int64_t orc_standard_nms(const double *polys, const double *scores, int64_t n, double thr, int64_t *keep_idx)
{
    if (n <= 0)
        return 0;
    double *neg = (double *)malloc(sizeof(double) * (size_t)n);
    int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    unsigned char *dead = (unsigned char *)calloc((size_t)n, 1);
    for (int64_t i = 0; i < n; i++)
        neg[i] = -scores[i];
    stable_argsort(neg, n, order);
    int64_t k = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t a = order[i];
        if (dead[a])
            continue;
        keep_idx[k++] = a;
        for (int64_t j = i + 1; j < n; j++) {
            int64_t b = order[j];
            if (dead[b])
                continue;
            if (orc_should_merge(polys + 8 * a, polys + 8 * b, thr))
                dead[b] = 1;
        }
    }
    free(neg);
    free(order);
    free(dead);
    return k;
}

/* ------------------------------------------------------------------------- */
/* RBOX geometry decode -- NOT a reference behaviour ("parity unpinned"): the reference's geometry head is QUAD only
 * (detectors/_east/east.py:99-100, utils.py:368-376).  BASELINE.json's north_star names it, so the closed
 * form of the EAST paper (Zhou et al. 2017) is stated here as the public implementations do (restore_rectangle_rbox):
 * geo5 = distances from the pixel to the top, right, bottom, left edge of the rotated rectangle and its angle.
 * Thresholding / quantisation / row order exactly as orc_decode_quads (utils.py:340-356).  float64 arithmetic. */
void orc_rbox_to_quad(double ox, double oy, double dt, double dr, double db, double dl, double ang, float *r)
{
    double px[4], py[4], qx, qy, c, s;
    if (ang >= 0.0) {
        px[0] = 0.0; py[0] = -dt - db;
        px[1] = dr + dl; py[1] = -dt - db;
        px[2] = dr + dl; py[2] = 0.0;
        px[3] = 0.0; py[3] = 0.0;
        qx = dl; qy = -db;
        c = cos(ang); s = sin(ang);
    } else {
        px[0] = -dr - dl; py[0] = -dt - db;
        px[1] = 0.0; py[1] = -dt - db;
        px[2] = 0.0; py[2] = 0.0;
        px[3] = -dr - dl; py[3] = 0.0;
        qx = -dr; qy = -db;
        c = cos(-ang); s = -sin(-ang);
    }
    double tx = ox - (c * qx + s * qy), ty = oy - (-s * qx + c * qy);
    for (int v = 0; v < 4; v++) {
        r[2 * v] = (float)(c * px[v] + s * py[v] + tx);
        r[2 * v + 1] = (float)(-s * px[v] + c * py[v] + ty);
    }
}

int64_t orc_decode_rbox(const float *score, const float *geo5, int H, int W, double thr, double scale, int q,
                        float *out, int64_t cap)
{
    if (q < 1) q = 1;
    int CH = (H + q - 1) / q, CW = (W + q - 1) / q;
    float thr32 = (float)thr;
    int64_t n = 0;
    size_t plane = (size_t)H * W;
    for (int cy = 0; cy < CH; cy++)
        for (int cx = 0; cx < CW; cx++) {
            int hit = 0;
            for (int dy = 0; dy < q && !hit; dy++)
                for (int dx = 0; dx < q && !hit; dx++) {
                    int y = cy * q + dy, x = cx * q + dx;
                    if (y < H && x < W && score[(size_t)y * W + x] > thr32) hit = 1;
                }
            if (!hit) continue;
            int y = q > 1 ? cy * q + q / 2 : cy, x = q > 1 ? cx * q + q / 2 : cx;
            if (y >= H || x >= W) return -2;
            if (n >= cap) return -3;
            size_t pix = (size_t)y * W + x;
            orc_rbox_to_quad(x * scale, y * scale, geo5[pix] * scale, geo5[plane + pix] * scale,
                             geo5[2 * plane + pix] * scale, geo5[3 * plane + pix] * scale, geo5[4 * plane + pix],
                             out + 9 * n);
            out[9 * n + 8] = score[pix];
            n++;
        }
    return n;
}

/* ------------------------------------------------------------------------- */
/* lanms.py:156-207  sort by x0, sequential weighted merge, then NMS          */
/* boxes: (n,9) f32.  out: (cap,9) f32.  Optional debug outputs (may be NULL): */
/*   cl_polys (n*8 f64), cl_scores (n f64), n_clusters, keep_cluster (n i64)   */
/* returns m (rows written) or -1 if cap too small                             */
/* ------------------------------------------------------------------------- */
int64_t orc_lanms(const float *boxes, int64_t n, double thr, float *out, int64_t cap,
                  double *cl_polys_out, double *cl_scores_out, int64_t *n_clusters_out,
                  int64_t *keep_cluster_out)
{
    if (n_clusters_out)
        *n_clusters_out = 0;
    if (n <= 0)
        return 0;
    double *keys = (double *)malloc(sizeof(double) * (size_t)n);
    int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    for (int64_t i = 0; i < n; i++)
        keys[i] = (double)boxes[9 * i]; /* lanms.py:166-168: argsort(boxes[:,0]) */
    stable_argsort(keys, n, order);

    double *cp = (double *)malloc(sizeof(double) * 8 * (size_t)n); /* merged_polys  */
    double *cs = (double *)malloc(sizeof(double) * (size_t)n);     /* merged_scores */
    double *cw = (double *)malloc(sizeof(double) * (size_t)n);     /* weight_sums   */
    int64_t c = 0;
    for (int64_t r = 0; r < n; r++) {
        const float *row = boxes + 9 * order[r];
        double poly[8];
        for (int k = 0; k < 8; k++)
            poly[k] = (double)row[k];
        double score = (double)row[8];
        if (c > 0) {
            double *last = cp + 8 * (c - 1);
            if (orc_should_merge(poly, last, thr)) { /* lanms.py:180: subject = new box */
                double al[8];
                orc_normalize_polygon(last, poly, al);
                double tw = cw[c - 1] + score;
                for (int k = 0; k < 8; k++)
                    last[k] = (last[k] * cw[c - 1] + al[k] * score) / tw; /* lanms.py:183-185 */
                cw[c - 1] = tw;
                cs[c - 1] = cs[c - 1] > score ? cs[c - 1] : score; /* python max(a,b): b if b>a else a */
                continue;
            }
        }
        memcpy(cp + 8 * c, poly, sizeof(poly));
        cs[c] = score;
        cw[c] = score;
        c++;
    }
    int64_t *keep = (int64_t *)malloc(sizeof(int64_t) * (size_t)c);
    int64_t m = orc_standard_nms(cp, cs, c, thr, keep);
    if (cl_polys_out)
        memcpy(cl_polys_out, cp, sizeof(double) * 8 * (size_t)c);
    if (cl_scores_out)
        memcpy(cl_scores_out, cs, sizeof(double) * (size_t)c);
    if (n_clusters_out)
        *n_clusters_out = c;
    if (keep_cluster_out)
        memcpy(keep_cluster_out, keep, sizeof(int64_t) * (size_t)m);
    int64_t ret = m;
    if (m > cap)
        ret = -1;
    else
        for (int64_t i = 0; i < m; i++) {
            for (int k = 0; k < 8; k++)
                out[9 * i + k] = (float)cp[8 * keep[i] + k]; /* lanms.py:207 astype(f32) */
            out[9 * i + 8] = (float)cs[keep[i]];
        }
    free(keys);
    free(order);
    free(cp);
    free(cs);
    free(cw);
    free(keep);
    return ret;
}

/* ------------------------------------------------------------------------- */
/* utils.py:328-381  decode_quads_from_maps                                   */
/* score (H,W) f32; geo planar (8,H,W) f32 (infer.py:321 passes the transposed */
/* view of this layout).  Returns N, or -1 cap too small, or -2 when the       */
/* quantised centre falls outside the map (reference raises IndexError).       */
/* numpy>=2 promotion (NEP 50): `score > thr` compares in f32; `d*scale` is    */
/* f32*f32(scale); `x*scale` is f64; sum f64; row cast to f32 (utils.py:381).  */
/* ------------------------------------------------------------------------- */
int64_t orc_decode_quads(const float *score, const float *geo, int H, int W, double thr, double scale, int q,
                         float *out, int64_t cap)
{
    const float thr32 = (float)thr;
    const float sc32 = (float)scale;
    if (q < 1)
        q = 1;
    int CH = (H + q - 1) / q, CW = (W + q - 1) / q;
    int64_t n = 0;
    size_t plane = (size_t)H * (size_t)W;
    for (int cy = 0; cy < CH; cy++)
        for (int cx = 0; cx < CW; cx++) {
            int hit = 0;
            for (int dy = 0; dy < q && !hit; dy++)
                for (int dx = 0; dx < q; dx++) {
                    int y = cy * q + dy, x = cx * q + dx;
                    if (y < H && x < W && score[(size_t)y * W + x] > thr32) {
                        hit = 1;
                        break;
                    }
                }
            if (!hit)
                continue;
            /* utils.py:349-350: (c//q)*q + q//2 ; q==1 keeps the pixel itself */
            int y = q > 1 ? cy * q + q / 2 : cy;
            int x = q > 1 ? cx * q + q / 2 : cx;
            if (y >= H || x >= W)
                return -2;
            if (n >= cap)
                return -1;
            float *row = out + 9 * n;
            size_t pix = (size_t)y * W + x;
            for (int v = 0; v < 4; v++) {
                float dxm = geo[(size_t)(2 * v) * plane + pix];
                float dym = geo[(size_t)(2 * v + 1) * plane + pix];
                float tx = dxm * sc32, ty = dym * sc32;
                double vx = (double)x * scale + (double)tx;
                double vy = (double)y * scale + (double)ty;
                row[2 * v] = (float)vx;
                row[2 * v + 1] = (float)vy;
            }
            row[8] = score[pix];
            n++;
        }
    return n;
}

/* ------------------------------------------------------------------------- */
/* utils.py:384-422  expand_boxes, all float32                                 */
/* ------------------------------------------------------------------------- */
void orc_expand_boxes(const float *quads, int64_t n, double expand_w, double expand_h, float *out)
{
    if (n <= 0)
        return;
    if (expand_w == 0 && expand_h == 0) { /* utils.py:388-389 returns input unchanged */
        memcpy(out, quads, sizeof(float) * 9 * (size_t)n);
        return;
    }
    /* utils.py:415-416: scale_xy f32 array, then (scale_xy - 1.0) in f32 */
    const float sx = (float)(1.0 + expand_w) - 1.0f;
    const float sy = (float)(1.0 + expand_h) - 1.0f;
    const float eps = (float)1e-6;
    for (int64_t b = 0; b < n; b++) {
        const float *p = quads + 9 * b;
        float *o = out + 9 * b;
        /* utils.py:394-397  signed shoelace; np.sum over 4 f32 terms */
        float t[4];
        for (int i = 0; i < 4; i++) {
            int j = (i + 1) & 3;
            t[i] = p[2 * i] * p[2 * j + 1] - p[2 * j] * p[2 * i + 1];
        }
        float area = ((t[0] + t[1]) + t[2]) + t[3]; /* numpy add.reduce over a 4-long axis: left to right (checked) */
        float sign = area > 0 ? 1.0f : (area < 0 ? -1.0f : 1.0f);
        if (area != area)
            sign = area; /* np.sign(nan)=nan */
        for (int i = 0; i < 4; i++) {
            int ip = (i + 3) & 3, in = (i + 1) & 3;
            float e1x = p[2 * i] - p[2 * ip], e1y = p[2 * i + 1] - p[2 * ip + 1];
            float e2x = p[2 * in] - p[2 * i], e2y = p[2 * in + 1] - p[2 * i + 1];
            float l1 = sqrtf(e1x * e1x + e1y * e1y);
            float l2 = sqrtf(e2x * e2x + e2y * e2y);
            float n1x = sign * e1y / (l1 + eps), n1y = sign * (-e1x) / (l1 + eps);
            float n2x = sign * e2y / (l2 + eps), n2y = sign * (-e2x) / (l2 + eps);
            float ax = n1x + n2x, ay = n1y + n2y;
            float nn = sqrtf(ax * ax + ay * ay);
            if (nn > 0) {
                ax = ax / nn;
                ay = ay / nn;
            } else {
                ax = 0.0f;
                ay = 0.0f;
            }
            float off = l1 < l2 ? l1 : l2; /* np.minimum */
            if (l1 != l1 || l2 != l2)
                off = NAN;
            o[2 * i] = p[2 * i] + (sx * off) * ax;
            o[2 * i + 1] = p[2 * i + 1] + (sy * off) * ay;
        }
        o[8] = p[8];
    }
}

/* ------------------------------------------------------------------------- */
/* EAST box filters: infer.py:134-233 (float32 numpy + cv2.pointPolygonTest)   */
/* ------------------------------------------------------------------------- */

/* infer.py:134-147  in-place f32 multiply by f32(orig/target) */
void orc_scale_boxes(float *quads, int64_t n, int orig_h, int orig_w, int target)
{
    const float sx = (float)((double)orig_w / (double)target);
    const float sy = (float)((double)orig_h / (double)target);
    for (int64_t b = 0; b < n; b++)
        for (int v = 0; v < 4; v++) {
            quads[9 * b + 2 * v] *= sx;
            quads[9 * b + 2 * v + 1] *= sy;
        }
}

/* infer.py:174-183  0.5*|sum(x*roll(y,-1) - y*roll(x,-1))| in f32, sum left to right */
float orc_quad_area_f32(const float *p)
{
    float t[4];
    for (int i = 0; i < 4; i++) {
        int j = (i + 1) & 3;
        t[i] = p[2 * i] * p[2 * j + 1] - p[2 * i + 1] * p[2 * j];
    }
    float s = ((t[0] + t[1]) + t[2]) + t[3];
    return 0.5f * fabsf(s);
}

/* cv2.pointPolygonTest(contour f32, pt, measureDist=False) as called from infer.py:185-192.
 * Third-party arithmetic (OpenCV imgproc/src/geometry.cpp, opencv-python>=4.5,<5; 4.13 here):
 * crossing-number test with exact-boundary detection, float32 points, double cross product.
 * Returns +1 inside, 0 on the boundary, -1 outside. */
int orc_point_polygon_test(const float *cnt, int n, float px, float py)
{
    int counter = 0;
    float vx = cnt[2 * (n - 1)], vy = cnt[2 * (n - 1) + 1];
    for (int i = 0; i < n; i++) {
        float v0x = vx, v0y = vy;
        vx = cnt[2 * i];
        vy = cnt[2 * i + 1];
        if ((v0y <= py && vy <= py) || (v0y > py && vy > py) || (v0x < px && vx < px)) {
            if (py == vy && (px == vx || (py == v0y && ((v0x <= px && px <= vx) || (vx <= px && px <= v0x)))))
                return 0;
            continue;
        }
        double dist = (double)(py - v0y) * (double)(vx - v0x) - (double)(px - v0x) * (double)(vy - v0y);
        if (dist == 0)
            return 0;
        if (vy < v0y)
            dist = -dist;
        counter += dist > 0;
    }
    return counter % 2 == 0 ? -1 : 1;
}

/* infer.py:185-192 */
static int quad_inside(const float *inner, const float *outer)
{
    for (int v = 0; v < 4; v++)
        if (orc_point_polygon_test(outer, 4, inner[2 * v], inner[2 * v + 1]) < 0)
            return 0;
    return 1;
}

typedef struct {
    float key;
    int64_t idx;
} orc_kvf;
static int kvf_cmp(const void *a, const void *b)
{
    const orc_kvf *x = (const orc_kvf *)a, *y = (const orc_kvf *)b;
    int xn = isnan(x->key), yn = isnan(y->key);
    if (xn || yn) {
        if (xn != yn)
            return xn - yn;
    } else {
        if (x->key < y->key)
            return -1;
        if (x->key > y->key)
            return 1;
    }
    return (x->idx > y->idx) - (x->idx < y->idx);
}

/* infer.py:194-214  ascending-area scan (stable tie rule), keep[] out, returns kept count */
int64_t orc_remove_contained(const float *quads, int64_t n, unsigned char *keep)
{
    for (int64_t i = 0; i < n; i++)
        keep[i] = 1;
    if (n <= 1)
        return n;
    float *area = (float *)malloc(sizeof(float) * (size_t)n);
    orc_kvf *kv = (orc_kvf *)malloc(sizeof(orc_kvf) * (size_t)n);
    for (int64_t i = 0; i < n; i++) {
        area[i] = orc_quad_area_f32(quads + 9 * i);
        kv[i].key = area[i];
        kv[i].idx = i;
    }
    qsort(kv, (size_t)n, sizeof(orc_kvf), kvf_cmp);
    const float eps = (float)1e-6;
    for (int64_t r = 0; r < n; r++) {
        int64_t i = kv[r].idx;
        if (!keep[i])
            continue;
        for (int64_t j = 0; j < n; j++) {
            if (j == i || !keep[j])
                continue;
            if (area[j] + eps < area[i]) /* infer.py:208: f32 + f32(1e-6) < f32 */
                continue;
            if (quad_inside(quads + 9 * i, quads + 9 * j)) {
                keep[i] = 0;
                break;
            }
        }
    }
    int64_t k = 0;
    for (int64_t i = 0; i < n; i++)
        k += keep[i];
    free(area);
    free(kv);
    return k;
}

/* numpy pairwise summation of a contiguous f32 vector (numpy/_core/src/umath/loops_utils.h,
 * pairwise_sum; third-party arithmetic, numpy>=1.19,<2.3 pinned, 2.3.5 here) as np.add.reduce
 * applies it to a whole contiguous 1-D array (checked against np.sum in tests). */
static float np_pairwise_f32(const float *a, int64_t n)
{
    if (n < 8) {
        float r = -0.0f;
        for (int64_t i = 0; i < n; i++)
            r += a[i];
        return r;
    } else if (n <= 128) {
        float r[8];
        for (int k = 0; k < 8; k++)
            r[k] = a[k];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; k++)
                r[k] += a[i + k];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++)
            res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_f32(a, n2) + np_pairwise_f32(a + n2, n - n2);
    }
}
float orc_np_sum_f32(const float *a, int64_t n)
{
    if (n <= 0)
        return 0.0f;
    return np_pairwise_f32(a, n);
}

/* infer.py:216-233  drop area > mean + sigma*std when more than min_count boxes */
int64_t orc_remove_area_anomalies(const float *quads, int64_t n, int enabled, double sigma, int64_t min_count,
                                  unsigned char *keep)
{
    for (int64_t i = 0; i < n; i++)
        keep[i] = 1;
    if (!enabled || n == 0 || n <= min_count)
        return n;
    float *area = (float *)malloc(sizeof(float) * (size_t)n);
    float *dev = (float *)malloc(sizeof(float) * (size_t)n);
    for (int64_t i = 0; i < n; i++)
        area[i] = orc_quad_area_f32(quads + 9 * i);
    /* np.mean f32: add.reduce then true_divide by n, in f32; returned as python float (double) */
    float mean32 = orc_np_sum_f32(area, n) / (float)n;
    /* np.std f32 (_methods._var): arrmean = sum/n ; x = arr - arrmean ; x = x*x ; ret = sum(x)/n ; sqrt */
    for (int64_t i = 0; i < n; i++) {
        float d = area[i] - mean32;
        dev[i] = d * d;
    }
    float var32 = orc_np_sum_f32(dev, n) / (float)n;
    float std32 = sqrtf(var32);
    double mean = (double)mean32, std = (double)std32;
    int64_t k = n;
    if (std != 0.0) {
        double thr = mean + sigma * std; /* python floats */
        k = 0;
        /* areas (f32 array) <= python float: compared in f32 (NEP 50) */
        float thr32 = (float)thr;
        for (int64_t i = 0; i < n; i++) {
            keep[i] = area[i] <= thr32;
            k += keep[i];
        }
        if (k == 0) {
            for (int64_t i = 0; i < n; i++)
                keep[i] = 1;
            k = n;
        }
    }
    free(area);
    free(dev);
    return k;
}

/* infer.py:149-172  bbox rectangle TL,TR,BR,BL */
void orc_axis_align(float *quads, int64_t n)
{
    for (int64_t b = 0; b < n; b++) {
        float *p = quads + 9 * b;
        float x0 = p[0], x1 = p[0], y0 = p[1], y1 = p[1];
        for (int v = 1; v < 4; v++) {
            /* np.min/np.max propagate NaN; quads are finite on every tested path */
            x0 = p[2 * v] < x0 ? p[2 * v] : x0;
            x1 = p[2 * v] > x1 ? p[2 * v] : x1;
            y0 = p[2 * v + 1] < y0 ? p[2 * v + 1] : y0;
            y1 = p[2 * v + 1] > y1 ? p[2 * v + 1] : y1;
        }
        p[0] = x0; p[1] = y0; p[2] = x1; p[3] = y0;
        p[4] = x1; p[5] = y1; p[6] = x0; p[7] = y1;
    }
}

/* ------------------------------------------------------------------------- */
/* Pipeline crop rectangle: _pipeline.py:125-137, 204-221                      */
/* polygon (float) -> int32 truncation -> bbox -> min-size filter -> python    */
/* slice clamp.  rect = x1,y1,x2,y2 (exclusive), returns 1 if a crop is made.   */
/* ------------------------------------------------------------------------- */
static int py_slice_len(int start, int stop, int dim, int *s_out)
{
    /* python slice semantics for image[a:b] with a>=0 possible-negative b */
    if (start < 0) { start += dim; if (start < 0) start = 0; } else if (start > dim) start = dim;
    if (stop < 0) { stop += dim; if (stop < 0) stop = 0; } else if (stop > dim) stop = dim;
    *s_out = start;
    return stop > start ? stop - start : 0;
}
int orc_word_rect(const float *quad8, int img_h, int img_w, int min_text_size, int *rect)
{
    int xi[4], yi[4];
    for (int v = 0; v < 4; v++) {
        xi[v] = (int)quad8[2 * v]; /* np.array(list of float, dtype=int32): C truncation */
        yi[v] = (int)quad8[2 * v + 1];
    }
    int xmin = xi[0], xmax = xi[0], ymin = yi[0], ymax = yi[0];
    for (int v = 1; v < 4; v++) {
        if (xi[v] < xmin) xmin = xi[v];
        if (xi[v] > xmax) xmax = xi[v];
        if (yi[v] < ymin) ymin = yi[v];
        if (yi[v] > ymax) ymax = yi[v];
    }
    rect[0] = rect[1] = rect[2] = rect[3] = 0;
    if (!(xmax - xmin >= min_text_size && ymax - ymin >= min_text_size))
        return 0;
    int x1 = xmin > 0 ? xmin : 0, y1 = ymin > 0 ? ymin : 0;
    int x2 = xmax < img_w ? xmax : img_w, y2 = ymax < img_h ? ymax : img_h;
    int sx, sy;
    int w = py_slice_len(x1, x2, img_w, &sx);
    int h = py_slice_len(y1, y2, img_h, &sy);
    if (w <= 0 || h <= 0)
        return 0; /* region_image.size == 0 -> None */
    rect[0] = sx; rect[1] = sy; rect[2] = sx + w; rect[3] = sy + h;
    return 1;
}

/* ------------------------------------------------------------------------- */
/* TRBA preprocessing: transforms.py:62-120,185-193; __init__.py:264-288       */
/* cv2.resize is third-party arithmetic (OpenCV imgproc/src/resize.cpp,        */
/* opencv-python>=4.5,<5 pinned; 4.13.0 here).  Restated: INTER_LINEAR u8       */
/* (11-bit fixed point), INTER_AREA (integer-ratio fast path incl. the 2x2      */
/* (s+2)>>2 special case, and the float decimation-table path).                 */
/* ------------------------------------------------------------------------- */
static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline unsigned char sat_u8(int v) { return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

void orc_resize_linear_u8c3(const unsigned char *src, int sh, int sw, int64_t sstride, unsigned char *dst, int dh,
                            int dw)
{
    const int cn = 3;
    double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    int *xofs = (int *)malloc(sizeof(int) * (size_t)dw);
    short *ia = (short *)malloc(sizeof(short) * 2 * (size_t)dw);
    int xmax = dw;
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx + 1 >= sw) {
            if (dx < xmax) xmax = dx;
            if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        }
        xofs[dx] = sx;
        ia[2 * dx] = (short)cv_round_f((1.f - fx) * 2048.f);
        ia[2 * dx + 1] = (short)cv_round_f(fx * 2048.f);
    }
    int *row0 = (int *)malloc(sizeof(int) * (size_t)dw * cn);
    int *row1 = (int *)malloc(sizeof(int) * (size_t)dw * cn);
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        short b0 = (short)cv_round_f((1.f - fy) * 2048.f), b1 = (short)cv_round_f(fy * 2048.f);
        int y0 = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
        int y1 = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
        const unsigned char *S0 = src + (int64_t)y0 * sstride, *S1 = src + (int64_t)y1 * sstride;
        for (int dx = 0; dx < dw; dx++)
            for (int c = 0; c < cn; c++) {
                int sx = xofs[dx] * cn + c;
                if (dx < xmax) {
                    row0[dx * cn + c] = S0[sx] * ia[2 * dx] + S0[sx + cn] * ia[2 * dx + 1];
                    row1[dx * cn + c] = S1[sx] * ia[2 * dx] + S1[sx + cn] * ia[2 * dx + 1];
                } else {
                    row0[dx * cn + c] = S0[sx] * 2048;
                    row1[dx * cn + c] = S1[sx] * 2048;
                }
            }
        unsigned char *D = dst + (int64_t)dy * dw * cn;
        for (int x = 0; x < dw * cn; x++)
            D[x] = (unsigned char)((((b0 * (row0[x] >> 4)) >> 16) + ((b1 * (row1[x] >> 4)) >> 16) + 2) >> 2);
    }
    free(xofs); free(ia); free(row0); free(row1);
}

typedef struct { int si, di; float alpha; } orc_dec;
static int area_tab(int ssize, int dsize, double scale, orc_dec *tab)
{
    int k = 0;
    for (int dx = 0; dx < dsize; dx++) {
        double f1 = dx * scale, f2 = f1 + scale;
        double cell = scale < ssize - f1 ? scale : ssize - f1;
        int s1 = (int)ceil(f1), s2 = (int)floor(f2);
        if (s2 > ssize - 1) s2 = ssize - 1;
        if (s1 > s2) s1 = s2;
        if (s1 - f1 > 1e-3) { tab[k].di = dx; tab[k].si = s1 - 1; tab[k++].alpha = (float)((s1 - f1) / cell); }
        for (int s = s1; s < s2; s++) { tab[k].di = dx; tab[k].si = s; tab[k++].alpha = (float)(1.0 / cell); }
        if (f2 - s2 > 1e-3) {
            double a = f2 - s2 < 1. ? f2 - s2 : 1.;
            a = a < cell ? a : cell;
            tab[k].di = dx; tab[k].si = s2; tab[k++].alpha = (float)(a / cell);
        }
    }
    return k;
}

void orc_resize_area_u8c3(const unsigned char *src, int sh, int sw, int64_t sstride, unsigned char *dst, int dh,
                          int dw)
{
    const int cn = 3;
    double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    int isx = (int)lrint(scale_x), isy = (int)lrint(scale_y);
    int fast = fabs(scale_x - isx) < 2.220446049250313e-16 && fabs(scale_y - isy) < 2.220446049250313e-16;
    if (fast) {
        int area = isx * isy;
        float inv = 1.f / area;
        for (int dy = 0; dy < dh; dy++)
            for (int dx = 0; dx < dw; dx++)
                for (int c = 0; c < cn; c++) {
                    int sum = 0;
                    for (int yy = 0; yy < isy; yy++)
                        for (int xx = 0; xx < isx; xx++)
                            sum += src[(int64_t)(dy * isy + yy) * sstride + (dx * isx + xx) * cn + c];
                    unsigned char v;
                    if (isx == 2 && isy == 2)
                        v = (unsigned char)((sum + 2) >> 2);
                    else
                        v = sat_u8(cv_round_f((float)sum * inv));
                    dst[((int64_t)dy * dw + dx) * cn + c] = v;
                }
        return;
    }
    orc_dec *xt = (orc_dec *)malloc(sizeof(orc_dec) * 2 * (size_t)sw);
    orc_dec *yt = (orc_dec *)malloc(sizeof(orc_dec) * 2 * (size_t)sh);
    int nx = area_tab(sw, dw, scale_x, xt), ny = area_tab(sh, dh, scale_y, yt);
    float *buf = (float *)malloc(sizeof(float) * (size_t)dw * cn);
    float *sum = (float *)calloc((size_t)dw * cn, sizeof(float));
    int prev = ny > 0 ? yt[0].di : 0;
    for (int j = 0; j < ny; j++) {
        float beta = yt[j].alpha;
        int dy = yt[j].di;
        const unsigned char *S = src + (int64_t)yt[j].si * sstride;
        for (int x = 0; x < dw * cn; x++)
            buf[x] = 0.f;
        for (int k = 0; k < nx; k++) {
            int s = xt[k].si * cn, d = xt[k].di * cn;
            float a = xt[k].alpha;
            for (int c = 0; c < cn; c++)
                buf[d + c] = buf[d + c] + S[s + c] * a;
        }
        if (dy != prev) {
            unsigned char *D = dst + (int64_t)prev * dw * cn;
            for (int x = 0; x < dw * cn; x++) {
                D[x] = sat_u8(cv_round_f(sum[x]));
                sum[x] = beta * buf[x];
            }
            prev = dy;
        } else {
            for (int x = 0; x < dw * cn; x++)
                sum[x] += beta * buf[x];
        }
    }
    if (ny > 0) {
        unsigned char *D = dst + (int64_t)prev * dw * cn;
        for (int x = 0; x < dw * cn; x++)
            D[x] = sat_u8(cv_round_f(sum[x]));
    }
    free(xt); free(yt); free(buf); free(sum);
}

/* python round(): half to even on a double */
static int py_round_int(double v) { return (int)nearbyint(v); }

/* transforms.py:85-120 geometry of ResizeAndPadA: returns new_w,new_h,x0,y0 and interp (0 copy,1 linear,2 area) */
void orc_resize_plan(int h, int w, int ih, int iw, int *plan)
{
    double s1 = (double)ih / (double)(h > 1 ? h : 1), s2 = (double)iw / (double)(w > 1 ? w : 1);
    double scale = s1 < s2 ? s1 : s2;
    int nw = py_round_int(w * scale), nh = py_round_int(h * scale);
    if (nw < 1) nw = 1;
    if (nh < 1) nh = 1;
    int interp = (nh < h || nw < w) ? 2 : 1; /* transforms.py:80-83 */
    if (nh == h && nw == w)
        interp = 0; /* cv2.resize with dsize == ssize copies */
    int x0 = 0, y0 = (ih - nh) / 2; /* align left / center */
    if (x0 > iw - nw) x0 = iw - nw;
    if (x0 < 0) x0 = 0;
    if (y0 > ih - nh) y0 = ih - nh;
    if (y0 < 0) y0 = 0;
    plan[0] = nw; plan[1] = nh; plan[2] = x0; plan[3] = y0; plan[4] = interp;
}

/* One crop: page (H,W,3) u8, rect x1,y1,x2,y2 -> canvas u8 (ih,iw,3) [optional] and
 * normalised CHW f32 (3,ih,iw): albumentations Normalize(mean=.5,std=.5,max_pixel_value=255)
 * = (x - 127.5) * (1/127.5) in f32, then ToTensorV2 (HWC->CHW)  (transforms.py:185-193).
 * albumentations is not installed in the build container: this normalisation is restated from
 * its published formula and is "parity unpinned" (tolerance, not bit-exact, on the f32 batch). */
int orc_crop_resize_pad(const unsigned char *page, int H, int W, const int *rect, int ih, int iw,
                        unsigned char *canvas_u8, float *out_chw)
{
    (void)H;
    int x1 = rect[0], y1 = rect[1], w = rect[2] - rect[0], h = rect[3] - rect[1];
    if (w <= 0 || h <= 0)
        return -1;
    int plan[5];
    orc_resize_plan(h, w, ih, iw, plan);
    int nw = plan[0], nh = plan[1], x0 = plan[2], y0 = plan[3];
    if (nw > iw || nh > ih)
        return -2;
    unsigned char *tmp = (unsigned char *)malloc((size_t)nw * nh * 3);
    const unsigned char *src = page + ((int64_t)y1 * W + x1) * 3;
    int64_t stride = (int64_t)W * 3;
    if (plan[4] == 0)
        for (int y = 0; y < nh; y++)
            memcpy(tmp + (size_t)y * nw * 3, src + y * stride, (size_t)nw * 3);
    else if (plan[4] == 1)
        orc_resize_linear_u8c3(src, h, w, stride, tmp, nh, nw);
    else
        orc_resize_area_u8c3(src, h, w, stride, tmp, nh, nw);
    unsigned char *cv = canvas_u8 ? canvas_u8 : (unsigned char *)malloc((size_t)ih * iw * 3);
    memset(cv, 255, (size_t)ih * iw * 3);
    for (int y = 0; y < nh; y++)
        memcpy(cv + ((size_t)(y0 + y) * iw + x0) * 3, tmp + (size_t)y * nw * 3, (size_t)nw * 3);
    if (out_chw) {
        const float inv = 1.0f / 127.5f;
        for (int c = 0; c < 3; c++)
            for (int y = 0; y < ih; y++)
                for (int x = 0; x < iw; x++)
                    out_chw[((size_t)c * ih + y) * iw + x] = ((float)cv[((size_t)y * iw + x) * 3 + c] - 127.5f) * inv;
    }
    if (!canvas_u8)
        free(cv);
    free(tmp);
    return 0;
}
