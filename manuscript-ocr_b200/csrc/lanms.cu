// Locality-aware merge + polygon-IoU NMS on the device, float64, bit-exact with the reference.
// Replaces locality_aware_nms / standard_nms / polygon_iou (reference detectors/_east/lanms.py).
//
// What the reference computes (lanms.py:156-207), per page:
//   1. stable sort of the (N,9) candidates by x0 = boxes[:,0]                       (lanms.py:166)
//   2. a SEQUENTIAL scan: box i merges into the last cluster iff IoU(box_i, last_poly) > thr,
//      weighted vertex average after normalize_polygon, score = max         (lanms.py:174-192)
//   3. greedy NMS over the clusters in descending score, subject = kept box   (lanms.py:133-153)
//   4. rows = kept clusters in that order, cast to f32                          (lanms.py:204-207)
//
// How it is parallelised without changing a single decision:
//   (2) A non-merge resets the scan state to the box itself, so the state after any cluster start is
//       known without history.  hot[i] = IoU(box_i, box_{i-1}) > thr is computed for every i in
//       parallel; every hot position speculatively runs the sequential merge from start(box_{i-1})
//       until the first non-merge (lanms_runs_kernel); a per-page pass then walks the hot positions
//       in order and accepts a run only if its predecessor really is a cluster start (i.e. is not
//       covered by an earlier accepted run).  Accepted runs + untouched singletons are exactly the
//       reference's clusters, in creation order.  Worst case (one giant cluster) degrades to the
//       sequential cost on the device, never to a wrong answer.
//   (3) suppression edges hi->lo (IoU(hi,lo) > thr, hi earlier in stable score order) are found by
//       a sweep over clusters in their x0 order with an inflated-bounding-box prefilter, then the
//       greedy recurrence keep[i] = !any(keep[j], j->i) is resolved in rounds (a node is decided as
//       soon as all its predecessors are).  Identical to the sequential loop because the IoU
//       predicate is a pure function of the ordered pair.
//       The bbox prefilter is exact only for convex, positively oriented clip quads (for those the
//       Sutherland-Hodgman result is empty when the boxes are apart); any other quad ("irregular":
//       concave, self-intersecting, clockwise, degenerate, non-finite) is paired with every box of
//       its page, and thr < 0 (where IoU==0 pairs merge too) disables the prefilter altogether.
//
// Roofline: this stage is NOT HBM bound (36 B/box in, 36 B/kept out); it is bounded by fp64 ALU
// latency of the clip and by the serial dependencies above.
#include "ms_internal.cuh"

namespace {

// neighbour-pair capacity per candidate is LanmsBuffers::ef (ms_ctx::edge_factor; overflow -> MS_FLAG_EDGE_OVERFLOW,
// the host entry points then retry with a larger factor)
constexpr int kGrid = 64;  // uniform grid per page for the neighbour search (cells about one word box high)
constexpr int kCells = kGrid * kGrid;
// neighbour pair word: lo in bits 0..29, hi (higher NMS priority) in bits 30..59
constexpr uint64_t kPairTrue = 1ull << 62;  // IoU already evaluated and > thr
constexpr uint64_t kPairDone = 1ull << 63;  // nothing left to learn from this pair

struct LanmsBuffers {
    int32_t *page_off;   // n_pages+1 exclusive offsets of candidate counts (packed space)
    int32_t *n_total;    // == page_off[n_pages]
    uint32_t *keys, *keys_tmp;  // orderable(x0) per candidate, sorted per page
    uint32_t *vals, *vals_tmp;
    double *sq;          // sorted candidate polys, 8 doubles each
    float *ss;           // sorted candidate scores (f32 exact)
    int32_t *pos_page;   // page of each packed position
    uint8_t *mflag;      // 0 anchor / 1 merged / 2 accepted run head
    uint8_t *hot;
    int32_t *hot_list;
    int32_t *hot_count;
    int32_t *run_end;
    double *run_poly;
    double *run_score;
    // clusters, stored at page_off[p] + c
    double *cl_poly;
    double *cl_score;
    float4 *cl_bbox;
    uint8_t *cl_irr;
    int32_t *cl_orig;    // priority tie-break index (creation order)
    int32_t *cl_count;   // per page
    float *page_ext;     // per page, 8 floats: minx, miny, cells per unit x / y, max bbox width / height
    int32_t *cell_cnt;   // per page, kCells: regular clusters per grid cell
    int32_t *cell_off;   // per page, kCells + 1
    int32_t *cell_cur;   // per page, kCells: scatter cursors
    int32_t *cl_cell;    // per cluster: its home cell
    float4 *sb_bbox;     // clusters in cell order: inflated bbox ...
    int32_t *sb_id;      // ... and cluster index
    uint64_t *edges;
    int32_t *edge_count; // per page
    uint8_t *state;      // 0 undecided / 1 kept / 2 suppressed
    uint8_t *blocked;
    int32_t *kept_list;
    int32_t *irr_list;   // packed slots of the irregular clusters (all pages)
    int32_t *irr_count;
    int ef;              // neighbour-pair capacity per candidate
    int32_t *page_redo;  // per page: a cluster had more than kMaxHits neighbours -> exact two-pass rebuild
    int32_t *nb_cnt;     // two-pass rebuild: neighbours per cluster, then their exclusive offsets
    int32_t *und_flags;  // per page, 2 ints: "some box still undecided" of the current / previous round
    uint64_t *kept_key;  // per kept entry: descending-score sort key
    int32_t *acc_count;  // per page: accepted merge runs
    int32_t *ext_acc;    // per page, 8 ints: extent accumulators (ordered-int images of floats)
    int32_t *ext_done;   // per page: cluster-building CTAs that have contributed
};

// ---- helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ bool prio_before(double sa, int ia, double sb, int ib)
{
    // position in np.argsort(-scores, kind="stable"): larger score first, NaN last, ties by index
    bool an = sa != sa, bn = sb != sb;
    if (an || bn) {
        if (an != bn) return bn;
        return ia < ib;
    }
    if (sa > sb) return true;
    if (sa < sb) return false;
    return ia < ib;
}

__device__ __forceinline__ void load_quad(const double *__restrict__ src, double *dst)
{
    const double2 *s2 = reinterpret_cast<const double2 *>(src);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        double2 t = s2[k];
        dst[2 * k] = t.x;
        dst[2 * k + 1] = t.y;
    }
}

// convex + positively oriented + finite, with a relative margin; also returns the inflated bbox
__device__ __forceinline__ bool quad_regular_bbox(const double *p, float4 &bb)
{
    double minx = p[0], maxx = p[0], miny = p[1], maxy = p[1];
    bool finite = true;
#pragma unroll
    for (int v = 0; v < 4; v++) {
        double x = p[2 * v], y = p[2 * v + 1];
        finite = finite && (fabs(x) < 1e30) && (fabs(y) < 1e30);  // false for NaN/Inf too
        minx = fmin(minx, x);
        maxx = fmax(maxx, x);
        miny = fmin(miny, y);
        maxy = fmax(maxy, y);
    }
    bool convex = true;
#pragma unroll
    for (int v = 0; v < 4; v++) {
        int a = v, b = (v + 1) & 3, c = (v + 2) & 3;
        double e1x = p[2 * b] - p[2 * a], e1y = p[2 * b + 1] - p[2 * a + 1];
        double e2x = p[2 * c] - p[2 * b], e2y = p[2 * c + 1] - p[2 * b + 1];
        double cr = e1x * e2y - e1y * e2x;
        double l1 = e1x * e1x + e1y * e1y, l2 = e2x * e2x + e2y * e2y;
        // sin(angle) > 1e-3 and both edges longer than 1e-3 px
        convex = convex && (cr > 0) && (cr * cr > 1e-6 * l1 * l2) && (l1 > 1e-6) && (l2 > 1e-6);
    }
    if (!(finite && convex)) {
        bb = make_float4(-INFINITY, -INFINITY, INFINITY, INFINITY);
        return false;
    }
    double ext = fmax(fmax(fabs(minx), fabs(maxx)), fmax(fabs(miny), fabs(maxy)));
    double m = 1e-3 + 1e-6 * ext;
    bb.x = __double2float_rd(minx - m);
    bb.y = __double2float_rd(miny - m);
    bb.z = __double2float_ru(maxx + m);
    bb.w = __double2float_ru(maxy + m);
    return true;
}

// ---- 0. page offsets ---------------------------------------------------------------------------------
__global__ void lanms_offsets_kernel(const int32_t *__restrict__ counts, int n_pages, int32_t *page_off,
                                     int32_t *n_total, int32_t *hot_count, int32_t *edge_count, int32_t *irr_count,
                                     int32_t *page_redo)
{
    ms_pdl_wait();
    // one thread: n_pages is small (<= a few thousand)
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int run = 0;
        for (int p = 0; p < n_pages; p++) {
            page_off[p] = run;
            run += counts[p];
            edge_count[p] = 0;
            page_redo[p] = 0;
        }
        page_off[n_pages] = run;
        *n_total = run;
        *hot_count = 0;
        *irr_count = 0;
    }
}

// ---- 1. sort keys (orderable x0; pages are sorted separately), values = strided source row ---------------------------------
__global__ void lanms_keys_kernel(const float *__restrict__ quads, const int32_t *__restrict__ counts,
                                  const int32_t *__restrict__ page_off, int n_pages, int cap,
                                  uint32_t *__restrict__ keys, uint32_t *__restrict__ vals,
                                  int32_t *__restrict__ pos_page)
{
    ms_pdl_wait();
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int p = (int)(g / cap), i = (int)(g % cap);
    if (p >= n_pages || i >= counts[p]) return;
    size_t row = (size_t)p * cap + i;
    int dst = page_off[p] + i;
    keys[dst] = ms_orderable_f32(quads[row * 9]);
    vals[dst] = (uint32_t)row;
    pos_page[dst] = p;
}

// ---- 2. gather into sorted order (f64) + hot bits -----------------------------------------------------
__global__ void __launch_bounds__(128) lanms_gather_hot_kernel(const float *__restrict__ quads,
                                                               const uint32_t *__restrict__ vals,
                                                               const int32_t *__restrict__ page_off,
                                                               const int32_t *__restrict__ n_total, double thr,
                                                               LanmsBuffers B)
{
    ms_pdl_wait();
    const int n = *n_total;
    double buf[4 * MS_MAXV];
    __shared__ int s_q[128 / 32][64];  // per warp: positions whose IoU with the predecessor has to be evaluated
    const int lane = threadIdx.x & 31;
    int *q = s_q[threadIdx.x >> 5];
    int qn = 0;  // warp-uniform
    // lanms.py:180 IoU(subject = box s, clip = box s - 1) > thr.  In x0 order most neighbours lie in different text
    // rows: when the clip quad is regular (convex, positively oriented -- Sutherland-Hodgman then returns the true
    // intersection) and the subject's corners all lie beyond one side of its inflated bounding box, the intersection is
    // empty and the IoU is exactly 0, which is not > thr for thr >= 0.  The few positions that need the float64 clip
    // are queued per warp and evaluated 32 at a time, so that a warp never clips for one lane's sake.
    auto evaluate = [&](int s) {
        const float *row = quads + (size_t)vals[s] * 9, *prow = quads + (size_t)vals[s - 1] * 9;
        double me[8], pv[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            me[k] = (double)row[k];
            pv[k] = (double)prow[k];
        }
        if (ms_quad_iou(me, pv, buf) > thr) {
            B.hot[s] = 1;
            B.hot_list[atomicAdd(B.hot_count, 1)] = s;
        }
    };
    const int stride = gridDim.x * blockDim.x;
    for (int s0 = blockIdx.x * blockDim.x; s0 < n; s0 += stride) {  // warp-uniform trip count
        const int s = s0 + threadIdx.x;
        bool need = false;
        // the row of position s, and its predecessor's from the lane below (lane 0 reads it itself)
        float r[8], pr[8];
        const float *row = quads + (size_t)(s < n ? vals[s] : 0) * 9;
#pragma unroll
        for (int k = 0; k < 8; k++) r[k] = s < n ? row[k] : 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) pr[k] = __shfl_up_sync(0xffffffffu, r[k], 1);
        if (lane == 0 && s > 0 && s < n) {
            const float *prow = quads + (size_t)vals[s - 1] * 9;
#pragma unroll
            for (int k = 0; k < 8; k++) pr[k] = prow[k];
        }
        if (s < n) {
            const int page = B.pos_page[s];
            double me[8];
#pragma unroll
            for (int k = 0; k < 8; k++) me[k] = (double)r[k];
            double2 *dst = reinterpret_cast<double2 *>(B.sq + (size_t)s * 8);
#pragma unroll
            for (int k = 0; k < 4; k++) dst[k] = make_double2(me[2 * k], me[2 * k + 1]);
            B.ss[s] = row[8];
            B.hot[s] = 0;
            if (s > page_off[page]) {
                double pv[8];
#pragma unroll
                for (int k = 0; k < 8; k++) pv[k] = (double)pr[k];
                bool disjoint = false;
                float4 pb;
                if (thr >= 0.0 && quad_regular_bbox(pv, pb)) {
                    const double mnx = fmin(fmin(me[0], me[2]), fmin(me[4], me[6]));
                    const double mxx = fmax(fmax(me[0], me[2]), fmax(me[4], me[6]));
                    const double mny = fmin(fmin(me[1], me[3]), fmin(me[5], me[7]));
                    const double mxy = fmax(fmax(me[1], me[3]), fmax(me[5], me[7]));
                    disjoint = mnx > (double)pb.z || mxx < (double)pb.x || mny > (double)pb.w || mxy < (double)pb.y;
                }
                need = !disjoint;
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, need);
        if (need) q[qn + __popc(m & ((1u << lane) - 1u))] = s;
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) {
            const int s2 = q[qn - 32 + lane];
            __syncwarp();
            evaluate(s2);
            qn -= 32;
            __syncwarp();
        }
    }
    if (lane < qn) evaluate(q[lane]);
}

// ---- 3. speculative runs from every hot position ---------------------------------------------------------
__global__ void __launch_bounds__(128) lanms_runs_kernel(const int32_t *__restrict__ page_off, double thr,
                                                         LanmsBuffers B)
{
    ms_pdl_wait();
    const int nh = *B.hot_count;
    double buf[4 * MS_MAXV];
    // Runs differ in length (most end at their first IoU test, a few swallow a whole word) and one step of a run is a
    // float64 polygon clip: with a run per loop trip the lanes of a warp waited for its longest run (8 of 32 lanes active
    // on average).  Here a lane that finishes its run takes the next one of its own strided list at once, so the lanes
    // of a warp are at different steps of different runs and the clip runs nearly full; the steps of a run and their
    // order are unchanged.
    const int stride = gridDim.x * blockDim.x;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = false;
    int h = 0, i = 0, page_end = 0;
    double L[8], nb[8], al[8];
    double wsum = 0.0, score = 0.0;
    while (true) {
        if (!active && j < nh) {
            h = B.hot_list[j];
            j += stride;
            page_end = page_off[B.pos_page[h] + 1];
            load_quad(B.sq + (size_t)(h - 1) * 8, L);
            wsum = (double)B.ss[h - 1];
            score = wsum;
            i = h;
            load_quad(B.sq + (size_t)i * 8, nb);
            active = true;
        }
        if (!__any_sync(0xffffffffu, active)) break;
        if (active) {
            // merge box i into the running cluster (lanms.py:181-188)
            double sc = (double)B.ss[i];
            ms_align_vertices(L, nb, al);
            double tw = wsum + sc;
#pragma unroll
            for (int k = 0; k < 8; k++) L[k] = (L[k] * wsum + al[k] * sc) / tw;
            wsum = tw;
            score = (sc > score) ? sc : score;  // python max(old, new)
            i++;
            bool fin = i >= page_end;
            if (!fin) {
                load_quad(B.sq + (size_t)i * 8, nb);
                fin = !(ms_quad_iou(nb, L, buf) > thr);
            }
            if (fin) {
                B.run_end[h] = i;
                double2 *dst = reinterpret_cast<double2 *>(B.run_poly + (size_t)h * 8);
#pragma unroll
                for (int k = 0; k < 4; k++) dst[k] = make_double2(L[2 * k], L[2 * k + 1]);
                B.run_score[h] = score;
                active = false;
            }
        }
    }
}

// ---- 4. accept runs in order, build clusters (one CTA per page) --------------------------------------------
constexpr int kResolveThreads = 1024;

__device__ __forceinline__ int block_excl_scan_1024(int v, int *s_warp, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    __syncthreads();  // protect s_warp reuse
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
        int winc = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, off);
            if (lane >= off) winc += t;
        }
        s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    total = s_warp[32];
    return s_warp[warp] + inc - v;
}

// 4a. one CTA per page: the page's hot positions in order, the accepted runs among them (a run is accepted iff its
//     head lies beyond the end of the last accepted run -- a greedy chain, marked by pointer doubling in shared memory),
//     and for every accepted run the number of merged positions before it.  Clusters = positions outside the runs.
constexpr int kAcceptBatch = 2048;  // hot positions walked per round out of shared memory

__global__ void __launch_bounds__(kResolveThreads) lanms_accept_kernel(const int32_t *__restrict__ page_off, LanmsBuffers B)
{
    ms_pdl_wait();
    const int page = blockIdx.x;
    const int p0 = page_off[page], p1 = page_off[page + 1];
    __shared__ int s_warp[33];
    __shared__ int s_h[kAcceptBatch], s_e[kAcceptBatch], s_sel[kAcceptBatch], s_ptr[kAcceptBatch];
    __shared__ unsigned char s_mark[kAcceptBatch];
    __shared__ int s_klast, s_nacc, s_merged, s_nv, s_root;
    int32_t *hot_sorted = B.cl_cell + p0;  // scratch until the neighbour grid is built
    int32_t *acc_h = B.kept_list + p0, *acc_e = B.nb_cnt + p0, *acc_pm = B.sb_id + p0;
    if (threadIdx.x == 0) {
        s_klast = p0 - 1;
        s_nacc = 0;
        s_merged = 0;
        B.ext_done[page] = 0;
        int32_t *acc = B.ext_acc + (size_t)page * 8;
        acc[0] = acc[1] = INT_MAX;  // min x, min y (ordered-int images of floats)
        acc[2] = acc[3] = INT_MIN;  // max x, max y
        acc[4] = acc[5] = 0;        // max width, max height (non-negative floats order like their bits)
    }
    for (int i = threadIdx.x; i < kCells; i += kResolveThreads) B.cell_cnt[(size_t)page * kCells + i] = 0;
    // hot positions of the page, ascending: 8 consecutive positions per thread and round
    int nh = 0;
    for (int base = p0; base < p1; base += kResolveThreads * 8) {
        const int s0 = base + threadIdx.x * 8;
        uint32_t bits = 0;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (s0 + k < p1 && B.hot[s0 + k]) bits |= 1u << k;
        int total;
        int pos = nh + block_excl_scan_1024(__popc(bits), s_warp, total);
        while (bits) {
            const int k = __ffs(bits) - 1;
            bits &= bits - 1;
            hot_sorted[pos++] = s0 + k;
        }
        nh += total;
    }
    __syncthreads();
    int nacc = 0, merged = 0;  // CTA-uniform: accepted runs so far, merged positions so far
    for (int b0 = 0; b0 < nh; b0 += kAcceptBatch) {
        const int nb = min(kAcceptBatch, nh - b0);
        for (int j = threadIdx.x; j < nb; j += kResolveThreads) {
            const int h = hot_sorted[b0 + j];
            s_h[j] = h;
            s_e[j] = B.run_end[h];
        }
        __syncthreads();
        // Which runs are accepted is a greedy chain: run j is accepted iff its head lies beyond the end of the last
        // accepted run, i.e. the accepted runs are root, next(root), next(next(root)), ... with
        // next(j) = first j' with h[j'] > e[j] (the heads are ascending: a binary search), root = first j with
        // h[j] > klast.  A one-thread walk of that chain was 50 us per launch (one dependent step per hot position);
        // here every position finds its successor at once and the chain is marked by pointer doubling: in round r
        // every marked position marks its 2^r-th successor, then the successor table is squared.  ceil(log2(nb))
        // rounds, each a few shared-memory accesses per thread.  (Before round r every position within 2^r steps of the
        // root is marked; round r marks those within 2^(r+1).)
        for (int j = threadIdx.x; j < nb; j += kResolveThreads) {
            const int key = s_e[j];  // first j' with s_h[j'] > key; j' > j because e >= h
            int lo = j + 1, hi = nb;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_h[mid] > key)
                    hi = mid;
                else
                    lo = mid + 1;
            }
            s_ptr[j] = lo;
            s_mark[j] = 0;
        }
        if (threadIdx.x == 0) {
            const int klast = s_klast;
            int lo = 0, hi = nb;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_h[mid] > klast)
                    hi = mid;
                else
                    lo = mid + 1;
            }
            s_root = lo;
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_root < nb) s_mark[s_root] = 1;
        __syncthreads();
        for (int span = 1; span < nb; span <<= 1) {
            // read phase (marks and successors as they stand), barrier, write phase: no location is read and written
            // in the same phase; several threads may store the same 1 into a mark
            int nxt[kAcceptBatch / kResolveThreads], tgt[kAcceptBatch / kResolveThreads];
#pragma unroll
            for (int u = 0; u < kAcceptBatch / kResolveThreads; u++) {
                const int j = threadIdx.x + u * kResolveThreads;
                nxt[u] = nb;
                tgt[u] = -1;
                if (j < nb) {
                    const int p = s_ptr[j];
                    if (p < nb) {
                        if (s_mark[j]) tgt[u] = p;
                        nxt[u] = s_ptr[p];
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < kAcceptBatch / kResolveThreads; u++) {
                const int j = threadIdx.x + u * kResolveThreads;
                if (tgt[u] >= 0) s_mark[tgt[u]] = 1;
                if (j < nb) s_ptr[j] = nxt[u];
            }
            __syncthreads();
        }
        // the marked positions in order -> s_sel; the last one's end is where the next batch continues
        {
            int nv_run = 0;
            for (int t0 = 0; t0 < nb; t0 += kResolveThreads) {
                const int j = t0 + threadIdx.x;
                const int m = (j < nb && s_mark[j]) ? 1 : 0;
                int total;
                const int pos = block_excl_scan_1024(m, s_warp, total);
                if (m) s_sel[nv_run + pos] = j;
                nv_run += total;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                s_nv = nv_run;
                if (nv_run > 0) s_klast = s_e[s_sel[nv_run - 1]];
            }
        }
        __syncthreads();
        const int nv = s_nv;
        // accepted runs of this batch, in order: head, end, merged positions before the run (exclusive running sum)
        for (int t0 = 0; t0 < nv; t0 += kResolveThreads) {
            const int t = t0 + threadIdx.x;
            int h = 0, e = 0;
            if (t < nv) {
                const int j = s_sel[t];
                h = s_h[j];
                e = s_e[j];
            }
            int total;
            const int before = block_excl_scan_1024(e - h, s_warp, total);
            if (t < nv) {
                acc_h[nacc + t] = h;
                acc_e[nacc + t] = e;
                acc_pm[nacc + t] = merged + before;
            }
            merged += total;
        }
        nacc += nv;
        __syncthreads();  // s_h / s_e / s_sel are refilled by the next batch
    }
    if (threadIdx.x == 0) {
        s_nacc = nacc;
        s_merged = merged;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        B.acc_count[page] = s_nacc;
        B.cl_count[page] = (p1 - p0) - s_merged;
    }
}

__device__ __forceinline__ int float_ordered(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// 4b. thread per position, any number of CTAs per page: a position outside every accepted run anchors a cluster --
//     the run that starts right after it if there is one, else the box itself; its cluster index is its position
//     minus the merged positions before it (binary search in the page's accepted runs).
constexpr int kBuildThreads = 256;
constexpr int kBuildRuns = 1024;  // accepted runs searched out of shared memory

__global__ void __launch_bounds__(kBuildThreads) lanms_build_clusters_kernel(const int32_t *__restrict__ page_off,
                                                                             int all_irregular, LanmsBuffers B)
{
    ms_pdl_wait();
    const int page = blockIdx.y;
    const int p0 = page_off[page], p1 = page_off[page + 1];
    __shared__ int s_h[kBuildRuns], s_e[kBuildRuns], s_pm[kBuildRuns];
    __shared__ float s_red[kBuildThreads / 32][6];
    const int nacc = B.acc_count[page];
    const int32_t *acc_h = B.kept_list + p0, *acc_e = B.nb_cnt + p0, *acc_pm = B.sb_id + p0;
    const bool in_smem = nacc <= kBuildRuns;
    if (in_smem) {
        for (int j = threadIdx.x; j < nacc; j += kBuildThreads) {
            s_h[j] = acc_h[j];
            s_e[j] = acc_e[j];
            s_pm[j] = acc_pm[j];
        }
    }
    __syncthreads();
    const int *gh = in_smem ? s_h : acc_h, *ge = in_smem ? s_e : acc_e, *gpm = in_smem ? s_pm : acc_pm;
    // last accepted run whose head is <= s (-1: none)
    auto run_at = [&](int s) {
        int lo = 0, hi = nacc;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (gh[mid] <= s)
                lo = mid + 1;
            else
                hi = mid;
        }
        return lo - 1;
    };
    float e_minx = INFINITY, e_miny = INFINITY, e_maxx = -INFINITY, e_maxy = -INFINITY, e_w = 0.f, e_h = 0.f;
    for (int s = p0 + blockIdx.x * kBuildThreads + threadIdx.x; s < p1; s += gridDim.x * kBuildThreads) {
        const int r = run_at(s);
        const bool inside = r >= 0 && s < ge[r];
        if (!inside) {
            const int merged_before = r >= 0 ? gpm[r] + (ge[r] - gh[r]) : 0;
            const int c = (s - p0) - merged_before;
            const size_t slot = (size_t)p0 + c;
            const bool has_run = r + 1 < nacc && gh[r + 1] == s + 1;
            double poly[8];
            double score;
            if (has_run) {
                load_quad(B.run_poly + (size_t)(s + 1) * 8, poly);
                score = B.run_score[s + 1];
            } else {
                load_quad(B.sq + (size_t)s * 8, poly);
                score = (double)B.ss[s];
            }
            double2 *dst = reinterpret_cast<double2 *>(B.cl_poly + slot * 8);
#pragma unroll
            for (int k = 0; k < 4; k++) dst[k] = make_double2(poly[2 * k], poly[2 * k + 1]);
            B.cl_score[slot] = score;
            float4 bb;
            bool reg = quad_regular_bbox(poly, bb) && !all_irregular;
            if (!reg) bb = make_float4(-INFINITY, -INFINITY, INFINITY, INFINITY);
            B.cl_bbox[slot] = bb;
            B.cl_irr[slot] = reg ? 0 : 1;
            if (!reg) B.irr_list[atomicAdd(B.irr_count, 1)] = (int)slot;
            B.cl_orig[slot] = c;
            B.state[slot] = 0;
            if (reg) {
                e_minx = fminf(e_minx, bb.x);
                e_miny = fminf(e_miny, bb.y);
                e_maxx = fmaxf(e_maxx, bb.z);
                e_maxy = fmaxf(e_maxy, bb.w);
                e_w = fmaxf(e_w, bb.z - bb.x);
                e_h = fmaxf(e_h, bb.w - bb.y);
            }
        }
    }
    // page extents of the regular clusters -> geometry of the neighbour-search grid: CTA reduction, one set of
    // atomics per CTA, and the last CTA of the page turns the accumulators into the grid geometry
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        e_minx = fminf(e_minx, __shfl_xor_sync(0xffffffffu, e_minx, off));
        e_miny = fminf(e_miny, __shfl_xor_sync(0xffffffffu, e_miny, off));
        e_maxx = fmaxf(e_maxx, __shfl_xor_sync(0xffffffffu, e_maxx, off));
        e_maxy = fmaxf(e_maxy, __shfl_xor_sync(0xffffffffu, e_maxy, off));
        e_w = fmaxf(e_w, __shfl_xor_sync(0xffffffffu, e_w, off));
        e_h = fmaxf(e_h, __shfl_xor_sync(0xffffffffu, e_h, off));
    }
    if ((threadIdx.x & 31) == 0) {
        float *r = s_red[threadIdx.x >> 5];
        r[0] = e_minx; r[1] = e_miny; r[2] = e_maxx; r[3] = e_maxy; r[4] = e_w; r[5] = e_h;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float minx = INFINITY, miny = INFINITY, maxx = -INFINITY, maxy = -INFINITY, mw = 0.f, mh = 0.f;
        for (int w = 0; w < kBuildThreads / 32; w++) {
            minx = fminf(minx, s_red[w][0]);
            miny = fminf(miny, s_red[w][1]);
            maxx = fmaxf(maxx, s_red[w][2]);
            maxy = fmaxf(maxy, s_red[w][3]);
            mw = fmaxf(mw, s_red[w][4]);
            mh = fmaxf(mh, s_red[w][5]);
        }
        int32_t *acc = B.ext_acc + (size_t)page * 8;
        if (maxx >= minx) {
            atomicMin(&acc[0], float_ordered(minx));
            atomicMin(&acc[1], float_ordered(miny));
            atomicMax(&acc[2], float_ordered(maxx));
            atomicMax(&acc[3], float_ordered(maxy));
            atomicMax(&acc[4], __float_as_int(mw));
            atomicMax(&acc[5], __float_as_int(mh));
        }
        __threadfence();
        if (atomicAdd(&B.ext_done[page], 1) == (int)gridDim.x - 1) {  // every CTA of the page passes here
            __threadfence();
            const volatile int32_t *va = acc;
            const bool any = va[2] != INT_MIN;
            const float fminx = ordered_float(va[0]), fminy = ordered_float(va[1]);
            const float fmaxx = ordered_float(va[2]), fmaxy = ordered_float(va[3]);
            const float sx = any ? fmaxf(fmaxx - fminx, 1e-3f) : 1.f, sy = any ? fmaxf(fmaxy - fminy, 1e-3f) : 1.f;
            float *ext = B.page_ext + (size_t)page * 8;
            ext[0] = any ? fminx : 0.f;
            ext[1] = any ? fminy : 0.f;
            ext[2] = (float)kGrid / sx;
            ext[3] = (float)kGrid / sy;
            ext[4] = __int_as_float(va[4]);
            ext[5] = __int_as_float(va[5]);
        }
    }
}

// ---- 5. neighbour pairs ---------------------------------------------------------------------------------------
// Only pairs whose (inflated) bounding boxes overlap can have IoU > thr >= 0 when both quads are regular (convex,
// positively oriented): those pairs are found with a uniform grid per page and stored UNEVALUATED -- the fp64 clip
// runs lazily in the resolve step, only for pairs whose higher-priority box turned out to be kept.  Irregular quads
// (and every quad when thr < 0) are paired with every other box of their page and evaluated eagerly.
constexpr int kPairThreads = 128;
constexpr int kPairWarps = kPairThreads / 32;
constexpr int kQueue = 96;  // per-warp pending pairs (flush at >= 32)

__device__ __forceinline__ int cell_coord(float v, float origin, float scale)
{
    // monotone in v: the same function maps box corners and search bounds
    float f = (v - origin) * scale;
    int c = f > 0.f ? (f < (float)(kGrid - 1) ? (int)f : kGrid - 1) : 0;
    return c;
}

__device__ __forceinline__ void push_pair(int hi, int lo, uint64_t flag, int p0, int page, const LanmsBuffers &B,
                                          int edge_cap, int32_t *flags)
{
    int e = atomicAdd(B.edge_count + page, 1);
    if (e < edge_cap)
        B.edges[(size_t)p0 * B.ef + e] = flag | ((uint64_t)(uint32_t)hi << 30) | (uint64_t)(uint32_t)lo;
    else
        atomicOr(flags + page, MS_FLAG_EDGE_OVERFLOW);
}

__device__ __forceinline__ void eval_pair(int a, int b, int p0, int page, double thr, const LanmsBuffers &B,
                                          int edge_cap, int32_t *flags, double *buf)
{
    // a, b: local cluster indices of `page`
    size_t sa = (size_t)p0 + a, sb = (size_t)p0 + b;
    double qa[8], qb[8];
    load_quad(B.cl_poly + sa * 8, qa);
    load_quad(B.cl_poly + sb * 8, qb);
    bool a_first = prio_before(B.cl_score[sa], B.cl_orig[sa], B.cl_score[sb], B.cl_orig[sb]);
    // lanms.py:149 should_merge(polys[idx] (kept, earlier), polys[idx_j] (later))
    double iou = a_first ? ms_quad_iou(qa, qb, buf) : ms_quad_iou(qb, qa, buf);
    if (iou > thr) push_pair(a_first ? a : b, a_first ? b : a, kPairTrue, p0, page, B, edge_cap, flags);
}

// Warp per listed cluster, paired with every other box of its page, IoU evaluated eagerly.  irregular_only != 0:
// the items are the irregular clusters (pairs of two irregular boxes are produced once, by the one with the
// smaller index); otherwise every slot is an item (standard_nms, where all boxes are flagged irregular).
__global__ void __launch_bounds__(kPairThreads) lanms_pairs_kernel(const int32_t *__restrict__ page_off,
                                                                   const int32_t *__restrict__ n_total, double thr,
                                                                   LanmsBuffers B, int32_t *flags, int irregular_only)
{
    ms_pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = *n_total;
    __shared__ int2 s_q[kPairWarps][kQueue];
    __shared__ int s_qpage[kPairWarps][kQueue];
    int2 *q = s_q[warp];
    int *qpage = s_qpage[warp];
    int qn = 0;  // warp-uniform
    double buf[4 * MS_MAXV];
    const int wid = blockIdx.x * kPairWarps + warp, nw = gridDim.x * kPairWarps;
    const int n_items = irregular_only ? *B.irr_count : n;

    for (int item = wid; item < n_items; item += nw) {
        const int slot = irregular_only ? B.irr_list[item] : item;
        const int page = B.pos_page[slot];
        const int p0 = page_off[page];
        const int a = slot - p0;
        const int C = B.cl_count[page];
        if (a >= C || B.cl_irr[slot] == 0) continue;
        for (int bb = 0; bb < C; bb += 32) {
            int b = bb + lane;
            bool hit = b < C && b != a && !(B.cl_irr[(size_t)p0 + b] != 0 && b < a);
            uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                int pos = qn + __popc(m & ((1u << lane) - 1u));
                q[pos] = make_int2(a, b);
                qpage[pos] = page;
            }
            qn += __popc(m);
            __syncwarp();
            while (qn >= 32) {
                int idx = qn - 32 + lane;
                int2 pr = q[idx];
                int pg = qpage[idx];
                __syncwarp();
                eval_pair(pr.x, pr.y, page_off[pg], pg, thr, B,
                          (page_off[pg + 1] - page_off[pg]) * B.ef, flags, buf);
                qn -= 32;
                __syncwarp();
            }
        }
    }
    if (lane < qn) {
        int2 pr = q[lane];
        int pg = qpage[lane];
        eval_pair(pr.x, pr.y, page_off[pg], pg, thr, B, (page_off[pg + 1] - page_off[pg]) * B.ef, flags,
                  buf);
    }
}

// grid binning of the regular clusters by the min corner of their inflated bbox
__global__ void __launch_bounds__(256) nms_bin_count_kernel(const int32_t *__restrict__ page_off,
                                                            const int32_t *__restrict__ n_total, LanmsBuffers B)
{
    ms_pdl_wait();
    const int n = *n_total;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
        const int page = B.pos_page[slot];
        const int c = slot - page_off[page];
        if (c >= B.cl_count[page] || B.cl_irr[slot]) continue;
        const float *ext = B.page_ext + (size_t)page * 8;
        const float4 bb = B.cl_bbox[slot];
        const int cell = cell_coord(bb.y, ext[1], ext[3]) * kGrid + cell_coord(bb.x, ext[0], ext[2]);
        B.cl_cell[slot] = cell;
        atomicAdd(B.cell_cnt + (size_t)page * kCells + cell, 1);
    }
}

__global__ void __launch_bounds__(1024) nms_bin_scan_kernel(LanmsBuffers B)
{
    ms_pdl_wait();
    // exclusive scan of the page's kCells cell counts: a thread owns kCells / 1024 consecutive cells
    static_assert(kCells % 1024 == 0, "cells per thread");
    constexpr int kPer = kCells / 1024;
    const int page = blockIdx.x;
    __shared__ int s_warp[33];
    int v[kPer], sum = 0;
#pragma unroll
    for (int i = 0; i < kPer; i++) {
        v[i] = B.cell_cnt[(size_t)page * kCells + threadIdx.x * kPer + i];
        sum += v[i];
    }
    int total;
    int off = block_excl_scan_1024(sum, s_warp, total);
#pragma unroll
    for (int i = 0; i < kPer; i++) {
        B.cell_off[(size_t)page * (kCells + 1) + threadIdx.x * kPer + i] = off;
        B.cell_cur[(size_t)page * kCells + threadIdx.x * kPer + i] = off;
        off += v[i];
    }
    if (threadIdx.x == 0) B.cell_off[(size_t)page * (kCells + 1) + kCells] = total;
}

__global__ void __launch_bounds__(256) nms_bin_scatter_kernel(const int32_t *__restrict__ page_off,
                                                              const int32_t *__restrict__ n_total, LanmsBuffers B)
{
    ms_pdl_wait();
    const int n = *n_total;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
        const int page = B.pos_page[slot];
        const int p0 = page_off[page];
        const int c = slot - p0;
        if (c >= B.cl_count[page] || B.cl_irr[slot]) continue;
        const int pos = atomicAdd(B.cell_cur + (size_t)page * kCells + B.cl_cell[slot], 1);
        B.sb_bbox[(size_t)p0 + pos] = B.cl_bbox[slot];
        B.sb_id[(size_t)p0 + pos] = c;
    }
}

// thread per regular cluster a: every b whose bbox overlaps a's and that a OWNS becomes an (unevaluated) neighbour
// pair, oriented hi -> lo by NMS priority.  Of two overlapping boxes the one with the smaller (min x, index) owns the
// pair: the other's min x then lies in [a.min x, a.max x], so the search needs no margin in x (it was the page's
// largest box width, half of the cells walked), and every pair is still produced exactly once.  b's home cell holds
// its min corner; in y it lies in [a.min y - (max bbox height of the page), a.max y]; each grid row of that range is
// one contiguous span of the cell-ordered arrays.  Hits are kept per thread and appended with ONE atomic per CTA (same-address atomics per
// hit serialise in L2 and cost more than the search itself).
constexpr int kBinThreads = 256;
constexpr int kMaxHits = 32;  // higher-index neighbours one cluster may have (overflow -> MS_FLAG_EDGE_OVERFLOW)

__global__ void __launch_bounds__(kBinThreads) nms_pairs_binned_kernel(const int32_t *__restrict__ page_off,
                                                                       LanmsBuffers B, int32_t *flags)
{
    ms_pdl_wait();
    const int page = blockIdx.y;
    const int p0 = page_off[page];
    const int C = B.cl_count[page];
    const float *ext = B.page_ext + (size_t)page * 8;
    const int32_t *coff = B.cell_off + (size_t)page * (kCells + 1);
    const int edge_cap = (page_off[page + 1] - p0) * B.ef;
    const float ox = ext[0], oy = ext[1], sx = ext[2], sy = ext[3], mh = ext[5];
    __shared__ int s_warp[kBinThreads / 32 + 1];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = blockIdx.x * kBinThreads; base < C; base += gridDim.x * kBinThreads) {
        const int a = base + threadIdx.x;
        int hits[kMaxHits];
        int cnt = 0;
        bool over = false;
        double sa = 0.0;
        if (a < C && B.cl_irr[(size_t)p0 + a] == 0) {
            const float4 ba = B.cl_bbox[(size_t)p0 + a];
            sa = B.cl_score[(size_t)p0 + a];
            // x: only boxes whose min x lies in [a.min x, a.max x] (see the ownership rule above) -- no max-width margin
            const int cx0 = cell_coord(ba.x, ox, sx), cx1 = cell_coord(ba.z, ox, sx);
            const int cy0 = cell_coord(ba.y - mh, oy, sy), cy1 = cell_coord(ba.w, oy, sy);
            for (int cy = cy0; cy <= cy1; cy++) {
                const int pos1 = coff[cy * kGrid + cx1 + 1];
                for (int pos = coff[cy * kGrid + cx0]; pos < pos1; pos++) {
                    const float4 o = B.sb_bbox[(size_t)p0 + pos];
                    if (o.x > ba.z || o.x < ba.x || o.y > ba.w || o.w < ba.y) continue;
                    const int b = B.sb_id[(size_t)p0 + pos];
                    if (o.x == ba.x && b <= a) continue;
                    if (cnt < kMaxHits)
                        hits[cnt++] = b;
                    else
                        over = true;
                }
            }
        }
        // CTA-wide exclusive scan of the hit counts -> one reservation in the page's pair list
        int inc = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        __syncthreads();
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
            for (int w = 0; w < kBinThreads / 32; w++) {
                int t = s_warp[w];
                s_warp[w] = run;
                run += t;
            }
            s_base = run > 0 ? atomicAdd(B.edge_count + page, run) : 0;
        }
        __syncthreads();
        int at = s_base + s_warp[warp] + inc - cnt;
        if (over) atomicOr(B.page_redo + page, 1);
        if (cnt > 0 && at + cnt > edge_cap) atomicOr(flags + page, MS_FLAG_EDGE_OVERFLOW);
        for (int k = 0; k < cnt; k++, at++) {
            if (at >= edge_cap) break;
            const int b = hits[k];
            // cl_orig == cluster index for LANMS clusters; both are regular here
            const bool a_first = prio_before(sa, a, B.cl_score[(size_t)p0 + b], b);
            const int hi = a_first ? a : b, lo = a_first ? b : a;
            B.edges[(size_t)p0 * B.ef + at] = ((uint64_t)(uint32_t)hi << 30) | (uint64_t)(uint32_t)lo;
        }
    }
}

// Exact rebuild of a page's neighbour pairs when some cluster has more than kMaxHits of them (very dense
// candidates, e.g. a score threshold below the background): count, scan, fill.  Pages that did not overflow
// return at once.
template <typename F>
__device__ __forceinline__ void walk_neighbours(const LanmsBuffers &B, int p0, int page, int a, F f)
{
    const float *ext = B.page_ext + (size_t)page * 8;
    const int32_t *coff = B.cell_off + (size_t)page * (kCells + 1);
    const float4 ba = B.cl_bbox[(size_t)p0 + a];
    const int cx0 = cell_coord(ba.x, ext[0], ext[2]), cx1 = cell_coord(ba.z, ext[0], ext[2]);
    const int cy0 = cell_coord(ba.y - ext[5], ext[1], ext[3]), cy1 = cell_coord(ba.w, ext[1], ext[3]);
    for (int cy = cy0; cy <= cy1; cy++) {
        const int pos1 = coff[cy * kGrid + cx1 + 1];
        for (int pos = coff[cy * kGrid + cx0]; pos < pos1; pos++) {
            const float4 o = B.sb_bbox[(size_t)p0 + pos];
            if (o.x > ba.z || o.x < ba.x || o.y > ba.w || o.w < ba.y) continue;  // same ownership rule as the binned kernel
            const int b = B.sb_id[(size_t)p0 + pos];
            if (o.x == ba.x && b <= a) continue;
            f(b);
        }
    }
}

__global__ void __launch_bounds__(kBinThreads) nms_redo_count_kernel(const int32_t *__restrict__ page_off, LanmsBuffers B)
{
    ms_pdl_wait();
    const int page = blockIdx.y;
    if (!B.page_redo[page]) return;
    const int p0 = page_off[page];
    const int C = B.cl_count[page];
    for (int a = blockIdx.x * kBinThreads + threadIdx.x; a < C; a += gridDim.x * kBinThreads) {
        int cnt = 0;
        if (B.cl_irr[(size_t)p0 + a] == 0) walk_neighbours(B, p0, page, a, [&](int) { cnt++; });
        B.nb_cnt[(size_t)p0 + a] = cnt;
    }
}

__global__ void __launch_bounds__(1024) nms_redo_scan_kernel(const int32_t *__restrict__ page_off, LanmsBuffers B,
                                                             int32_t *flags)
{
    ms_pdl_wait();
    const int page = blockIdx.x;
    if (!B.page_redo[page]) return;
    const int p0 = page_off[page];
    const int C = B.cl_count[page];
    __shared__ int s_warp[33];
    long long run = 0;
    for (int base = 0; base < C; base += 1024) {
        const int c = base + threadIdx.x;
        const int v = c < C ? B.nb_cnt[(size_t)p0 + c] : 0;
        int total;
        const int off = block_excl_scan_1024(v, s_warp, total);
        if (c < C) B.nb_cnt[(size_t)p0 + c] = (int)min(run + off, (long long)0x7fffffff);
        run += total;
    }
    if (threadIdx.x == 0) {
        const long long cap = (long long)(page_off[page + 1] - p0) * B.ef;
        if (run > cap) atomicOr(flags + page, MS_FLAG_EDGE_OVERFLOW);
        B.edge_count[page] = (int)min(run, cap);
    }
}

__global__ void __launch_bounds__(kBinThreads) nms_redo_fill_kernel(const int32_t *__restrict__ page_off, LanmsBuffers B)
{
    ms_pdl_wait();
    const int page = blockIdx.y;
    if (!B.page_redo[page]) return;
    const int p0 = page_off[page];
    const int C = B.cl_count[page];
    const long long cap = (long long)(page_off[page + 1] - p0) * B.ef;
    for (int a = blockIdx.x * kBinThreads + threadIdx.x; a < C; a += gridDim.x * kBinThreads) {
        if (B.cl_irr[(size_t)p0 + a]) continue;
        long long at = B.nb_cnt[(size_t)p0 + a];
        const double sa = B.cl_score[(size_t)p0 + a];
        walk_neighbours(B, p0, page, a, [&](int b) {
            if (at < cap) {
                const bool a_first = prio_before(sa, a, B.cl_score[(size_t)p0 + b], b);
                const int hi = a_first ? a : b, lo = a_first ? b : a;
                B.edges[(size_t)p0 * B.ef + at] = ((uint64_t)(uint32_t)hi << 30) | (uint64_t)(uint32_t)lo;
            }
            at++;
        });
    }
}

// ---- 6. greedy recurrence in rounds (one CTA per page), IoU evaluated lazily ----------------------------------
// keep[i] = !any(keep[j] and IoU(j, i) > thr for neighbours j of higher priority).  A box is decided once all its
// higher-priority neighbours are; a pair is clipped (fp64 Sutherland-Hodgman, subject = the kept box as in
// lanms.py:149) only when its higher box is kept and its lower box is still undecided -- in a text page that is
// ~1/4 of the neighbour pairs.  Identical to the sequential loop because the predicate is a pure function of the
// ordered pair.
constexpr int kResolveWarps = 32;

// IoU(a, b) > thr decided WITHOUT clipping, for two regular quads (convex, positively oriented, edges and angles
// bounded away from degenerate by quad_regular_bbox) and thr >= 0.  Shrink `in` about the mean of its vertices by
// lam: the result lies in `in` (convexity); if its four vertices are strictly inside every edge of `out` (the cross
// product of lanms.py:40-45, required to exceed 1e-9 of its own two products -- seven orders above its rounding
// error) it lies in `out` too, so the true intersection area is at least lam^2 * area(in).  lam is chosen so that
// this bound clears   inter * (1 + thr) > thr * (a1 + a2)   (<=> IoU > thr) with 6 % to spare; the reference's
// float64 Sutherland-Hodgman area differs from the true one by ~1e-12 relative for such quads (a crossing point
// stays on its subject edge, and the shoelace sum has no cancellation), so its decision is the same.  Candidates of
// one word decode to nearly the same quad: almost every pair the NMS clips is decided here, in ~1/10 of the
// instructions and without the divergent vertex loops.  false = not proven: the caller clips.
__device__ __forceinline__ bool shrunk_inside(const double *in, const double *out, double lam)
{
    const double cx = 0.25 * (in[0] + in[2] + in[4] + in[6]), cy = 0.25 * (in[1] + in[3] + in[5] + in[7]);
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double vx = cx + lam * (in[2 * k] - cx), vy = cy + lam * (in[2 * k + 1] - cy);
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int e1 = (e + 1) & 3;
            const double t1 = (out[2 * e1] - out[2 * e]) * (vy - out[2 * e + 1]);
            const double t2 = (out[2 * e1 + 1] - out[2 * e + 1]) * (vx - out[2 * e]);
            ok = ok && (t1 - t2 > 1e-9 * (fabs(t1) + fabs(t2)));
        }
    }
    return ok;
}

__device__ __forceinline__ bool iou_above_by_containment(const double *a, const double *b, double thr)
{
    if (!(thr >= 0.0)) return false;
    const double a1 = ms_shoelace(a, 4), a2 = ms_shoelace(b, 4);
    if (!(a1 > 0.0) || !(a2 > 0.0)) return false;
    const double need = 1.06 * thr * (a1 + a2) / (1.0 + thr);  // intersection area that proves IoU > thr
    const double la = fmax(need / a1, 0.0025), lb = fmax(need / a2, 0.0025);  // lam^2, floored (thr == 0)
    if (la < 0.9 && shrunk_inside(a, b, sqrt(la) * 1.0000001)) return true;
    if (lb < 0.9 && shrunk_inside(b, a, sqrt(lb) * 1.0000001)) return true;
    return false;
}

__device__ __forceinline__ void cluster_sync_all()
{
    // release / acquire at cluster scope: global-memory writes of every CTA of the cluster are visible afterwards
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Launched as one thread-block cluster per page (1, 2, 4 or 8 CTAs): pages are few, so the cluster spreads one
// page's pairs over several SMs; state lives in global memory and rounds are separated by cluster barriers.
__global__ void __launch_bounds__(1024) lanms_resolve_kernel(const int32_t *__restrict__ page_off, double thr,
                                                             LanmsBuffers B, int32_t *__restrict__ und_flags)
{
    ms_pdl_wait();
    uint32_t crank, csize;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    const int page = blockIdx.x / csize;
    const int p0 = page_off[page];
    const int C = B.cl_count[page];
    int E = B.edge_count[page];
    const int ecap = (page_off[page + 1] - p0) * B.ef;
    if (E > ecap) E = ecap;
    uint64_t *edges = B.edges + (size_t)p0 * B.ef;
    volatile uint8_t *state = B.state + p0;
    volatile uint8_t *blocked = B.blocked + p0;
    volatile int32_t *und = und_flags + 2 * page;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gthread = crank * blockDim.x + threadIdx.x, gthreads = csize * blockDim.x;
    __shared__ int s_q[kResolveWarps][64], s_q2[kResolveWarps][64];
    int *q = s_q[warp], *q2 = s_q2[warp];
    int q2n = 0;  // warp-uniform: pairs the containment test could not decide, waiting for the clip
    double buf[4 * MS_MAXV];

    auto clip_full = [&](int e) {
        const uint64_t ed = edges[e];
        const int hi = (int)((ed >> 30) & 0x3fffffffu), lo = (int)(ed & 0x3fffffffu);
        double qh[8], ql[8];
        load_quad(B.cl_poly + (size_t)(p0 + hi) * 8, qh);
        load_quad(B.cl_poly + (size_t)(p0 + lo) * 8, ql);
        if (ms_quad_iou(qh, ql, buf) > thr)
            state[lo] = 2;
        else
            edges[e] = ed | kPairDone;
    };
    // one pair per active lane (e < 0: none): decided by containment, or queued for the clip, 32 at a time
    auto clip = [&](int e) {
        bool slow = false;
        if (e >= 0) {
            const uint64_t ed = edges[e];
            const int hi = (int)((ed >> 30) & 0x3fffffffu), lo = (int)(ed & 0x3fffffffu);
            double qh[8], ql[8];
            load_quad(B.cl_poly + (size_t)(p0 + hi) * 8, qh);
            load_quad(B.cl_poly + (size_t)(p0 + lo) * 8, ql);
            const bool reg = B.cl_irr[(size_t)p0 + hi] == 0 && B.cl_irr[(size_t)p0 + lo] == 0;
            if (reg && iou_above_by_containment(qh, ql, thr))
                state[lo] = 2;
            else
                slow = true;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, slow);
        if (slow) q2[q2n + __popc(m & ((1u << lane) - 1u))] = e;
        q2n += __popc(m);
        __syncwarp();
        if (q2n >= 32) {
            const int e2 = q2[q2n - 32 + lane];
            __syncwarp();
            clip_full(e2);
            q2n -= 32;
            __syncwarp();
        }
    };

    for (int round = 0;; round++) {
        for (int c = gthread; c < C; c += gthreads) blocked[c] = 0;
        if (gthread == 0) und[round & 1] = 0;
        cluster_sync_all();
        int qn = 0;  // warp-uniform
        // kEB pairs per lane and iteration: the pair words, then the two state bytes of every pair, are independent
        // loads in flight together (the loop is bound by their latency, not by bandwidth).  A state byte read a
        // little early is at worst a stale 0, which only defers the pair to the next round.
        constexpr int kEB = 4;
        for (int base = (crank * kResolveWarps + warp) * 32 * kEB; base < E; base += csize * kResolveWarps * 32 * kEB) {
            uint64_t ed[kEB];
            uint8_t slo[kEB], shi[kEB];
#pragma unroll
            for (int k = 0; k < kEB; k++) {
                const int e = base + k * 32 + lane;
                ed[k] = e < E ? edges[e] : kPairDone;  // each pair word is only ever touched by this thread
            }
#pragma unroll
            for (int k = 0; k < kEB; k++) {
                slo[k] = shi[k] = 0;
                if (!(ed[k] & kPairDone)) {
                    slo[k] = state[(int)(ed[k] & 0x3fffffffu)];
                    shi[k] = state[(int)((ed[k] >> 30) & 0x3fffffffu)];
                }
            }
#pragma unroll
            for (int k = 0; k < kEB; k++) {
                const int e = base + k * 32 + lane;
                bool need = false;
                if (!(ed[k] & kPairDone)) {
                    const int lo = (int)(ed[k] & 0x3fffffffu);
                    if (slo[k] != 0) {
                        edges[e] = ed[k] | kPairDone;
                    } else if (shi[k] == 1) {
                        if (ed[k] & kPairTrue)
                            state[lo] = 2;
                        else
                            need = true;
                    } else if (shi[k] == 0) {
                        blocked[lo] = 1;
                    } else {
                        edges[e] = ed[k] | kPairDone;  // a suppressed box suppresses nothing
                    }
                }
                const uint32_t m = __ballot_sync(0xffffffffu, need);
                if (need) q[qn + __popc(m & ((1u << lane) - 1u))] = e;
                qn += __popc(m);
                __syncwarp();
                if (qn >= 32) {
                    const int e2 = q[qn - 32 + lane];
                    __syncwarp();
                    clip(e2);
                    qn -= 32;
                    __syncwarp();
                }
            }
        }
        clip(lane < qn ? q[lane] : -1);
        if (lane < q2n) clip_full(q2[lane]);
        q2n = 0;
        cluster_sync_all();
        bool any_und = false;
        for (int c = gthread; c < C; c += gthreads) {
            if (state[c] == 0) {
                if (!blocked[c])
                    state[c] = 1;
                else
                    any_und = true;
            }
        }
        if (any_und) und[round & 1] = 1;
        cluster_sync_all();
        if (!und[round & 1]) break;
    }
}

static int launch_resolve(ms_ctx *ctx, int n_pages, const int32_t *page_off, double thr, const LanmsBuffers &B,
                          int32_t *und_flags, cudaStream_t st)
{
    int cs = 1;
    while (cs < 8 && n_pages * cs * 2 <= ctx->num_sms) cs *= 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_pages * cs), 1, 1);
    cfg.blockDim = dim3(1024, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ms_pdl_enabled() ? 2 : 1;
    MS_CUDA(cudaLaunchKernelEx(&cfg, lanms_resolve_kernel, page_off, thr, B, und_flags));
    MS_LAUNCH_CHECK(ctx);
    return MS_OK;
}

// position key of np.argsort(-scores, kind="stable"): ascending key == larger score first, NaN last, -0 == +0
__device__ __forceinline__ uint64_t desc_score_key(double sc)
{
    if (sc != sc) return ~0ull;
    if (sc == 0.0) sc = 0.0;
    uint64_t u = (uint64_t)__double_as_longlong(sc);
    u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);  // ascending image
    return ~u;
}

// ---- 7. kept clusters -> rows in stable descending-score order ----------------------------------------------------
__global__ void __launch_bounds__(1024) lanms_kept_kernel(const int32_t *__restrict__ page_off, LanmsBuffers B,
                                                          int32_t *__restrict__ counts_out, int f32_scores)
{
    ms_pdl_wait();
    const int page = blockIdx.x;
    const int p0 = page_off[page];
    const int C = B.cl_count[page];
    __shared__ int s_warp[33];
    int32_t *kept = B.kept_list + p0;
    int run_base = 0;
    // 8 consecutive clusters per thread and round: the state bytes of a round are independent loads in flight together
    // and a 14k-cluster page needs two block scans instead of fourteen
    for (int base = 0; base < C; base += 1024 * 8) {
        const int c0 = base + threadIdx.x * 8;
        uint32_t bits = 0;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (c0 + k < C && B.state[p0 + c0 + k] == 1) bits |= 1u << k;
        int total;
        int pos = run_base + block_excl_scan_1024(__popc(bits), s_warp, total);
        while (bits) {
            const int c = c0 + __ffs(bits) - 1;
            bits &= bits - 1;
            kept[pos] = c;
            B.kept_key[p0 + pos] = desc_score_key(B.cl_score[p0 + c]);
            if (f32_scores) {
                // LANMS cluster scores are float32 values: a 32-bit descending key for the large-page radix sort
                const float sc = (float)B.cl_score[p0 + c];
                B.keys[p0 + pos] = (sc != sc) ? 0xFFFFFFFFu : ~ms_orderable_f32(sc);
                B.vals[p0 + pos] = (uint32_t)c;
            }
            pos++;
        }
        run_base += total;
    }
    if (threadIdx.x == 0) counts_out[page] = run_base;
}

// Pages with at most kSortMax kept boxes: order them with a bitonic sort in shared memory.  LANMS cluster scores are
// float32 values (a max over float32 inputs), so (descending-score key, cluster index) packs into one 64-bit key
// whose ascending order is exactly np.argsort(-scores, kind="stable").  Larger pages are left to the ranking kernel.
constexpr int kSortMax = 4096;

__global__ void __launch_bounds__(1024) lanms_sort_emit_kernel(const int32_t *__restrict__ page_off, int cap,
                                                               LanmsBuffers B, float *__restrict__ out,
                                                               const int32_t *__restrict__ counts_out)
{
    ms_pdl_wait();
    const int page = blockIdx.x;
    const int p0 = page_off[page];
    const int K = counts_out[page];
    if (K > kSortMax) return;  // ordered by the segmented radix sort + lanms_emit_sorted_kernel
    __shared__ uint64_t s_key[kSortMax];
    const int32_t *kept = B.kept_list + p0;
    int n = 1;
    while (n < K) n <<= 1;
    for (int i = threadIdx.x; i < n; i += 1024) {
        uint64_t key = ~0ull;
        if (i < K) {
            const int c = kept[i];
            const float sc = (float)B.cl_score[p0 + c];
            const uint32_t k32 = (sc != sc) ? 0xFFFFFFFFu : ~ms_orderable_f32(sc);
            key = ((uint64_t)k32 << 32) | (uint32_t)c;
        }
        s_key[i] = key;
    }
    __syncthreads();
    // Bitonic network, one compare-exchange PAIR per thread and stage (pair pr -> element i: a zero bit inserted at
    // log2 j), so no lane idles.  A page is one CTA on one SM and the sort is bound by that SM's issue rate, so
    // instructions per stage are what counts.  Stages with j <= 32 stay inside the 64-element block a warp owns
    // (pairs 32w .. 32w + 31): between two such stages a warp barrier is enough; any stage next to a wider one needs
    // the CTA barrier.
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool first = j == (k >> 1);
            if (first ? j <= 32 : j <= 16)
                __syncwarp();
            else
                __syncthreads();
            for (int pr = threadIdx.x; pr < (n >> 1); pr += 1024) {
                const int i = ((pr & ~(j - 1)) << 1) | (pr & (j - 1));
                const int ixj = i | j;
                const uint64_t a = s_key[i], b = s_key[ixj];
                const bool up = (i & k) == 0;
                if ((a > b) == up) {
                    s_key[i] = b;
                    s_key[ixj] = a;
                }
            }
        }
    }
    __syncthreads();
    // the order only: the rows are gathered by lanms_emit_sorted_kernel, which has the whole GPU instead of one SM
    for (int r = threadIdx.x; r < K; r += 1024) B.vals[p0 + r] = (uint32_t)(s_key[r] & 0xffffffffu);
}

// B.vals holds every page's kept cluster indices in output order (shared-memory sort, or the radix sort for pages
// with more than kSortMax kept boxes): gather the rows
__global__ void __launch_bounds__(256) lanms_emit_sorted_kernel(const int32_t *__restrict__ page_off, int cap,
                                                                LanmsBuffers B, float *__restrict__ out,
                                                                const int32_t *__restrict__ counts_out)
{
    ms_pdl_wait();
    const int page = blockIdx.y;
    const int K = counts_out[page];
    const int p0 = page_off[page];
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < K; r += gridDim.x * blockDim.x) {
        const int c = (int)B.vals[p0 + r];
        float *row = out + ((size_t)page * cap + r) * 9;
        const double *poly = B.cl_poly + (size_t)(p0 + c) * 8;
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) row[k2] = (float)poly[k2];  // lanms.py:207 astype(float32)
        row[8] = (float)B.cl_score[p0 + c];
    }
}

constexpr int kRankThreads = 256;

// rank of every kept cluster among the kept ones of its page (np.argsort(-scores, kind="stable") position)
__global__ void __launch_bounds__(kRankThreads) lanms_emit_kernel(const int32_t *__restrict__ page_off, int cap,
                                                                  LanmsBuffers B, float *__restrict__ out,
                                                                  const int32_t *__restrict__ counts_out,
                                                                  int32_t *__restrict__ keep_idx_out,
                                                                  const int32_t *__restrict__ rank_needed)
{
    ms_pdl_wait();
    const int page = blockIdx.y;
    if (rank_needed && !rank_needed[page]) return;
    const int p0 = page_off[page];
    const int K = counts_out[page];
    const int32_t *kept = B.kept_list + p0;
    __shared__ uint64_t s_key[kRankThreads];
    __shared__ int s_oc[kRankThreads];
    const uint64_t *kkey = B.kept_key + p0;
    for (int base = blockIdx.x * kRankThreads; base < K; base += gridDim.x * kRankThreads) {
        const int i = base + threadIdx.x;
        const bool live = i < K;
        int c = 0, oc = 0;
        uint64_t key = 0;
        if (live) {
            c = kept[i];
            key = kkey[i];
            oc = B.cl_orig[p0 + c];
        }
        int rank = 0;
        for (int tb = 0; tb < K; tb += kRankThreads) {
            __syncthreads();
            const int j = tb + threadIdx.x;
            if (j < K) {
                s_key[threadIdx.x] = kkey[j];
                s_oc[threadIdx.x] = B.cl_orig[p0 + kept[j]];
            }
            __syncthreads();
            const int nb = min(kRankThreads, K - tb);
            if (live)
                for (int t = 0; t < nb; t++) {
                    const uint64_t k2 = s_key[t];
                    rank += (k2 < key || (k2 == key && s_oc[t] < oc)) ? 1 : 0;
                }
        }
        if (live) {
            if (out) {
                float *row = out + ((size_t)page * cap + rank) * 9;
                const double *poly = B.cl_poly + (size_t)(p0 + c) * 8;
#pragma unroll
                for (int k2 = 0; k2 < 8; k2++) row[k2] = (float)poly[k2];  // lanms.py:207 astype(float32)
                row[8] = (float)B.cl_score[p0 + c];
            }
            if (keep_idx_out) keep_idx_out[(size_t)page * cap + rank] = oc;
        }
    }
}

// ---- standalone helpers ---------------------------------------------------------------------------------------------
__global__ void iou_pairs_kernel(const double *__restrict__ subj, const double *__restrict__ clip, int64_t n,
                                 double *__restrict__ iou)
{
    ms_pdl_wait();
    double buf[4 * MS_MAXV];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double a[8], b[8];
        load_quad(subj + i * 8, a);
        load_quad(clip + i * 8, b);
        iou[i] = ms_quad_iou(a, b, buf);
    }
}

// test-only (ms_test_iou_proved_host): the two predicates the resolve kernel uses to skip the float64 clip, evaluated
// by the device functions themselves.  bit 0: both quads regular (quad_regular_bbox); bit 1: regular and
// iou_above_by_containment(a, b, thr) -- "IoU(a, b) > thr is proven".
__global__ void iou_proved_kernel(const double *__restrict__ subj, const double *__restrict__ clip, int64_t n,
                                  double thr, uint8_t *__restrict__ out)
{
    ms_pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double a[8], b[8];
        load_quad(subj + i * 8, a);
        load_quad(clip + i * 8, b);
        float4 bb;
        const bool reg = quad_regular_bbox(a, bb) && quad_regular_bbox(b, bb);
        out[i] = (uint8_t)((reg ? 1 : 0) | ((reg && iou_above_by_containment(a, b, thr)) ? 2 : 0));
    }
}

// clusters given directly (standard_nms): every box is treated as irregular => exact all-pairs mode
__global__ void nms_prepare_kernel(const double *__restrict__ polys, const double *__restrict__ scores, int n,
                                   LanmsBuffers B)
{
    ms_pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        B.page_off[0] = 0;
        B.page_off[1] = n;
        *B.n_total = n;
        B.cl_count[0] = n;
        B.edge_count[0] = 0;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 8; k++) B.cl_poly[(size_t)i * 8 + k] = polys[(size_t)i * 8 + k];
        B.cl_score[i] = scores[i];
        B.cl_bbox[i] = make_float4(-INFINITY, -INFINITY, INFINITY, INFINITY);
        B.cl_irr[i] = 1;
        B.cl_orig[i] = i;
        B.state[i] = 0;
        B.pos_page[i] = 0;
    }
}

size_t carve(ms_bump &bump, LanmsBuffers &B, int n_pages, size_t n_max, bool full, int ef)
{
    B = LanmsBuffers{};
    B.page_off = bump.take<int32_t>(n_pages + 1);
    B.n_total = bump.take<int32_t>(1);
    B.hot_count = bump.take<int32_t>(1);
    B.cl_count = bump.take<int32_t>(n_pages);
    B.edge_count = bump.take<int32_t>(n_pages);
    B.page_ext = bump.take<float>((size_t)n_pages * 8);
    B.cell_cnt = bump.take<int32_t>((size_t)n_pages * kCells);
    B.cell_off = bump.take<int32_t>((size_t)n_pages * (kCells + 1));
    B.cell_cur = bump.take<int32_t>((size_t)n_pages * kCells);
    B.cl_cell = bump.take<int32_t>(n_max);
    B.sb_bbox = bump.take<float4>(n_max);
    B.sb_id = bump.take<int32_t>(n_max);
    B.pos_page = bump.take<int32_t>(n_max);
    if (full) {
        B.keys = bump.take<uint32_t>(n_max);
        B.keys_tmp = bump.take<uint32_t>(n_max);
        B.vals = bump.take<uint32_t>(n_max);
        B.vals_tmp = bump.take<uint32_t>(n_max);
        B.sq = bump.take<double>(n_max * 8);
        B.ss = bump.take<float>(n_max);
        B.mflag = bump.take<uint8_t>(n_max);
        B.hot = bump.take<uint8_t>(n_max);
        B.hot_list = bump.take<int32_t>(n_max);
        B.run_end = bump.take<int32_t>(n_max);
        B.run_poly = bump.take<double>(n_max * 8);
        B.run_score = bump.take<double>(n_max);
    }
    B.cl_poly = bump.take<double>(n_max * 8);
    B.cl_score = bump.take<double>(n_max);
    B.cl_bbox = bump.take<float4>(n_max);
    B.cl_irr = bump.take<uint8_t>(n_max);
    B.cl_orig = bump.take<int32_t>(n_max);
    B.ef = ef;
    B.edges = bump.take<uint64_t>(n_max * (size_t)ef);
    B.page_redo = bump.take<int32_t>(n_pages);
    B.nb_cnt = bump.take<int32_t>(n_max);
    B.state = bump.take<uint8_t>(n_max);
    B.blocked = bump.take<uint8_t>(n_max);
    B.irr_list = bump.take<int32_t>(n_max);
    B.irr_count = bump.take<int32_t>(1);
    B.und_flags = bump.take<int32_t>((size_t)n_pages * 2);
    B.kept_key = bump.take<uint64_t>(n_max);
    B.kept_list = bump.take<int32_t>(n_max);
    B.acc_count = bump.take<int32_t>(n_pages);
    B.ext_acc = bump.take<int32_t>((size_t)n_pages * 8);
    B.ext_done = bump.take<int32_t>(n_pages);
    return bump.off;
}

}  // namespace

size_t msk_lanms_scratch(int n_pages, int cap_per_page, int ef)
{
    ms_bump probe{nullptr, 0, 0};
    LanmsBuffers B;
    size_t n_max = (size_t)n_pages * cap_per_page;
    size_t core = carve(probe, B, n_pages, n_max, true, ef);
    return core + msk_sort_pages_scratch(n_pages, cap_per_page) + 4096;
}

size_t msk_standard_nms_scratch(int n, int ef)
{
    ms_bump probe{nullptr, 0, 0};
    LanmsBuffers B;
    return carve(probe, B, 1, (size_t)n, false, ef) + 4096;
}

int msk_lanms(ms_ctx *ctx, const float *quads, const int32_t *counts, int n_pages, int cap_per_page, double thr,
              float *quads_out, int32_t *counts_out, int32_t *flags, ms_bump bump, cudaStream_t st)
{
    if (n_pages <= 0) return MS_OK;
    const size_t n_max = (size_t)n_pages * cap_per_page;
    if (n_max >= (size_t)1 << 31) {
        ms_set_error("lanms: n_pages*cap_per_page too large");
        return MS_ERR_INVALID;
    }
    LanmsBuffers B;
    carve(bump, B, n_pages, n_max, true, ctx->edge_factor);
    if (!B.kept_list) {
        ms_set_error("lanms: scratch too small");
        return MS_ERR_CAPACITY;
    }
    const int sms = ctx->num_sms;
    ms_launch(lanms_offsets_kernel, 1, 32, 0, st, counts, n_pages, B.page_off, B.n_total, B.hot_count, B.edge_count,
                                           B.irr_count, B.page_redo);
    MS_LAUNCH_CHECK(ctx);
    {
        size_t threads = n_max;
        int grid = (int)((threads + 255) / 256);
        ms_launch(lanms_keys_kernel, grid, 256, 0, st, quads, counts, B.page_off, n_pages, cap_per_page, B.keys, B.vals,
                                                B.pos_page);
        MS_LAUNCH_CHECK(ctx);
    }
    int rc = msk_sort_pages(ctx, B.keys, B.vals, B.keys_tmp, B.vals_tmp, B.page_off, nullptr, -1, n_pages, cap_per_page,
                            bump, st);
    if (rc != MS_OK) return rc;
    ms_launch(lanms_gather_hot_kernel, sms * 8, 128, 0, st, quads, B.vals, B.page_off, B.n_total, thr, B);
    MS_LAUNCH_CHECK(ctx);
    ms_launch(lanms_runs_kernel, sms * 4, 128, 0, st, B.page_off, thr, B);
    MS_LAUNCH_CHECK(ctx);
    ms_launch(lanms_accept_kernel, n_pages, kResolveThreads, 0, st, B.page_off, B);
    MS_LAUNCH_CHECK(ctx);
    {
        // a few CTAs per page, each striding over the page's positions (the counts live on the device)
        int gx = (cap_per_page + kBuildThreads - 1) / kBuildThreads;
        const int want = (sms * 8 + n_pages - 1) / n_pages;
        if (gx > want) gx = want;
        if (gx < 1) gx = 1;
        ms_launch(lanms_build_clusters_kernel, dim3((unsigned)gx, (unsigned)n_pages), kBuildThreads, 0, st, B.page_off, thr < 0 ? 1 : 0, B);
    }
    MS_LAUNCH_CHECK(ctx);
    {
        int g = (int)((n_max + 255) / 256);
        if (g > sms * 8) g = sms * 8;
        ms_launch(nms_bin_count_kernel, g, 256, 0, st, B.page_off, B.n_total, B);
        MS_LAUNCH_CHECK(ctx);
        ms_launch(nms_bin_scan_kernel, n_pages, 1024, 0, st, B);
        MS_LAUNCH_CHECK(ctx);
        ms_launch(nms_bin_scatter_kernel, g, 256, 0, st, B.page_off, B.n_total, B);
        MS_LAUNCH_CHECK(ctx);
        int gx = (cap_per_page + kBinThreads - 1) / kBinThreads;
        if (gx > 64) gx = 64;
        ms_launch(nms_pairs_binned_kernel, dim3(gx, n_pages), kBinThreads, 0, st, B.page_off, B, flags);
        MS_LAUNCH_CHECK(ctx);
        ms_launch(nms_redo_count_kernel, dim3(gx, n_pages), kBinThreads, 0, st, B.page_off, B);
        MS_LAUNCH_CHECK(ctx);
        ms_launch(nms_redo_scan_kernel, n_pages, 1024, 0, st, B.page_off, B, flags);
        MS_LAUNCH_CHECK(ctx);
        ms_launch(nms_redo_fill_kernel, dim3(gx, n_pages), kBinThreads, 0, st, B.page_off, B);
        MS_LAUNCH_CHECK(ctx);
    }
    ms_launch(lanms_pairs_kernel, sms * 4, kPairThreads, 0, st, B.page_off, B.n_total, thr, B, flags, 1);
    MS_LAUNCH_CHECK(ctx);
    {
        int rc2 = launch_resolve(ctx, n_pages, B.page_off, thr, B, B.und_flags, st);
        if (rc2 != MS_OK) return rc2;
    }
    ms_launch(lanms_kept_kernel, n_pages, 1024, 0, st, B.page_off, B, counts_out, 1);
    MS_LAUNCH_CHECK(ctx);
    {
        int gx = (cap_per_page + kRankThreads - 1) / kRankThreads;
        if (gx > 32) gx = 32;
        ms_launch(lanms_sort_emit_kernel, n_pages, 1024, 0, st, B.page_off, cap_per_page, B, quads_out, counts_out);
        MS_LAUNCH_CHECK(ctx);
        // pages with more kept boxes than the shared-memory sort holds: stable radix sort of (descending score key,
        // cluster index) -- the kernels return at once for the other pages
        int rc2 = msk_sort_pages(ctx, B.keys, B.vals, B.keys_tmp, B.vals_tmp, B.page_off, counts_out, kSortMax, n_pages,
                                 cap_per_page, bump, st);
        if (rc2 != MS_OK) return rc2;
        ms_launch(lanms_emit_sorted_kernel, dim3(gx, n_pages), 256, 0, st, B.page_off, cap_per_page, B, quads_out, counts_out);
        MS_LAUNCH_CHECK(ctx);
    }
    return MS_OK;
}

int msk_standard_nms(ms_ctx *ctx, const double *polys, const double *scores, int n, double thr, int32_t *keep_idx,
                     int32_t *k_out, int32_t *flags, ms_bump bump, cudaStream_t st)
{
    if (n <= 0) return MS_OK;
    LanmsBuffers B;
    carve(bump, B, 1, (size_t)n, false, ctx->edge_factor);
    if (!B.kept_list) {
        ms_set_error("standard_nms: scratch too small");
        return MS_ERR_CAPACITY;
    }
    const int sms = ctx->num_sms;
    ms_launch(nms_prepare_kernel, sms, 256, 0, st, polys, scores, n, B);
    MS_LAUNCH_CHECK(ctx);
    ms_launch(lanms_pairs_kernel, sms * 8, kPairThreads, 0, st, B.page_off, B.n_total, thr, B, flags, 0);
    MS_LAUNCH_CHECK(ctx);
    {
        int rc2 = launch_resolve(ctx, 1, B.page_off, thr, B, B.und_flags, st);
        if (rc2 != MS_OK) return rc2;
    }
    ms_launch(lanms_kept_kernel, 1, 1024, 0, st, B.page_off, B, k_out, 0);
    MS_LAUNCH_CHECK(ctx);
    {
        int gx = (n + kRankThreads - 1) / kRankThreads;
        if (gx > 148) gx = 148;
        ms_launch(lanms_emit_kernel, dim3(gx, 1), kRankThreads, 0, st, B.page_off, n, B, nullptr, k_out, keep_idx, nullptr);
        MS_LAUNCH_CHECK(ctx);
    }
    return MS_OK;
}

int msk_iou_proved(ms_ctx *ctx, const double *subj, const double *clip, int64_t n, double thr, uint8_t *out,
                   cudaStream_t st)
{
    if (n <= 0) return MS_OK;
    int grid = (int)((n + 127) / 128);
    if (grid > ctx->num_sms * 16) grid = ctx->num_sms * 16;
    ms_launch(iou_proved_kernel, grid, 128, 0, st, subj, clip, n, thr, out);
    MS_LAUNCH_CHECK(ctx);
    return MS_OK;
}

int msk_polygon_iou(ms_ctx *ctx, const double *subj, const double *clip, int64_t n, double *iou, cudaStream_t st)
{
    if (n <= 0) return MS_OK;
    int grid = (int)((n + 127) / 128);
    if (grid > ctx->num_sms * 16) grid = ctx->num_sms * 16;
    ms_launch(iou_pairs_kernel, grid, 128, 0, st, subj, clip, n, iou);
    MS_LAUNCH_CHECK(ctx);
    return MS_OK;
}
