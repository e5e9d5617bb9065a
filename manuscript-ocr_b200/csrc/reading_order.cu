// Reading-order sort of a page's word boxes on the device -- the host logic that sits between NMS and the crop
// loop in the reference: resolve_intersections / sort_boxes_reading_order /
// sort_boxes_reading_order_with_resolutions (reference detectors/_east/utils.py:500-644) and the word re-matching
// of Pipeline.predict (_pipeline.py:105-123).  Result for result identical, including the quirks (the dict that
// keeps the LAST box of equal compressed coordinates, the re-match that picks the FIRST word of equal integer box).
//
// The reference's shrink loop is a sequential recurrence over (i, j) pairs in index order.  Two facts make it cheap
// without changing a single step: boxes only ever shrink towards their top-left corner, so a pair that does not
// intersect initially never does (the sweeps visit only the initially intersecting pairs, in the reference's order,
// and drop pairs that stopped intersecting); box centres are multiples of 0.5, so per-line means are exact integer
// sums divided once.  Within a sweep two pairs commute unless they share a box: every pair gets a dependency level
// (1 + the level of the latest earlier pair touching either of its boxes) and the pairs of one level are applied in
// parallel -- any order that respects the levels reproduces the sequential sweep.
// One CTA per page, everything out of shared memory; the line assignment is a one-warp sequential scan.
#include "ms_internal.cuh"

namespace {

constexpr int kRoThreads = 1024;
constexpr int kRoMaxBoxes = 4096;   // boxes per page held in shared memory
constexpr int kRoMaxPairs = 28672;  // initially intersecting pairs per page

__device__ __forceinline__ int ro_trunc(float v)
{
    // np.array(polygon, dtype=np.int32): C cast of a double; out of range / NaN -> INT_MIN as on x86
    if (!(v > -2147483904.0f && v < 2147483648.0f)) return INT_MIN;
    return (int)v;
}

__device__ __forceinline__ bool ro_intersect(const int4 a, const int4 b)
{
    // utils.py:516-519
    return !(a.z <= b.x || b.z <= a.x || a.w <= b.y || b.w <= a.y);
}

__device__ __forceinline__ int ro_shrink(int lo, int hi)
{
    // int(x1 - (x1 - x0) * 0.1) in Python floats (utils.py:531-542)
    return (int)((double)hi - (double)(hi - lo) * 0.1);
}

__device__ __forceinline__ int ro_block_scan(int v, int *s_warp, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    __syncthreads();
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

__device__ __forceinline__ void ro_bitonic(uint64_t *keys, int n)
{
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += kRoThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t a = keys[i], b = keys[ixj];
                    if ((a > b) == ((i & k) == 0)) {
                        keys[i] = b;
                        keys[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// orderable image of a double (ascending), NaN last
__device__ __forceinline__ uint64_t ro_orderable(double d)
{
    if (d != d) return ~0ull;
    if (d == 0.0) d = 0.0;
    uint64_t u = (uint64_t)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// boxes8: rows of `row_stride` floats whose first 8 are x0,y0..x3,y3 (page-strided with `cap` rows per page)
__global__ void __launch_bounds__(kRoThreads) reading_order_kernel(const float *__restrict__ boxes8, int row_stride,
                                                                   const int32_t *__restrict__ counts, int cap,
                                                                   int4 *__restrict__ obox_g, int32_t *__restrict__ order,
                                                                   float *__restrict__ reordered, int32_t *__restrict__ flags)
{
    extern __shared__ __align__(16) unsigned char ro_smem[];
    int4 *box = reinterpret_cast<int4 *>(ro_smem);                                   // kRoMaxBoxes
    uint32_t *pairs = reinterpret_cast<uint32_t *>(ro_smem + kRoMaxBoxes * 16);      // kRoMaxPairs
    uint8_t *lv = reinterpret_cast<uint8_t *>(pairs + kRoMaxPairs);                  // kRoMaxPairs dependency levels
    uint16_t *last_lv = reinterpret_cast<uint16_t *>(lv + kRoMaxPairs);              // kRoMaxBoxes
    // after the sweeps the pair + level region (148 KB) is reused:
    unsigned char *R = reinterpret_cast<unsigned char *>(pairs);
    uint64_t *keys = reinterpret_cast<uint64_t *>(R);                         // kRoMaxBoxes u64          @   0 K
    long long *line_sum = reinterpret_cast<long long *>(R + 32 * 1024);       // per line: sum of y0 + y1  @  32 K
    double *line_cy = reinterpret_cast<double *>(R + 64 * 1024);              // per line: mean centre     @  64 K
    int *line_cnt = reinterpret_cast<int *>(R + 96 * 1024);                   // per line: members         @  96 K
    uint16_t *line_rank = reinterpret_cast<uint16_t *>(R + 112 * 1024);       // per line                  @ 112 K
    uint16_t *line_of = reinterpret_cast<uint16_t *>(R + 120 * 1024);         // per box                   @ 120 K
    uint16_t *seq_of = reinterpret_cast<uint16_t *>(R + 128 * 1024);          // per box                   @ 128 K
    int4 *obox_s = reinterpret_cast<int4 *>(R + 32 * 1024);                   // phase E: original boxes   @  32 K
    __shared__ int s_warp[33];
    __shared__ int s_np, s_lines, s_avg_pos, s_changed, s_maxl;
    __shared__ long long s_hsum;
    __shared__ double s_ytol;

    const int page = blockIdx.x;
    const int K = counts[page];
    const size_t pb = (size_t)page * cap;
    if (K > kRoMaxBoxes) {
        if (threadIdx.x == 0) atomicOr(flags + page, MS_FLAG_CAND_OVERFLOW);
        for (int k = threadIdx.x; k < K; k += kRoThreads) order[pb + k] = k;
        return;
    }
    // A. integer boxes (_pipeline.py:105-109)
    for (int k = threadIdx.x; k < K; k += kRoThreads) {
        const float *q = boxes8 + (pb + k) * row_stride;
        int xmin, xmax, ymin, ymax;
        xmin = xmax = ro_trunc(q[0]);
        ymin = ymax = ro_trunc(q[1]);
#pragma unroll
        for (int v = 1; v < 4; v++) {
            const int x = ro_trunc(q[2 * v]), y = ro_trunc(q[2 * v + 1]);
            xmin = min(xmin, x);
            xmax = max(xmax, x);
            ymin = min(ymin, y);
            ymax = max(ymax, y);
        }
        const int4 b = make_int4(xmin, ymin, xmax, ymax);
        box[k] = b;
        obox_g[pb + k] = b;
    }
    __syncthreads();

    // B. initially intersecting pairs (i < j) in (i, j) order: count, scan, fill
    int run = 0;
    bool overflow = false;
    for (int base = 0; base < K; base += kRoThreads) {
        const int i = base + threadIdx.x;
        int cnt = 0;
        int4 bi = make_int4(0, 0, 0, 0);
        if (i < K) {
            bi = box[i];
            for (int j = i + 1; j < K; j++) cnt += ro_intersect(bi, box[j]) ? 1 : 0;
        }
        int total;
        int off = run + ro_block_scan(cnt, s_warp, total);
        if (i < K && cnt > 0) {
            for (int j = i + 1; j < K; j++) {
                if (ro_intersect(bi, box[j])) {
                    if (off < kRoMaxPairs) pairs[off] = ((uint32_t)i << 16) | (uint32_t)j;
                    off++;
                }
            }
        }
        run += total;
    }
    if (run > kRoMaxPairs) {
        overflow = true;
        run = kRoMaxPairs;
    }
    if (threadIdx.x == 0) {
        s_np = run;
        if (overflow) atomicOr(flags + page, MS_FLAG_EDGE_OVERFLOW);
    }
    __syncthreads();

    // C. the shrink sweeps (utils.py:521-545).  Dependency levels first (one sequential pass over the pair list) ...
    for (int k = threadIdx.x; k < K; k += kRoThreads) last_lv[k] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int np = s_np;
        int maxl = 0;
        for (int p = 0; p < np; p++) {
            const uint32_t pr = pairs[p];
            const int i = (int)(pr >> 16), j = (int)(pr & 0xffffu);
            const int l = max((int)last_lv[i], (int)last_lv[j]) + 1;
            if (l > 255) {
                maxl = -1;  // too deep for the byte-sized levels: sequential sweeps below
                break;
            }
            last_lv[i] = last_lv[j] = (uint16_t)l;
            lv[p] = (uint8_t)l;
            maxl = max(maxl, l);
        }
        s_maxl = maxl;
    }
    __syncthreads();
    if (s_maxl >= 0) {
        // ... then up to 50 sweeps, each level's pairs in parallel; a pair that stopped intersecting gets level 0
        const int np = s_np, maxl = s_maxl;
        for (int sweep = 0; sweep < 50; sweep++) {
            if (threadIdx.x == 0) s_changed = 0;
            __syncthreads();
            for (int l = 1; l <= maxl; l++) {
                for (int p = threadIdx.x; p < np; p += kRoThreads) {
                    if (lv[p] != l) continue;
                    const uint32_t pr = pairs[p];
                    const int i = (int)(pr >> 16), j = (int)(pr & 0xffffu);
                    int4 a = box[i], c = box[j];
                    if (ro_intersect(a, c)) {
                        a.z = ro_shrink(a.x, a.z);
                        a.w = ro_shrink(a.y, a.w);
                        c.z = ro_shrink(c.x, c.z);
                        c.w = ro_shrink(c.y, c.w);
                        box[i] = a;
                        box[j] = c;
                        s_changed = 1;
                    } else {
                        lv[p] = 0;
                    }
                }
                __syncthreads();
            }
            const int changed = s_changed;
            __syncthreads();
            if (!changed) break;
        }
    } else if (threadIdx.x == 0) {
        int np = s_np;
        for (int sweep = 0; sweep < 50; sweep++) {
            bool changed = false;
            int w = 0;
            for (int p = 0; p < np; p++) {
                const uint32_t pr = pairs[p];
                const int i = (int)(pr >> 16), j = (int)(pr & 0xffffu);
                int4 a = box[i], c = box[j];
                if (ro_intersect(a, c)) {
                    a.z = ro_shrink(a.x, a.z);
                    a.w = ro_shrink(a.y, a.w);
                    c.z = ro_shrink(c.x, c.z);
                    c.w = ro_shrink(c.y, c.w);
                    box[i] = a;
                    box[j] = c;
                    changed = true;
                    pairs[w++] = pr;
                }
            }
            np = w;
            if (!changed) break;
        }
    }
    __syncthreads();

    // D1. avg_h (utils.py:581) and the vertical tolerance
    {
        long long hs = 0;
        for (int k = threadIdx.x; k < K; k += kRoThreads) hs += (long long)box[k].w - (long long)box[k].y;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) hs += __shfl_xor_sync(0xffffffffu, hs, off);
        if (threadIdx.x == 0) s_hsum = 0;
        __syncthreads();
        if ((threadIdx.x & 31) == 0) atomicAdd(reinterpret_cast<unsigned long long *>(&s_hsum), (unsigned long long)hs);
        __syncthreads();
        if (threadIdx.x == 0) {
            const double avg = K > 0 ? (double)s_hsum / (double)K : 0.0;
            s_ytol = avg * 0.6;
            s_avg_pos = avg > 0.0 ? 1 : 0;  // (b[0] - last_x1) <= avg_h * inf: true iff avg_h > 0 (0 * inf is NaN)
            s_lines = 0;
        }
    }
    __syncthreads();

    // D2. stable order by centre y (utils.py:584): key = y0 + y1 (cy = key / 2), ties by index
    int n2 = 1;
    while (n2 < K) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += kRoThreads) {
        uint64_t key = ~0ull;
        if (i < K) {
            const long long s2 = (long long)box[i].y + (long long)box[i].w;
            key = ((uint64_t)(s2 + (1ll << 33)) << 12) | (uint64_t)i;
        }
        keys[i] = key;
    }
    __syncthreads();
    ro_bitonic(keys, n2);

    // D3. sequential line assignment (utils.py:584-603) by warp 0; line l keeps the exact integer sum of y0 + y1
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const double ytol = s_ytol;
        const bool xgap_ok = s_avg_pos != 0;
        int L = 0;
        for (int r = 0; r < K; r++) {
            const int k = (int)(keys[r] & 0xfffu);
            const int4 b = box[k];
            const long long s2 = (long long)b.y + (long long)b.w;
            const double cy = (double)s2 / 2.0;
            int found = -1;
            if (xgap_ok) {
                for (int l0 = 0; l0 < L && found < 0; l0 += 32) {
                    const int l = l0 + lane;
                    bool ok = false;
                    if (l < L) ok = fabs(cy - line_cy[l]) <= ytol;
                    const uint32_t m = __ballot_sync(0xffffffffu, ok);
                    if (m) found = l0 + __ffs(m) - 1;
                }
            }
            if (lane == 0) {
                if (found >= 0) {
                    line_sum[found] += s2;
                    line_cnt[found] += 1;
                } else {
                    found = L;
                    line_sum[L] = s2;
                    line_cnt[L] = 1;
                }
                // np.mean of the members' centres: exact sum (multiples of 0.5), one division
                line_cy[found] = ((double)line_sum[found] * 0.5) / (double)line_cnt[found];
                line_of[k] = (uint16_t)found;
                seq_of[k] = (uint16_t)r;
            }
            found = __shfl_sync(0xffffffffu, found, 0);
            if (found == L) L++;
            __syncwarp();
        }
        if (lane == 0) s_lines = L;
    }
    __syncthreads();

    // D4. lines ordered by mean centre (utils.py:605, stable)
    {
        const int L = s_lines;
        for (int l = threadIdx.x; l < L; l += kRoThreads) {
            const uint64_t key = ro_orderable(line_cy[l]);
            int rank = 0;
            for (int m = 0; m < L; m++) {
                const uint64_t km = ro_orderable(line_cy[m]);
                rank += (km < key || (km == key && m < l)) ? 1 : 0;
            }
            line_rank[l] = (uint16_t)rank;
        }
    }
    __syncthreads();

    // D5. boxes by (line, x0, insertion order) (utils.py:606-609)
    {
        uint64_t mine[kRoMaxBoxes / kRoThreads];
        int cnt = 0;
        for (int i = threadIdx.x; i < n2; i += kRoThreads, cnt++) {
            uint64_t key = ~0ull;
            if (i < K) {
                const uint64_t lr = (uint64_t)line_rank[line_of[i]];
                const uint64_t x0 = (uint64_t)((long long)box[i].x + (1ll << 31));
                key = (lr << 44) | (x0 << 12) | (uint64_t)seq_of[i];
            }
            mine[cnt] = key;
        }
        __syncthreads();  // every thread has read line_rank / line_of / seq_of (they do not alias keys, but keep order)
        cnt = 0;
        for (int i = threadIdx.x; i < n2; i += kRoThreads, cnt++) keys[i] = mine[cnt];
    }
    __syncthreads();
    ro_bitonic(keys, n2);
    // position r holds the box with insertion order seq: map seq -> box index through the first sort's result.
    // seq_of[] is indexed by box; invert it into line_cnt[] (free now)
    for (int k = threadIdx.x; k < K; k += kRoThreads) {
        line_cnt[seq_of[k]] = k;
        obox_s[k] = obox_g[pb + k];  // line_sum / line_cy are dead
    }
    __syncthreads();

    // E. utils.py:639 (dict: the LAST box with equal compressed coordinates wins) and _pipeline.py:113-123 (the FIRST
    //    word with the same integer box is taken)
    for (int r = threadIdx.x; r < K; r += kRoThreads) {
        const int k = line_cnt[(int)(keys[r] & 0xfffu)];
        const int4 ck = box[k];
        int last = k;
        for (int m = K - 1; m > k; m--) {
            const int4 cm = box[m];
            if (cm.x == ck.x && cm.y == ck.y && cm.z == ck.z && cm.w == ck.w) {
                last = m;
                break;
            }
        }
        const int4 ob = obox_s[last];
        int first = last;
        for (int w = 0; w < last; w++) {
            const int4 ow = obox_s[w];
            if (ow.x == ob.x && ow.y == ob.y && ow.z == ob.z && ow.w == ob.w) {
                first = w;
                break;
            }
        }
        order[pb + r] = first;
        if (reordered) {
            const float *src = boxes8 + (pb + first) * row_stride;
            float *dst = reordered + (pb + r) * row_stride;
            for (int c = 0; c < row_stride; c++) dst[c] = src[c];
        }
    }
}

}  // namespace

size_t msk_reading_order_scratch(int n_pages, int cap_per_page)
{
    return (size_t)n_pages * cap_per_page * sizeof(int4) + 1024;
}

// order (n_pages*cap) int32: order[p*cap + r] = index of the word at reading position r; `reordered` (may be NULL)
// receives the rows of boxes8 in that order (it must not alias boxes8).
int msk_reading_order(ms_ctx *ctx, const float *boxes8, int row_stride, const int32_t *counts, int n_pages,
                      int cap_per_page, int32_t *order, float *reordered, int32_t *flags, ms_bump bump, cudaStream_t st)
{
    if (n_pages <= 0) return MS_OK;
    int4 *obox = bump.take<int4>((size_t)n_pages * cap_per_page);
    if (!obox) {
        ms_set_error("reading_order: scratch too small");
        return MS_ERR_CAPACITY;
    }
    const size_t smem = (size_t)kRoMaxBoxes * 16 + (size_t)kRoMaxPairs * 5 + (size_t)kRoMaxBoxes * 2;
    static_assert(kRoMaxPairs * 5 + kRoMaxBoxes * 2 >= 136 * 1024 && kRoMaxBoxes == 4096,
                  "pair + level region must hold the sort / line arrays laid out in the kernel");
    if ((int)smem > ctx->smem_attr[3]) {  // a synchronous driver call: once per context
        MS_CUDA(cudaFuncSetAttribute(reading_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->smem_attr[3] = (int)smem;
    }
    reading_order_kernel<<<n_pages, kRoThreads, smem, st>>>(boxes8, row_stride, counts, cap_per_page, obox, order,
                                                           reordered, flags);
    MS_LAUNCH_CHECK(ctx);
    return MS_OK;
}
