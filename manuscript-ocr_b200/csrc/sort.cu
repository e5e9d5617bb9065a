// Stable LSD radix sort of (u64 key, u32 value) pairs whose length lives on the device.
//
// Used for the one ordering the reference's algorithm depends on: locality_aware_nms sorts the
// candidates of a page by x0 (lanms.py:166-168).  Keys are (page << 32 | orderable(x0)); the sort is
// stable and the input is in original-index order, which realises the documented tie rule
// (numpy kind="stable").  One sort covers every page of a batch.
//
// Per 8-bit pass: tile histograms -> exclusive scan (digit-major) -> stable scatter.  Grids are sized
// for n_max and read the true n from device memory, so no host round trip is needed.
#include "ms_internal.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItemsPerWarp = 1024;                 // each warp owns a contiguous strip of the tile
constexpr int kTile = kWarps * kItemsPerWarp;       // 8192 keys per CTA
constexpr int kRadix = 256;

__device__ __forceinline__ int digit_of(uint64_t k, int shift) { return (int)((k >> shift) & 0xFF); }

// tile histograms, digit-major: hist[d * tiles + t] with tiles = ceil(n / kTile) (true n)
__global__ void __launch_bounds__(kThreads) rs_hist_kernel(const uint64_t *__restrict__ keys,
                                                           const int32_t *__restrict__ n_dev, int shift,
                                                           int32_t *__restrict__ hist)
{
    const int n = *n_dev;
    const int tiles = (n + kTile - 1) / kTile;
    __shared__ int s_h[kRadix];
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int d = threadIdx.x; d < kRadix; d += kThreads) s_h[d] = 0;
        __syncthreads();
        int base = tile * kTile;
        for (int i = threadIdx.x; i < kTile; i += kThreads) {
            int idx = base + i;
            if (idx < n) atomicAdd(&s_h[digit_of(keys[idx], shift)], 1);
        }
        __syncthreads();
        for (int d = threadIdx.x; d < kRadix; d += kThreads) hist[(size_t)d * tiles + tile] = s_h[d];
        __syncthreads();
    }
}

// exclusive scan over the flattened (digit, tile) histogram -- one CTA, each thread owns a
// contiguous chunk (serial), chunk sums are scanned across the CTA
__global__ void __launch_bounds__(1024) rs_scan_kernel(int32_t *__restrict__ hist,
                                                       const int32_t *__restrict__ n_dev)
{
    const int n = *n_dev;
    const int tiles = (n + kTile - 1) / kTile;
    const int total = tiles * kRadix;
    const int per = (total + 1023) / 1024;
    __shared__ int s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int lo = threadIdx.x * per, hi = lo + per;
    if (lo > total) lo = total;
    if (hi > total) hi = total;
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += hist[i];
    int inc = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
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
    }
    __syncthreads();
    int run = s_warp[warp] + inc - sum;
    for (int i = lo; i < hi; i++) {
        int v = hist[i];
        hist[i] = run;
        run += v;
    }
}

// stable scatter: each warp walks its strip of the tile in order, 32 keys at a time
__global__ void __launch_bounds__(kThreads) rs_scatter_kernel(const uint64_t *__restrict__ keys,
                                                              const uint32_t *__restrict__ vals,
                                                              const int32_t *__restrict__ n_dev, int shift,
                                                              const int32_t *__restrict__ hist,
                                                              uint64_t *__restrict__ keys_out,
                                                              uint32_t *__restrict__ vals_out)
{
    const int n = *n_dev;
    const int tiles = (n + kTile - 1) / kTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int s_cnt[kWarps][kRadix];  // per-warp digit counts, then running write cursors
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int i = threadIdx.x; i < kWarps * kRadix; i += kThreads) (&s_cnt[0][0])[i] = 0;
        __syncthreads();
        const int wbase = tile * kTile + warp * kItemsPerWarp;
        // pass 1: per-warp digit counts
        for (int i = lane; i < kItemsPerWarp; i += 32) {
            int idx = wbase + i;
            if (idx < n) atomicAdd(&s_cnt[warp][digit_of(keys[idx], shift)], 1);
        }
        __syncthreads();
        // cursors: global base of (digit, tile) + counts of earlier warps of this tile
        for (int d = threadIdx.x; d < kRadix; d += kThreads) {
            int run = hist[(size_t)d * tiles + tile];
#pragma unroll
            for (int w = 0; w < kWarps; w++) {
                int c = s_cnt[w][d];
                s_cnt[w][d] = run;
                run += c;
            }
        }
        __syncthreads();
        // pass 2: ordered walk; equal digits inside a 32-key group are ranked by lane
        for (int i0 = 0; i0 < kItemsPerWarp; i0 += 32) {
            int idx = wbase + i0 + lane;
            bool live = idx < n;
            if (!__any_sync(0xffffffffu, live)) break;
            uint64_t k = live ? keys[idx] : 0;
            uint32_t v = live ? vals[idx] : 0;
            int d = live ? digit_of(k, shift) : -1 - lane;  // dead lanes match nobody
            uint32_t peers = __match_any_sync(0xffffffffu, d);
            int rank = __popc(peers & ((1u << lane) - 1u));
            int leader = __ffs(peers) - 1;
            int pos = 0;
            if (live) pos = s_cnt[warp][d] + rank;
            __syncwarp();
            if (live && lane == leader) s_cnt[warp][d] += __popc(peers);
            __syncwarp();
            if (live) {
                keys_out[pos] = k;
                vals_out[pos] = v;
            }
        }
        __syncthreads();
    }
}


// ---- segmented variant: every page's (u32 key, u32 value) segment [page_off[p], page_off[p+1]) is sorted on its own.
// Pages hold a few tiles each, so a CTA derives its tile's digit bases directly from the page's tile histograms
// (no scan kernel) and a 32-bit key needs 4 passes of 2 launches.
__global__ void __launch_bounds__(kThreads) seg_hist_kernel(const uint32_t *__restrict__ keys,
                                                            const int32_t *__restrict__ page_off, int tiles_max, int shift,
                                                            int32_t *__restrict__ hist)
{
    const int page = blockIdx.y, tile = blockIdx.x;
    const int p0 = page_off[page], n = page_off[page + 1] - p0;
    if (tile * kTile >= n) return;
    __shared__ int s_h[kRadix];
    for (int d = threadIdx.x; d < kRadix; d += kThreads) s_h[d] = 0;
    __syncthreads();
    const int base = tile * kTile;
    for (int i = threadIdx.x; i < kTile; i += kThreads) {
        const int idx = base + i;
        if (idx < n) atomicAdd(&s_h[(keys[p0 + idx] >> shift) & 0xFF], 1);
    }
    __syncthreads();
    int32_t *h = hist + ((size_t)page * tiles_max + tile) * kRadix;
    for (int d = threadIdx.x; d < kRadix; d += kThreads) h[d] = s_h[d];
}

__global__ void __launch_bounds__(kThreads) seg_scatter_kernel(const uint32_t *__restrict__ keys,
                                                               const uint32_t *__restrict__ vals,
                                                               const int32_t *__restrict__ page_off, int tiles_max,
                                                               int shift, const int32_t *__restrict__ hist,
                                                               uint32_t *__restrict__ keys_out,
                                                               uint32_t *__restrict__ vals_out)
{
    const int page = blockIdx.y, tile = blockIdx.x;
    const int p0 = page_off[page], n = page_off[page + 1] - p0;
    if (tile * kTile >= n) return;
    const int tiles = (n + kTile - 1) / kTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int s_cnt[kWarps][kRadix];  // per-warp digit counts, then running write cursors
    __shared__ int s_base[kRadix];         // first output slot of (digit, this tile) inside the page
    __shared__ int s_wsum[kWarps];
    // digit d of this tile starts after every smaller digit of the whole page and digit d of the earlier tiles
    {
        const int d = threadIdx.x;  // kThreads == kRadix
        const int32_t *h = hist + (size_t)page * tiles_max * kRadix;
        int tot = 0, before = 0;
        for (int t = 0; t < tiles; t++) {
            const int v = h[(size_t)t * kRadix + d];
            tot += v;
            if (t < tile) before += v;
        }
        // exclusive scan of the per-digit totals over the CTA
        int inc = tot;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) s_wsum[warp] = inc;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < warp; w++) wbase += s_wsum[w];
        s_base[d] = wbase + inc - tot + before;
    }
    for (int i = threadIdx.x; i < kWarps * kRadix; i += kThreads) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const int wbase = tile * kTile + warp * kItemsPerWarp;
    for (int i = lane; i < kItemsPerWarp; i += 32) {
        const int idx = wbase + i;
        if (idx < n) atomicAdd(&s_cnt[warp][(keys[p0 + idx] >> shift) & 0xFF], 1);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < kRadix; d += kThreads) {
        int run = s_base[d];
#pragma unroll
        for (int w = 0; w < kWarps; w++) {
            int c = s_cnt[w][d];
            s_cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    // ordered walk; equal digits inside a 32-key group are ranked by lane
    for (int i0 = 0; i0 < kItemsPerWarp; i0 += 32) {
        const int idx = wbase + i0 + lane;
        const bool live = idx < n;
        if (!__any_sync(0xffffffffu, live)) break;
        const uint32_t k = live ? keys[p0 + idx] : 0;
        const uint32_t v = live ? vals[p0 + idx] : 0;
        const int d = live ? (int)((k >> shift) & 0xFF) : -1 - lane;  // dead lanes match nobody
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const int leader = __ffs(peers) - 1;
        int pos = 0;
        if (live) pos = s_cnt[warp][d] + rank;
        __syncwarp();
        if (live && lane == leader) s_cnt[warp][d] += __popc(peers);
        __syncwarp();
        if (live) {
            keys_out[p0 + pos] = k;
            vals_out[p0 + pos] = v;
        }
    }
}

}  // namespace

static int sort_tiles_max(int64_t n_max) { return (int)((n_max + kTile - 1) / kTile); }

size_t msk_sort_scratch(int64_t n_max)
{
    return (size_t)sort_tiles_max(n_max) * kRadix * sizeof(int32_t) + 1024;
}

// Sorts in place from the caller's point of view: the result ends in (keys, vals); the tmp arrays
// are clobbered.  Bits [0, end_bit) are sorted, in ceil(end_bit/8) passes (odd pass counts copy back).
int msk_sort_pairs(ms_ctx *ctx, uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp,
                   const int32_t *n_dev, int64_t n_max, int end_bit, ms_bump bump, cudaStream_t st)
{
    if (n_max <= 0) return MS_OK;
    const int tiles_max = sort_tiles_max(n_max);
    int32_t *hist = bump.take<int32_t>((size_t)tiles_max * kRadix);
    if (!hist) {
        ms_set_error("sort: scratch too small");
        return MS_ERR_CAPACITY;
    }
    int passes = (end_bit + 7) / 8;
    if (passes & 1) passes++;  // even number of passes leaves the result in (keys, vals)
    const int grid = tiles_max < ctx->num_sms * 4 ? tiles_max : ctx->num_sms * 4;
    uint64_t *kin = keys, *kout = keys_tmp;
    uint32_t *vin = vals, *vout = vals_tmp;
    for (int p = 0; p < passes; p++) {
        int shift = 8 * p;
        rs_hist_kernel<<<grid, kThreads, 0, st>>>(kin, n_dev, shift, hist);
        MS_LAUNCH_CHECK(ctx);
        rs_scan_kernel<<<1, 1024, 0, st>>>(hist, n_dev);
        MS_LAUNCH_CHECK(ctx);
        rs_scatter_kernel<<<grid, kThreads, 0, st>>>(kin, vin, n_dev, shift, hist, kout, vout);
        MS_LAUNCH_CHECK(ctx);
        uint64_t *tk = kin;
        kin = kout;
        kout = tk;
        uint32_t *tv = vin;
        vin = vout;
        vout = tv;
    }
    return MS_OK;
}

size_t msk_sort_pages_scratch(int n_pages, int cap_per_page)
{
    const int tiles_max = (cap_per_page + kTile - 1) / kTile;
    return (size_t)n_pages * tiles_max * kRadix * sizeof(int32_t) + 1024;
}

// Stable sort of every page's (u32 key, u32 value) segment; the result ends in (keys, vals), tmp is clobbered.
int msk_sort_pages(ms_ctx *ctx, uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp,
                   const int32_t *page_off, int n_pages, int cap_per_page, ms_bump bump, cudaStream_t st)
{
    if (n_pages <= 0 || cap_per_page <= 0) return MS_OK;
    static_assert(kThreads == kRadix, "seg_scatter_kernel maps one thread per digit");
    const int tiles_max = (cap_per_page + kTile - 1) / kTile;
    int32_t *hist = bump.take<int32_t>((size_t)n_pages * tiles_max * kRadix);
    if (!hist) {
        ms_set_error("sort: scratch too small");
        return MS_ERR_CAPACITY;
    }
    uint32_t *kin = keys, *kout = keys_tmp, *vin = vals, *vout = vals_tmp;
    const dim3 grid(tiles_max, n_pages);
    for (int pass = 0; pass < 4; pass++) {
        seg_hist_kernel<<<grid, kThreads, 0, st>>>(kin, page_off, tiles_max, 8 * pass, hist);
        MS_LAUNCH_CHECK(ctx);
        seg_scatter_kernel<<<grid, kThreads, 0, st>>>(kin, vin, page_off, tiles_max, 8 * pass, hist, kout, vout);
        MS_LAUNCH_CHECK(ctx);
        uint32_t *t = kin;
        kin = kout;
        kout = t;
        t = vin;
        vin = vout;
        vout = t;
    }
    return MS_OK;
}
