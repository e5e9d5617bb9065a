// Stable LSD radix sort of per-page segments of (u32 key, u32 value) pairs whose lengths live on the device.
//
// Used for the one ordering the reference's algorithm depends on: locality_aware_nms sorts the
// candidates of a page by x0 (lanms.py:166-168).  Keys are orderable(x0); the sort is stable and the
// input is in original-index order, which realises the documented tie rule (numpy kind="stable").
// One call sorts every page of a batch; grids are sized for the per-page capacity and read the true
// segment bounds from device memory, so no host round trip is needed.
#include "ms_internal.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItemsPerWarp = 256;                  // each warp owns a contiguous strip of the tile
constexpr int kTile = kWarps * kItemsPerWarp;       // 2048 keys per CTA: a 15k-candidate page is 8 CTAs
constexpr int kRadix = 256;

// Every page's (u32 key, u32 value) segment [page_off[p], page_off[p+1]) is sorted on its own.
// Pages hold a few tiles each, so a CTA derives its tile's digit bases directly from the page's tile histograms
// (no scan kernel) and a 32-bit key needs 4 passes of 2 launches.
// seg_len == NULL: segment p is [page_off[p], page_off[p+1]); otherwise it starts at page_off[p] and holds seg_len[p]
// items.  Segments of at most skip_le items are left alone (the caller orders those another way).
__global__ void __launch_bounds__(kThreads) seg_hist_kernel(const uint32_t *__restrict__ keys,
                                                            const int32_t *__restrict__ page_off,
                                                            const int32_t *__restrict__ seg_len, int skip_le,
                                                            int tiles_max, int shift, int32_t *__restrict__ hist)
{
    ms_pdl_wait();
    const int page = blockIdx.y;
    const int p0 = page_off[page], n = seg_len ? seg_len[page] : page_off[page + 1] - p0;
    if (n <= skip_le) return;
    __shared__ int s_h[kRadix];
    // the segment lengths live on the device: a few CTAs per page stride over its tiles
    for (int tile = blockIdx.x; tile * kTile < n; tile += gridDim.x) {
        for (int d = threadIdx.x; d < kRadix; d += kThreads) s_h[d] = 0;
        __syncthreads();
        const int base = tile * kTile;
        // all of the thread's keys are requested before the first one is counted (one round trip, not eight)
        uint32_t kk[kTile / kThreads];
#pragma unroll
        for (int u = 0; u < kTile / kThreads; u++) {
            const int idx = base + u * kThreads + threadIdx.x;
            kk[u] = idx < n ? keys[p0 + idx] : 0u;
        }
#pragma unroll
        for (int u = 0; u < kTile / kThreads; u++)
            if (base + u * kThreads + threadIdx.x < n) atomicAdd(&s_h[(kk[u] >> shift) & 0xFF], 1);
        __syncthreads();
        int32_t *h = hist + ((size_t)page * tiles_max + tile) * kRadix;
        for (int d = threadIdx.x; d < kRadix; d += kThreads) h[d] = s_h[d];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads) seg_scatter_kernel(const uint32_t *__restrict__ keys,
                                                               const uint32_t *__restrict__ vals,
                                                               const int32_t *__restrict__ page_off,
                                                               const int32_t *__restrict__ seg_len, int skip_le,
                                                               int tiles_max, int shift,
                                                               const int32_t *__restrict__ hist,
                                                               uint32_t *__restrict__ keys_out,
                                                               uint32_t *__restrict__ vals_out)
{
    ms_pdl_wait();
    const int page = blockIdx.y;
    const int p0 = page_off[page], n = seg_len ? seg_len[page] : page_off[page + 1] - p0;
    if (n <= skip_le) return;
    const int tiles = (n + kTile - 1) / kTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int s_cnt[kWarps][kRadix];  // per-warp digit counts, then running write cursors
    __shared__ int s_base[kRadix];         // first output slot of (digit, this tile) inside the page
    __shared__ int s_wsum[kWarps];
    constexpr int kPerLane = kItemsPerWarp / 32;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    // this lane's keys and values of the warp's strip: requested once, up front, together with the histogram reads
    // below (the count and the ordered walk both use them; three dependent round trips to L2 become one)
    uint32_t kk[kPerLane], vv[kPerLane];
    {
        const int wb = tile * kTile + warp * kItemsPerWarp;
#pragma unroll
        for (int u = 0; u < kPerLane; u++) {
            const int idx = wb + u * 32 + lane;
            kk[u] = idx < n ? keys[p0 + idx] : 0u;
            vv[u] = idx < n ? vals[p0 + idx] : 0u;
        }
    }
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
#pragma unroll
    for (int u = 0; u < kPerLane; u++)
        if (wbase + u * 32 + lane < n) atomicAdd(&s_cnt[warp][(kk[u] >> shift) & 0xFF], 1);
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
#pragma unroll
    for (int u = 0; u < kPerLane; u++) {
        const int idx = wbase + u * 32 + lane;
        const bool live = idx < n;
        if (!__any_sync(0xffffffffu, live)) break;
        const uint32_t k = kk[u];
        const uint32_t v = vv[u];
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
    __syncthreads();  // the shared tables are rebuilt for the next tile
    }
}

}  // namespace

size_t msk_sort_pages_scratch(int n_pages, int cap_per_page)
{
    const int tiles_max = (cap_per_page + kTile - 1) / kTile;
    return (size_t)n_pages * tiles_max * kRadix * sizeof(int32_t) + 1024;
}

// Stable sort of every page's (u32 key, u32 value) segment; the result ends in (keys, vals), tmp is clobbered.
int msk_sort_pages(ms_ctx *ctx, uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp,
                   const int32_t *page_off, const int32_t *seg_len, int skip_le, int n_pages, int cap_per_page,
                   ms_bump bump, cudaStream_t st)
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
    // segment lengths are device data: launch at most a few CTAs per page and let them stride over the page's tiles
    int gx = tiles_max;
    const int want = (ctx->num_sms * 8 + n_pages - 1) / n_pages;
    if (gx > want) gx = want;
    if (gx < 1) gx = 1;
    const dim3 grid(gx, n_pages);
    for (int pass = 0; pass < 4; pass++) {
        ms_launch(seg_hist_kernel, grid, kThreads, 0, st, kin, page_off, seg_len, skip_le, tiles_max, 8 * pass, hist);
        MS_LAUNCH_CHECK(ctx);
        ms_launch(seg_scatter_kernel, grid, kThreads, 0, st, kin, vin, page_off, seg_len, skip_le, tiles_max, 8 * pass, hist, kout,
                                                      vout);
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
