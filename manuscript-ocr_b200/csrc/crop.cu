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
// HBM roofline: per crop 3*w*h source bytes read + 3*ih*iw*4 bytes written (49 152 B at 32x128): the
// kernel is write dominated.
//
// Kernels:
//   crop_plan_kernel        thread per crop: the float64 sizing arithmetic of transforms.py:91-98, interpolation mode,
//                           staging decision -> Plan
//   crop_resize_pad_kernel  the crops with Plan::fast (INTER_AREA with shrink factors below 3 -- what shrinking a
//                           word box to a 32- or 64-pixel-high canvas gives): persistent, warp specialised, 4 CTAs
//                           of 7 warps per SM.  Warp 0 (copy warp) stages the next crop's source rows with TMA bulk
//                           copies (cp.async.bulk global->shared, one 16-byte-aligned span per row, completion on an
//                           mbarrier) into a double-buffered stage; warp 1 (table warp) builds OpenCV's coefficient
//                           tables in shared memory (structure of arrays) and writes the 255-padding as constant
//                           16-byte streaming stores; warps 2..6 (consumers) resample column strips out of shared
//                           memory (aligned 32-bit loads + funnel shift + PRMT u8->f32, 3 or 4 taps) and store
//                           normalised pixels straight into the CHW batch.  full / empty mbarriers, no CTA-wide
//                           barrier in the loop.
//   crop_generic_kernel     every other crop (copy, INTER_LINEAR upscale, integer-ratio or long-table INTER_AREA,
//                           rows too long for the stage): one CTA per crop from global memory, same arithmetic.
#include "crop_common.cuh"

namespace {

// CTA shape of the persistent kernel.  MS_CROP_PAD_WARP = 0: copy warp + table/padding warp + 5 consumer warps, 4 CTAs
// per SM (224 threads, 72 registers).  MS_CROP_PAD_WARP = 1: copy warp + table warp + PADDING warp + 7 consumer warps,
// 3 CTAs per SM (320 threads, 64 registers): the padding no longer serialises behind the tables in one warp.
#ifndef MS_CROP_PAD_WARP
#define MS_CROP_PAD_WARP 0
#endif
#if MS_CROP_PAD_WARP
constexpr int kThreads = 320;
constexpr int kCtasPerSm = 3;
constexpr int kSlots = 4;       // crops in flight per CTA (ring slots: metadata + tables per slot, bytes from the ring)
#else
constexpr int kThreads = 224;
constexpr int kCtasPerSm = 4;
#ifndef MS_CROP_SLOTS
#define MS_CROP_SLOTS 3
#endif
constexpr int kSlots = MS_CROP_SLOTS;  // (4 slots: 3.5 KB less ring per CTA -- measured r2p, see profiles/README.md)
#endif
constexpr bool kPadWarp = MS_CROP_PAD_WARP != 0;
constexpr int kSrcBuf = 40 * 1024;  // largest staged crop (bytes); the staging ring of a CTA holds at least one
constexpr int kBandsY = 64, kCellsX = 8;  // work-list buckets per page: (row band, x cell)
#ifndef MS_COPY_PAD_CHANNELS
#define MS_COPY_PAD_CHANNELS 0
#endif
constexpr int kCopyPadChannels = MS_COPY_PAD_CHANNELS;  // float32 padding channels written by the copy warp (A/B switch)
#ifndef MS_PAD_LAG
#define MS_PAD_LAG 0
#endif
constexpr int kPadLag = MS_PAD_LAG;  // crops between building a crop's tables and writing its padding (A/B switch)

// Pages are either one (n_pages, img_h, img_w, 3) tensor, or -- page_ptrs != NULL -- separate images of their own
// sizes: page_ptrs[p] -> (page_hw[2p], page_hw[2p+1], 3) bytes.
__device__ __forceinline__ void make_plan(const int32_t *cr, int n_pages, int img_h, int img_w, int ih, int iw,
                                          const uint8_t *pages, const uint8_t *const *page_ptrs,
                                          const int32_t *page_hw, Plan &p)
{
    p.page = cr[0];
    p.x1 = cr[1];
    p.y1 = cr[2];
    p.w = cr[3] - cr[1];
    p.h = cr[4] - cr[2];
    p.staged = 0;
    p.fast = 0;
    p.pitch = 0;
    p.stride = 0;
    p.src = nullptr;
    p.nw = p.nh = p.y0 = p.interp = p.isx = p.isy = 0;
    p.scale_x = p.scale_y = 1.0;
    p.ok = p.page >= 0 && p.page < n_pages;
    if (!p.ok) return;
    const int H = page_ptrs ? page_hw[2 * p.page] : img_h, W = page_ptrs ? page_hw[2 * p.page + 1] : img_w;
    const uint8_t *page0 = page_ptrs ? page_ptrs[p.page] : pages + (size_t)p.page * img_h * (size_t)img_w * 3;
    p.ok = page0 != nullptr && H > 0 && W > 0 && p.w > 0 && p.h > 0 && p.x1 >= 0 && p.y1 >= 0 && cr[3] <= W && cr[4] <= H;
    if (!p.ok) return;
    const int w = p.w, h = p.h;
    plan_resize(w, h, ih, iw, p);
    // staging: every row is copied as the 16-byte-aligned span that covers it; the pitch is exactly that span (the
    // 4-tap reads may run up to 16 bytes past a row -- into the next row, or into the slack kept after the last one).
    // When the page stride is a multiple of 16 every row has the same misalignment, otherwise assume the worst (15).
    const size_t stride = (size_t)W * 3;
    p.stride = (int)stride;
    p.src = page0 + (size_t)p.y1 * stride + (size_t)p.x1 * 3;
    const uintptr_t first = reinterpret_cast<uintptr_t>(p.src);
    const uintptr_t last_end = first + (size_t)(h - 1) * stride + (size_t)w * 3;
    // the copies may not leave the memory that holds the pages: the whole tensor, or this page's own image
    const uintptr_t lo = reinterpret_cast<uintptr_t>(page_ptrs ? page0 : pages);
    const uintptr_t hi = page_ptrs ? lo + (size_t)H * stride : lo + (size_t)n_pages * img_h * (size_t)img_w * 3;
    const int mis = (stride & 15) == 0 ? (int)(first & 15) : 15;
    const int pitch = (mis + w * 3 + 15) & ~15;
    const bool fits = (size_t)pitch * h + 16 <= (size_t)kSrcBuf;
    const bool tail_ok = ((last_end + 15) & ~(uintptr_t)15) <= hi;
    if (fits && tail_ok && (lo & 15) == 0) {
        p.staged = 1;
        p.pitch = pitch;
    }
    // a decimation window shorter than 3 source pixels touches at most 4 of them: every table entry has <= 4 taps
    // ... and the exact 2 x 2 integer ratio has its own staged routine
    p.fast = p.staged && p.w < 65536 && p.h < 65536 &&
             ((p.interp == 3 && p.scale_x < 2.999 && p.scale_y < 2.999) || (p.interp == 2 && p.isx == 2 && p.isy == 2));
}

// plans for all crops (thread per crop): the float64 sizing arithmetic of transforms.py:91-98 runs here, off the
// critical path of the persistent resampling kernel.  Fast crops are also counted into their work-list bucket
// (page, row band, x cell): the persistent kernel walks the crops in bucket order, so that overlapping word boxes are
// resampled within microseconds of each other and their shared source rows come out of L2 instead of DRAM.
__device__ __forceinline__ int crop_bucket(const Plan &p, const int32_t *page_hw, int img_h, int img_w, int n_pages)
{
    const int H = page_hw ? page_hw[2 * p.page] : img_h, W = page_hw ? page_hw[2 * p.page + 1] : img_w;
    const int by = min(kBandsY - 1, (int)(((int64_t)p.y1 * kBandsY) / max(H, 1)));
    const int bx = min(kCellsX - 1, (int)(((int64_t)p.x1 * kCellsX) / max(W, 1)));
    return (min(p.page, n_pages - 1) * kBandsY + by) * kCellsX + bx;
}

__global__ void __launch_bounds__(256) crop_plan_kernel(const uint8_t *__restrict__ pages,
                                                        const uint8_t *const *__restrict__ page_ptrs,
                                                        const int32_t *__restrict__ page_hw, int n_pages, int img_h,
                                                        int img_w, const int32_t *__restrict__ crops,
                                                        const int32_t *__restrict__ n_crops_dev,
                                                        const int32_t *__restrict__ range, int64_t crops_cap, int ih,
                                                        int iw, Plan *__restrict__ plans, int32_t *__restrict__ hist)
{
    ms_pdl_wait();
    int64_t begin = range ? range[0] : 0;
    int64_t n_crops = range ? range[1] : *n_crops_dev;
    if (n_crops > crops_cap) n_crops = crops_cap;
    for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_crops;
         i += (int64_t)gridDim.x * blockDim.x) {
        Plan p;
        make_plan(crops + i * 5, n_pages, img_h, img_w, ih, iw, pages, page_ptrs, page_hw, p);
        plans[i] = p;
        if (p.fast) atomicAdd(&hist[crop_bucket(p, page_ptrs ? page_hw : nullptr, img_h, img_w, n_pages)], 1);
    }
}

// exclusive scan of the bucket counts (one CTA); hist becomes the buckets' write cursors, *n_work the number of fast
// crops.  A thread owns 32 consecutive counts, fetched with eight independent 16-byte loads (the first version walked
// them one dependent load at a time: 57 us for 32 768 buckets); tiles of 32 768 counts are chained by a carry.
__global__ void __launch_bounds__(1024) crop_bucket_scan_kernel(int32_t *__restrict__ hist, int n_buckets,
                                                                int32_t *__restrict__ n_work)
{
    ms_pdl_wait();
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int tile = 0; tile < n_buckets; tile += 32 * 1024) {
        const int lo = tile + (int)threadIdx.x * 32;
        int v[32];
        if (lo + 32 <= n_buckets) {  // n_buckets is a multiple of 4 and hist is 256-byte aligned
            const int4 *q = reinterpret_cast<const int4 *>(hist + lo);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int4 x = q[i];
                v[4 * i] = x.x, v[4 * i + 1] = x.y, v[4 * i + 2] = x.z, v[4 * i + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = lo + i < n_buckets ? hist[lo + i] : 0;
        }
        int sum = 0;
#pragma unroll
        for (int i = 0; i < 32; i++) {
            const int c = v[i];
            v[i] = sum;
            sum += c;
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        const int carry = s_carry;
        if (warp == 0) {
            const int w = s_warp[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += u;
            }
            s_warp[lane] = wi - w;
            if (lane == 31) s_carry = carry + wi;
        }
        __syncthreads();
        const int base = carry + s_warp[warp] + incl - sum;
        if (lo + 32 <= n_buckets) {
            int4 *q = reinterpret_cast<int4 *>(hist + lo);
#pragma unroll
            for (int i = 0; i < 8; i++) q[i] = make_int4(base + v[4 * i], base + v[4 * i + 1], base + v[4 * i + 2], base + v[4 * i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++)
                if (lo + i < n_buckets) hist[lo + i] = base + v[i];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_work = s_carry;
}

// What the copy warp of the persistent kernel needs of a crop, in work-list order (32 bytes, one broadcast load).
struct __align__(16) CopyDesc {
    const uint8_t *src;  // first source byte of the crop
    int stride;          // bytes between source rows
    int pitch;           // bytes between staged rows
    int wh;              // crop size in pixels: w | h << 16 (both < 65536 for a fast crop)
    int nwh;             // pasted size: nw | nh << 16
    int ci;              // crop index (output slot)
    int y0;              // first canvas row of the pasted rectangle
};

// work[cursor of the crop's bucket ++] = the crop's copy descriptor (order inside a bucket is arbitrary: it only shapes
// the schedule)
__global__ void __launch_bounds__(256) crop_bucket_scatter_kernel(const Plan *__restrict__ plans,
                                                                  const int32_t *__restrict__ page_hw, int n_pages,
                                                                  int img_h, int img_w,
                                                                  const int32_t *__restrict__ n_crops_dev,
                                                                  const int32_t *__restrict__ range, int64_t crops_cap,
                                                                  int32_t *__restrict__ cursors, CopyDesc *__restrict__ work)
{
    ms_pdl_wait();
    int64_t begin = range ? range[0] : 0;
    int64_t n_crops = range ? range[1] : *n_crops_dev;
    if (n_crops > crops_cap) n_crops = crops_cap;
    for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_crops;
         i += (int64_t)gridDim.x * blockDim.x) {
        const Plan p = plans[i];
        if (p.fast) {
            CopyDesc d;
            d.src = p.src;
            d.stride = p.stride;
            d.pitch = p.pitch;
            d.wh = p.w | (p.h << 16);
            d.nwh = p.nw | (p.nh << 16);
            d.ci = (int)i;
            d.y0 = p.y0;
            work[atomicAdd(&cursors[crop_bucket(p, page_hw, img_h, img_w, n_pages)], 1)] = d;
        }
    }
}

// Warp-specialised persistent kernel for the crops with Plan::fast.  Two producer warps run ahead of the consumers:
// warp 0 (copy warp) takes the next crop of the work list (a global ticket counter: a CTA that drew narrow words simply
// draws more of them), reserves the crop's bytes in the CTA's staging RING, tells the table warp which crop it is and
// issues the TMA row copies; warp 1 (table warp) builds the axis tables, signals, and then writes the crop's padding
// (which nobody waits for).  Up to kSlots crops are in flight per CTA; a slot owns its metadata and tables, its source
// rows take exactly the bytes they need from the ring (a narrow word does not hold a 24 KB stage), so the producers
// can be several crops ahead of the consumers and the differences between crops average out.  full[s] completes when
// both producers have arrived and the copied bytes have landed.  Warps 2..6 (consumers) wait on full[s], resample the
// pasted rectangle straight into the CHW batch and release the slot (and its ring bytes) on empty[s].  No CTA-wide
// barrier inside the loop.  History (profiles/README.md): with two fixed 24 KB stages the consumers waited on `full`
// 24 % of their time and the copy warp on `empty` 33 % of its -- producer and consumer periods are equal on average,
// so a two-deep pipeline stalls on every fluctuation.
// Each consumer thread resamples a column strip (one destination column, G consecutive destination rows): with a
// shrink factor near 2 neighbouring destination rows share their boundary source row, so a strip needs ~(2G + 1)
// horizontal row sums instead of 3G; per-pixel arithmetic and its order are unchanged.
constexpr int kProducerWarps = kPadWarp ? 3 : 2;
constexpr int kConsumerWarps = kThreads / 32 - kProducerWarps;

template <bool kWriteF32, bool kWriteU8>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
    crop_resize_pad_kernel(const Plan *__restrict__ plans, const CopyDesc *__restrict__ work,
                           const int32_t *__restrict__ n_work_dev, int32_t *__restrict__ ticket, int ring_bytes, int ih,
                           int iw, float *__restrict__ batch, uint8_t *__restrict__ canvas_out, int vec_ok,
                           uint8_t *__restrict__ redo)
{
    ms_pdl_wait();
    extern __shared__ __align__(128) unsigned char smem[];  // [ring_bytes] staging ring, then kSlots table sets
    const int tab_n = area_tab_words(ih, iw);  // words of one slot's table set
    uint32_t *tabs = reinterpret_cast<uint32_t *>(smem + ring_bytes);  // [kSlots][tab_n]
    __shared__ __align__(8) uint64_t s_full[kSlots], s_empty[kSlots], s_tick[kSlots];
    // per slot: the crop (index, -1 = no more crops) and what the consumers need of its plan, so that they never touch
    // the plan array in global memory
    __shared__ int s_next[kSlots];    // copy warp -> table warp: crop index of the slot
    __shared__ int s_off[kSlots];     // copy warp -> consumers: ring offset of the crop's first row
    __shared__ int s_pad[kSlots][2];  // copy warp -> padding warp: nw | nh << 16, y0
    __shared__ long long s_ci[kSlots];
    __shared__ int s_meta[kSlots][8];  // nw, nh, y0, pitch, misalignment of row 0, its change per row, <= 3 x taps, strip height

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int plane = ih * iw;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kSlots; i++) {
            mbar_init(&s_full[i], 2);  // the copy warp (with the byte count) and the table warp
            mbar_init(&s_empty[i], kConsumerWarps + (kPadWarp ? 1 : 0));  // + the padding warp (it read the slot's descriptor)
            mbar_init(&s_tick[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ---------------- copy warp: tickets, ring space, one bulk copy per source row ----------------
        // Software pipeline: while crop k is being copied, the descriptor of crop k+1 and the ticket of crop k+2 are in
        // flight, so neither the atomic nor the descriptor load is on the warp's critical path.
        const int n_work = *n_work_dev;
        int q_off[kSlots], q_len[kSlots];  // ring intervals of the slots (len 0: free)
#pragma unroll
        for (int i = 0; i < kSlots; i++) q_off[i] = q_len[i] = 0;
        int head = 0;        // where the newest crop ended
        int oldest = 0;      // first crop (sequence number) whose slot has not been seen released
        auto load_desc = [&](int t) {
            CopyDesc d;
            if (t < n_work) {
                const uint4 *q = reinterpret_cast<const uint4 *>(work + t);
                const uint4 a = q[0], b = q[1];
                d.src = reinterpret_cast<const uint8_t *>(((uint64_t)a.y << 32) | a.x);
                d.stride = (int)a.z;
                d.pitch = (int)a.w;
                d.wh = (int)b.x;
                d.nwh = (int)b.y;
                d.ci = (int)b.z;
                d.y0 = (int)b.w;
            } else {
                d.src = nullptr;
                d.stride = d.pitch = d.wh = d.nwh = d.y0 = 0;
                d.ci = -1;
            }
            return d;
        };
        int t1 = 0;
        if (lane == 0) t1 = atomicAdd(ticket, 1);
        CopyDesc cur = load_desc(__shfl_sync(0xffffffffu, t1, 0));
        if (lane == 0) t1 = atomicAdd(ticket, 1);  // ticket of crop 1
        for (int k = 0;; k++) {
            const int s = k % kSlots;
            const CopyDesc nxt = load_desc(__shfl_sync(0xffffffffu, t1, 0));  // crop k+1: lands during this crop's copies
            if (lane == 0 && cur.ci >= 0) t1 = atomicAdd(ticket, 1);           // ticket of crop k+2
            // the slot's previous crop (k - kSlots) must be done before its metadata / tables are reused
            while (oldest + kSlots <= k) {
                mbar_wait(&s_empty[oldest % kSlots], (uint32_t)((oldest / kSlots) & 1));
                q_len[oldest % kSlots] = 0;
                oldest++;
            }
            if (cur.ci < 0) {  // no more crops: tell the table warp, which tells the consumers
                if (lane == 0) {
                    s_next[s] = -1;
                    mbar_arrive(&s_tick[s]);
                    mbar_arrive(&s_full[s]);
                }
                break;
            }
            const int cw = cur.wh & 0xffff, chh = (int)((uint32_t)cur.wh >> 16);
            const int need = (cur.pitch * chh + 16 + 127) & ~127;
            int off;
            for (;;) {  // first fit at `head`, else at 0; else wait for the oldest crop in flight
                auto free_at = [&](int c) {
                    bool ok = c + need <= ring_bytes;
#pragma unroll
                    for (int i = 0; i < kSlots; i++) ok = ok && (q_len[i] == 0 || c + need <= q_off[i] || q_off[i] + q_len[i] <= c);
                    return ok;
                };
                if (free_at(head)) {
                    off = head;
                    break;
                }
                if (free_at(0)) {
                    off = 0;
                    break;
                }
                // oldest < k here: with nothing in flight the whole ring is free and `need` fits by construction
                mbar_wait(&s_empty[oldest % kSlots], (uint32_t)((oldest / kSlots) & 1));
                q_len[oldest % kSlots] = 0;
                oldest++;
            }
            q_off[s] = off;
            q_len[s] = need;
            head = off + need;
            if (lane == 0) {
                s_next[s] = cur.ci;
                s_off[s] = off;
                s_pad[s][0] = cur.nwh;
                s_pad[s][1] = cur.y0;
                mbar_arrive(&s_tick[s]);  // release: the table warp and (through full) the consumers see both
            }
            uint32_t bytes = 0;
            {
                const uint8_t *src = cur.src;
                const size_t stride = (size_t)cur.stride;
                unsigned char *buf = smem + off;
                for (int r = lane; r < chh; r += 32) {
                    const uint8_t *g = src + (size_t)r * stride;
                    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 15);
                    const uint32_t sz = (a + (uint32_t)cw * 3 + 15) & ~15u;
#if !defined(MS_EXP_NO_COPY)
                    tma_bulk_g2s(buf + (size_t)r * cur.pitch, g - a, sz, &s_full[s]);
                    bytes += sz;
#endif
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
            }
            if (lane == 0) mbar_expect_tx(&s_full[s], bytes);
            // this warp's share of the crop's padding (kCopyPadChannels of the three channels): a lone warp's stores
            // proceed at one 128-byte line per ~25 cycles, so all the padding on the table warp made that warp the
            // kernel's critical path (profiles/README.md, r2d experiments)
            if (kCopyPadChannels > 0)
                write_padding<kWriteF32, kWriteU8>(cur.nwh & 0xffff, (int)((uint32_t)cur.nwh >> 16), cur.y0, ih, iw,
                                                   kWriteF32 ? batch + (size_t)cur.ci * 3 * plane : nullptr, nullptr, vec_ok,
                                                   lane, 32, 3 - kCopyPadChannels, 3);
            cur = nxt;
        }
        return;
    }
    if (warp == 1) {
        // ---------------- table warp: OpenCV's axis tables, then all of the padding ----------------
        // kPadLag > 0: the padding of a crop is written kPadLag crops later, i.e. about when the consumers write its
        // pixels, so that a canvas reaches DRAM in one visit instead of two
        int lag_ci[kPadLag + 1], lag_nwh[kPadLag + 1], lag_y0[kPadLag + 1];
#pragma unroll
        for (int i = 0; i <= kPadLag; i++) lag_ci[i] = -1, lag_nwh[i] = 0, lag_y0[i] = 0;
        auto pad_one = [&](int pci, int nwh, int py0) {
            if (pci >= 0)
                write_padding<kWriteF32, kWriteU8>(nwh & 0xffff, (int)((uint32_t)nwh >> 16), py0, ih, iw,
                                                   kWriteF32 ? batch + (size_t)pci * 3 * plane : nullptr,
                                                   kWriteU8 ? canvas_out + (size_t)pci * 3 * plane : nullptr, vec_ok, lane,
                                                   32, 0, kWriteF32 ? 3 - kCopyPadChannels : 3);
        };
        for (int k = 0;; k++) {
            const int s = k % kSlots;
            mbar_wait(&s_tick[s], (uint32_t)((k / kSlots) & 1));
            const int ci = s_next[s];
            if (ci < 0) {
                if (lane == 0) {
                    s_ci[s] = -1;
                    mbar_arrive(&s_full[s]);
                }
#if !defined(MS_EXP_NO_PAD)
#pragma unroll
                for (int i = 0; i < kPadLag; i++) pad_one(lag_ci[i], lag_nwh[i], lag_y0[i]);
#endif
                break;
            }
            const Plan p = plans[ci];
            const bool two = p.interp == 2;  // exact 2 x 2 ratio: no tables
            const bool x3 = two || __all_sync(0xffffffffu, build_tables(p, tabs + (size_t)s * tab_n, tab_n, iw, lane));
            if (lane == 0) {
                s_ci[s] = ci;
                s_meta[s][0] = p.nw;
                s_meta[s][1] = p.nh;
                s_meta[s][2] = p.y0;
                s_meta[s][3] = p.pitch;
                s_meta[s][4] = (int)(reinterpret_cast<uintptr_t>(p.src) & 15);
                s_meta[s][5] = p.stride & 15;  // per-row change of the 16-byte misalignment
                s_meta[s][6] = two ? 2 : (x3 ? 1 : 0);
                s_meta[s][7] = strip_height<kConsumerWarps * 32>(p.nw, p.nh);
            }
            __syncwarp();  // every lane's table stores are ordered before lane 0's releasing arrive
            if (lane == 0) mbar_arrive(&s_full[s]);
#if !defined(MS_EXP_NO_PAD)
            if (kPadWarp) continue;
            lag_ci[kPadLag] = ci, lag_nwh[kPadLag] = p.nw | (p.nh << 16), lag_y0[kPadLag] = p.y0;
            pad_one(lag_ci[0], lag_nwh[0], lag_y0[0]);
#pragma unroll
            for (int i = 0; i < kPadLag; i++) lag_ci[i] = lag_ci[i + 1], lag_nwh[i] = lag_nwh[i + 1], lag_y0[i] = lag_y0[i + 1];
#endif
        }
        return;
    }

    if (kPadWarp && warp == 2) {
        // ---------------- padding warp: the 255-padding of every crop, as soon as the copy warp has drawn it ----------------
        for (int k = 0;; k++) {
            const int s = k % kSlots;
            mbar_wait(&s_tick[s], (uint32_t)((k / kSlots) & 1));
            const int ci = s_next[s];
            if (ci < 0) break;
            const int nwh = s_pad[s][0], py0 = s_pad[s][1];
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);  // the slot's descriptor has been read
#if !defined(MS_EXP_NO_PAD)
            write_padding<kWriteF32, kWriteU8>(nwh & 0xffff, (int)((uint32_t)nwh >> 16), py0, ih, iw,
                                               kWriteF32 ? batch + (size_t)ci * 3 * plane : nullptr,
                                               kWriteU8 ? canvas_out + (size_t)ci * 3 * plane : nullptr, vec_ok, lane, 32, 0, 3);
#endif
        }
        return;
    }

    // ---------------- consumers ----------------
    const int ct = (warp - kProducerWarps) * 32 + lane;
    constexpr int kCT = kConsumerWarps * 32;
    for (int k = 0;; k++) {
        const int s = k % kSlots;
        mbar_wait(&s_full[s], (uint32_t)((k / kSlots) & 1));
        const int64_t ci = s_ci[s];
        if (ci < 0) break;
        const int nw = s_meta[s][0], nh = s_meta[s][1], y0 = s_meta[s][2];
        const uint32_t pitch = (uint32_t)s_meta[s][3], a0 = (uint32_t)s_meta[s][4], sstep = (uint32_t)s_meta[s][5];
        const int G = s_meta[s][7];
        const uint32_t stage_off = (uint32_t)s_off[s];
        float *dstf = kWriteF32 ? batch + (size_t)ci * 3 * plane : nullptr;
        uint8_t *dstu = kWriteU8 ? canvas_out + (size_t)ci * 3 * plane : nullptr;
        const uint32_t *tab = tabs + (size_t)s * tab_n;
        // three x taps at most (shrink factor below 2, most word boxes): a quarter of the horizontal work less; a page
        // stride that is a multiple of 16 bytes (sstep == 0): no per-row misalignment arithmetic
        bool bad = false;
#if defined(MS_EXP_NO_CONSUME)
        if (false)
#else
        if (s_meta[s][6] == 2)
            area2x2_pixels<kWriteF32, kWriteU8, kCT>(smem, stage_off, pitch, a0, sstep, ih, iw, nw, nh, y0, dstf, dstu, ct);
        else if (sstep == 0)
#endif
            bad = s_meta[s][6] == 1 ? area4_strips<kWriteF32, kWriteU8, kCT, 3, true>(smem, stage_off, pitch, a0, 0u, tab, tab_n, ih,
                                                                                 iw, nw, nh, y0, dstf, dstu, ct, G)
                               : area4_strips<kWriteF32, kWriteU8, kCT, 4, true>(smem, stage_off, pitch, a0, 0u, tab, tab_n, ih,
                                                                                 iw, nw, nh, y0, dstf, dstu, ct, G);
#if !defined(MS_EXP_NO_CONSUME)
        else
            bad = s_meta[s][6] == 1 ? area4_strips<kWriteF32, kWriteU8, kCT, 3>(smem, stage_off, pitch, a0, sstep, tab, tab_n, ih, iw,
                                                                           nw, nh, y0, dstf, dstu, ct, G)
                               : area4_strips<kWriteF32, kWriteU8, kCT, 4>(smem, stage_off, pitch, a0, sstep, tab, tab_n, ih, iw,
                                                                           nw, nh, y0, dstf, dstu, ct, G);
#endif
        if (bad) redo[ci] = 1;  // a table entry with more than 4 taps: the generic kernel redoes the crop
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s]);
    }
}

// Crops the persistent kernel does not take (same size copy, INTER_LINEAR upscale, integer-ratio INTER_AREA, shrink
// factors >= 3, rows too long for the stage, invalid rectangles, crops flagged `redo`) are listed first ...
__global__ void __launch_bounds__(256) crop_generic_list_kernel(const Plan *__restrict__ plans,
                                                                const int32_t *__restrict__ n_crops_dev,
                                                                const int32_t *__restrict__ range, int64_t crops_cap,
                                                                const uint8_t *__restrict__ redo,
                                                                int32_t *__restrict__ list, int32_t *__restrict__ list_n)
{
    ms_pdl_wait();
    const int64_t begin = range ? range[0] : 0;
    int64_t n_crops = range ? range[1] : *n_crops_dev;
    if (n_crops > crops_cap) n_crops = crops_cap;
    for (int64_t ci = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ci < n_crops;
         ci += (int64_t)gridDim.x * blockDim.x)
        if (!plans[ci].fast || redo[ci]) list[atomicAdd(list_n, 1)] = (int32_t)ci;
}

// ... and resampled one CTA per crop from global memory with the same arithmetic: the CTA builds the crop's
// coefficient tables in shared memory, then one thread per destination pixel.
constexpr int kGenericMaxTab = 1024;  // table entries kept in shared memory (longer axes recompute per pixel)

template <bool kWriteF32, bool kWriteU8>
__global__ void __launch_bounds__(256) crop_generic_kernel(const Plan *__restrict__ plans,
                                                           const int32_t *__restrict__ list,
                                                           const int32_t *__restrict__ list_n, int ih, int iw,
                                                           float *__restrict__ batch,
                                                           uint8_t *__restrict__ canvas_out, int vec_ok)
{
    ms_pdl_wait();
    __shared__ AxisEnt s_tab[kGenericMaxTab];
    const int plane = ih * iw;
    const float inv = 1.0f / 127.5f;
    const int todo = *list_n;
    for (int li = blockIdx.x; li < todo; li += gridDim.x) {
        const int64_t ci = list[li];
        const Plan p = plans[ci];
        float *dstf = kWriteF32 ? batch + (size_t)ci * 3 * plane : nullptr;
        uint8_t *dstu = kWriteU8 ? canvas_out + (size_t)ci * 3 * plane : nullptr;
        write_padding<kWriteF32, kWriteU8>(p, ih, iw, dstf, dstu, vec_ok, threadIdx.x, blockDim.x);
        const int nw = p.ok ? p.nw : 0, nh = p.ok ? p.nh : 0;
        const bool need_tab = p.ok && (p.interp == 1 || p.interp == 3);
        const bool tab_ok = need_tab && nw + nh <= kGenericMaxTab;
        __syncthreads();  // the previous crop's table is no longer read
        if (tab_ok) {
            for (int t = threadIdx.x; t < nw + nh; t += blockDim.x) {
                const bool isx = t < nw;
                const int d = isx ? t : t - nw;
                s_tab[t] = p.interp == 3 ? (isx ? area_entry(d, p.scale_x, p.w) : area_entry(d, p.scale_y, p.h))
                                         : (isx ? linear_entry_x(d, p.scale_x, p.w) : linear_entry_y(d, p.scale_y, p.h));
            }
        }
        __syncthreads();
        const PitchedSrc gsrc{p.src, (size_t)p.stride};
        const int npx = nw * nh;
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
            resample_px(p, dx, dy, gsrc, ex, ey, o0, o1, o2);
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
}

// ---- detector input (the caller side of the path, EAST.predict infer.py:301-305) ---------------------------------
// cv2.resize(img, (T, T)) -- INTER_LINEAR on uint8 (11-bit fixed-point coefficients, aspect NOT preserved), then
// torchvision ToTensor (x / 255) and Normalize(0.5, 0.5) ((x - 0.5) / 0.5), written as the (3, T, T) float32 network
// input.  Same arithmetic as resample_px's linear branch (the oracle's orc_resize_linear_u8c3, pinned to cv2 for
// shrinking and enlarging); a thread per destination pixel, the source page is read through L2.  This is what lets
// the page cross PCIe once: the original image is uploaded, the resized detector input is made on the device.
__global__ void __launch_bounds__(256) detector_input_kernel(const uint8_t *__restrict__ page, int H, int W, int th,
                                                             int tw, float *__restrict__ out_f32,
                                                             uint8_t *__restrict__ out_u8)
{
    ms_pdl_wait();
    const double scale_x = 1.0 / ((double)tw / (double)W), scale_y = 1.0 / ((double)th / (double)H);
    const bool same = (th == H && tw == W);  // cv2.resize returns a copy
    const size_t plane = (size_t)th * tw;
    const PitchedSrc src{page, (size_t)W * 3};
    Plan p;
    p.interp = same ? 0 : 1;
    p.isx = p.isy = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
        const int dy = (int)(i / tw), dx = (int)(i - (size_t)dy * tw);
        AxisEnt ex = {}, ey = {};
        if (!same) {
            ex = linear_entry_x(dx, scale_x, W);
            ey = linear_entry_y(dy, scale_y, H);
        }
        unsigned char o0, o1, o2;
        resample_px(p, dx, dy, src, ex, ey, o0, o1, o2);
        if (out_f32) {
            out_f32[i] = ((float)o0 / 255.0f - 0.5f) / 0.5f;
            out_f32[plane + i] = ((float)o1 / 255.0f - 0.5f) / 0.5f;
            out_f32[2 * plane + i] = ((float)o2 / 255.0f - 0.5f) / 0.5f;
        }
        if (out_u8) {
            out_u8[i * 3] = o0;
            out_u8[i * 3 + 1] = o1;
            out_u8[i * 3 + 2] = o2;
        }
    }
}

}  // namespace

int msk_detector_input(ms_ctx *ctx, const uint8_t *page, int img_h, int img_w, int target_h, int target_w,
                       float *out_f32, uint8_t *out_u8, cudaStream_t st)
{
    if (!page || img_h <= 0 || img_w <= 0 || target_h <= 0 || target_w <= 0 || (!out_f32 && !out_u8)) {
        ms_set_error("detector_input: bad arguments");
        return MS_ERR_INVALID;
    }
    const size_t n = (size_t)target_h * target_w;
    size_t grid = (n + 255) / 256;
    if (grid > (size_t)ctx->num_sms * 16) grid = (size_t)ctx->num_sms * 16;
    ms_launch(detector_input_kernel, (int)grid, 256, 0, st, page, img_h, img_w, target_h, target_w, out_f32, out_u8);
    MS_LAUNCH_CHECK(ctx);
    return MS_OK;
}

// scratch that depends on the page count as well: the work-list buckets
static size_t crop_bucket_count(int n_pages)
{
    size_t b = (size_t)(n_pages > 0 ? n_pages : 1) * kBandsY * kCellsX;
    return b > ((size_t)1 << 22) ? ((size_t)1 << 22) : b;
}

size_t msk_crop_scratch(int64_t crops_cap, int n_pages)
{
    const size_t n = (size_t)(crops_cap > 0 ? crops_cap : 0);
    return n * sizeof(Plan) + n + n * sizeof(int32_t) + n * sizeof(CopyDesc) + crop_bucket_count(n_pages) * sizeof(int32_t) + 8192;
}

int msk_crop(ms_ctx *ctx, const uint8_t *pages, int n_pages, int img_h, int img_w, const int32_t *crops,
             const int32_t *n_crops, const int32_t *range, int64_t crops_cap, int out_h, int out_w, float *batch_f32,
             uint8_t *canvas_u8, ms_bump bump, cudaStream_t st, const uint8_t *const *page_ptrs, const int32_t *page_hw)
{
    if (crops_cap <= 0 || n_pages <= 0) return MS_OK;
    const bool ragged = page_ptrs != nullptr && page_hw != nullptr;
    if (out_h <= 0 || out_w <= 0 || (!ragged && (img_h <= 0 || img_w <= 0 || !pages)) || (!batch_f32 && !canvas_u8)) {
        ms_set_error("crop: bad arguments");
        return MS_ERR_INVALID;
    }
    if ((int64_t)out_h * out_w > (1 << 20)) {
        ms_set_error("crop: canvas %dx%d too large", out_h, out_w);
        return MS_ERR_INVALID;
    }
    // shared memory of the persistent kernel: kSlots table sets + the staging ring, which gets what is left of an SM's
    // 227 KB split over the CTAs per SM (1 KB per CTA is reserved by the system, ~0.5 KB is static)
    const size_t tab_bytes = (size_t)kSlots * (size_t)area_tab_words(out_h, out_w) * sizeof(uint32_t);
    const size_t min_ring = (size_t)kSrcBuf + 128;
    int per_sm = kCtasPerSm;
    while (per_sm > 1 && (227 * 1024) / per_sm < (int)(tab_bytes + min_ring + 1024 + 512)) per_sm--;
    if ((227 * 1024) / per_sm < (int)(tab_bytes + min_ring + 1024 + 512)) {
        ms_set_error("crop: canvas %dx%d needs %zu bytes of shared memory", out_h, out_w, tab_bytes + min_ring);
        return MS_ERR_INVALID;
    }
    size_t ring = ((size_t)(227 * 1024) / per_sm - 1024 - 512 - tab_bytes) & ~(size_t)127;
    if (ring > 96 * 1024) ring = 96 * 1024;
    const size_t smem = ring + tab_bytes;
    const int n_buckets = (int)crop_bucket_count(n_pages);
    Plan *plans = bump.take<Plan>((size_t)crops_cap);
    uint8_t *redo = bump.take<uint8_t>((size_t)crops_cap);
    int32_t *glist = bump.take<int32_t>((size_t)crops_cap);
    CopyDesc *work = bump.take<CopyDesc>((size_t)crops_cap);
    int32_t *hist = bump.take<int32_t>((size_t)n_buckets);
    int32_t *cnt = bump.take<int32_t>(4);  // generic-list length, fast-crop count, ticket counter
    if (!plans || !redo || !cnt) {
        ms_set_error("crop: scratch too small");
        return MS_ERR_CAPACITY;
    }
    int32_t *glist_n = cnt, *n_work = cnt + 1, *ticket = cnt + 2;
    MS_CUDA(cudaMemsetAsync(redo, 0, (size_t)crops_cap, st));
    MS_CUDA(cudaMemsetAsync(hist, 0, (size_t)n_buckets * sizeof(int32_t), st));
    MS_CUDA(cudaMemsetAsync(cnt, 0, 4 * sizeof(int32_t), st));
    int64_t lgrid = (crops_cap + 255) / 256;
    if (lgrid > (int64_t)ctx->num_sms * 8) lgrid = (int64_t)ctx->num_sms * 8;
    ms_launch(crop_plan_kernel, (int)lgrid, 256, 0, st, pages, ragged ? page_ptrs : nullptr, page_hw, n_pages, img_h, img_w, crops,
                                                 n_crops, range, crops_cap, out_h, out_w, plans, hist);
    MS_LAUNCH_CHECK(ctx);
    ms_launch(crop_bucket_scan_kernel, 1, 1024, 0, st, hist, n_buckets, n_work);
    MS_LAUNCH_CHECK(ctx);
    ms_launch(crop_bucket_scatter_kernel, (int)lgrid, 256, 0, st, plans, ragged ? page_hw : nullptr, n_pages, img_h, img_w, n_crops,
                                                           range, crops_cap, hist, work);
    MS_LAUNCH_CHECK(ctx);
    int64_t grid = (int64_t)ctx->num_sms * per_sm;
    if (grid > crops_cap) grid = crops_cap;
    // one CTA per listed crop, many in flight: a generic crop reads global memory byte by byte and is latency bound
    int64_t ggrid = (int64_t)ctx->num_sms * 8;
    if (ggrid > crops_cap) ggrid = crops_cap;
    const int vec_ok = ((out_w & 3) == 0 && (reinterpret_cast<uintptr_t>(batch_f32) & 15) == 0) ? 1 : 0;
    // cudaFuncSetAttribute is a synchronous driver call (milliseconds): once per context and kernel, not per launch
#define MS_CROP_LAUNCH(F32, U8)                                                                                        \
    do {                                                                                                               \
        auto kfn = crop_resize_pad_kernel<F32, U8>;                                                                    \
        int &granted = ctx->smem_attr[(F32 ? 1 : 0) + (U8 ? 2 : 0) - 1];                                               \
        if ((int)smem > granted) {                                                                                     \
            MS_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
            granted = (int)smem;                                                                                       \
        }                                                                                                              \
        ms_launch(kfn, (int)grid, kThreads, smem, st, plans, work, n_work, ticket, (int)ring, out_h, out_w, batch_f32,      \
                                              canvas_u8, vec_ok, redo);                                               \
        MS_LAUNCH_CHECK(ctx);                                                                                          \
        ms_launch(crop_generic_list_kernel, (int)lgrid, 256, 0, st, plans, n_crops, range, crops_cap, redo, glist, glist_n);  \
        MS_LAUNCH_CHECK(ctx);                                                                                          \
        ms_launch(crop_generic_kernel<F32, U8>, (int)ggrid, 256, 0, st, plans, glist, glist_n, out_h,                        \
                                                                 out_w, batch_f32, canvas_u8, vec_ok);                \
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
