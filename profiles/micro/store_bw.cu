// Micro-benchmark (profiles/micro): how fast can W warps per CTA (4 CTAs / SM, 148 SMs) stream float4 stores to HBM?
// Answers whether ONE padding warp per CTA of crop_resize_pad_kernel can be write-bandwidth bound by itself.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu && ./store_bw
#include <cstdio>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

// modes 5: what a crop kernel with its canvas rows assembled in shared memory would write -- per canvas ONE TMA tensor
// store of the pixel box (kPix floats x 96 rows out of a dense tile) + 12 tensor stores of the padding tails ((128 - kPix)
// floats x 8 rows out of a block of 1.0f); one thread per CTA issues, `depth` canvases in flight.
struct Maps2 {
    CUtensorMap pix, tail;
};
__global__ void __launch_bounds__(224, 4) tensor_kernel(const __grid_constant__ Maps2 maps, size_t canvases, int pix, int depth)
{
    extern __shared__ __align__(128) unsigned char sm[];  // [96][pix] tile, then [8][128 - pix] ones
    float *tile = reinterpret_cast<float *>(sm), *ones = tile + 96 * pix;
    for (int i = threadIdx.x; i < 96 * pix + 8 * (128 - pix); i += blockDim.x) tile[i] = 1.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x != 0) return;
    const unsigned stile = (unsigned)__cvta_generic_to_shared(tile), sones = (unsigned)__cvta_generic_to_shared(ones);
    for (size_t c = blockIdx.x; c < canvases; c += gridDim.x) {
        const int row0 = (int)(c * 96);
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&maps.pix), "r"(0),
                     "r"(row0), "r"(stile)
                     : "memory");
        if (pix < 128)
            for (int b = 0; b < 12; b++)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&maps.tail),
                             "r"(pix), "r"(row0 + 8 * b), "r"(sones)
                             : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (depth == 1)
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        else
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int kMode>  // 0: contiguous per warp-iteration (512 B), 1: rows of 288 B inside 512 B rows (the padding pattern)
__global__ void __launch_bounds__(224, 4) store_kernel(float4 *out, size_t n4, int warps_active)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= warps_active) return;
    const size_t nw = (size_t)gridDim.x * warps_active, w = (size_t)blockIdx.x * warps_active + warp;
    const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
    if (kMode == 0) {
        for (size_t i = w * 32 + lane; i < n4; i += nw * 32) __stcs(out + i, one);
    } else {
        // canvas rows of 128 floats (32 float4); the last 18 float4 of every row are padding
        const size_t rows = n4 / 32;
        for (size_t r = w; r < rows; r += nw)
            if (lane < 18) __stcs(out + r * 32 + 14 + lane, one);
    }
}

// mode 2: the crop kernel's real write pattern.  A CTA takes canvases (96 rows of 512 bytes: 3 channels x 32 rows x 128
// floats) in turn; warp 0 writes the last 288 bytes of every row as 16-byte stores (the padding warp), warps 1..5 write
// the first 224 bytes as 4-byte stores, 32 consecutive floats per warp instruction (the consumers).  No synchronisation.
template <int kPix>  // floats of a row written by the pixel warps (a multiple of 4); the padding warp writes the rest
__global__ void __launch_bounds__(224, 4) split_kernel(float *out, size_t canvases)
{
    constexpr int kTail4 = (128 - kPix) / 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp > 5) return;
    for (size_t c = blockIdx.x; c < canvases; c += gridDim.x) {
        float *cv = out + c * 96 * 128;
        if (warp == 0) {
            const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
            for (int i = lane; i < 96 * kTail4; i += 32)
                __stcs(reinterpret_cast<float4 *>(cv + (i / kTail4) * 128 + kPix) + i % kTail4, one);
        } else {
            for (int i = (warp - 1) * 32 + lane; i < 96 * kPix; i += 160) __stcs(cv + (i / kPix) * 128 + i % kPix, 0.5f);
        }
    }
}

// the same split with the pixel part written kVec floats per lane (8- or 16-byte stores): what a consumer mapping with 2 or
// 4 adjacent columns per thread would issue
template <int kPix, int kVec>
__global__ void __launch_bounds__(224, 4) split_vec_kernel(float *out, size_t canvases)
{
    constexpr int kTail4 = (128 - kPix) / 4, kPv = kPix / kVec;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp > 5) return;
    for (size_t c = blockIdx.x; c < canvases; c += gridDim.x) {
        float *cv = out + c * 96 * 128;
        if (warp == 0) {
            const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
            for (int i = lane; i < 96 * kTail4; i += 32)
                __stcs(reinterpret_cast<float4 *>(cv + (i / kTail4) * 128 + kPix) + i % kTail4, one);
        } else {
            for (int i = (warp - 1) * 32 + lane; i < 96 * kPv; i += 160) {
                float *at = cv + (i / kPv) * 128 + (i % kPv) * kVec;
                if (kVec == 2)
                    __stcs(reinterpret_cast<float2 *>(at), make_float2(0.5f, 0.5f));
                else
                    __stcs(reinterpret_cast<float4 *>(at), make_float4(0.5f, 0.5f, 0.5f, 0.5f));
            }
        }
    }
}

// modes 3 / 4: whole canvases (or quarter canvases) as TMA bulk stores out of shared memory, one elected thread per CTA,
// `depth` stores in flight
__global__ void __launch_bounds__(224, 4) bulk_kernel(float *out, size_t canvases, int piece_bytes, int depth)
{
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < piece_bytes / 4; i += blockDim.x) reinterpret_cast<float *>(sm)[i] = 1.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int pieces = 96 * 512 / piece_bytes;
    const unsigned src = (unsigned)__cvta_generic_to_shared(sm);
    for (size_t c = blockIdx.x; c < canvases; c += gridDim.x)
        for (int p = 0; p < pieces; p++) {
            char *dst = reinterpret_cast<char *>(out) + c * 96 * 512 + (size_t)p * piece_bytes;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(piece_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (depth == 1)
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else if (depth == 2)
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else
                asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
        }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// mode 6: pixel part of every canvas row by five warps with 4-byte stores (as the crop kernel's consumers write it), the
// padding tails of the SAME canvas at the same time as TMA tensor stores issued by one more thread
__global__ void __launch_bounds__(224, 4) mixed_kernel(const __grid_constant__ Maps2 maps, float *out, size_t canvases, int pix)
{
    extern __shared__ __align__(128) unsigned char sm[];
    float *ones = reinterpret_cast<float *>(sm);
    for (int i = threadIdx.x; i < 8 * (128 - pix); i += blockDim.x) ones[i] = 1.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp > 5) return;
    const unsigned sones = (unsigned)__cvta_generic_to_shared(ones);
    for (size_t c = blockIdx.x; c < canvases; c += gridDim.x) {
        float *cv = out + c * 96 * 128;
        if (warp == 0) {
            if (lane < 12)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&maps.tail),
                             "r"(pix), "r"((int)(c * 96) + 8 * lane), "r"(sones)
                             : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        } else {
            for (int i = (warp - 1) * 32 + lane; i < 96 * pix; i += 160) __stcs(cv + (i / pix) * 128 + i % pix, 0.5f);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main()
{
    const size_t bytes = (size_t)6 << 30;
    float4 *buf;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; mode++)
        for (int w : {1, 2, 4, 7}) {
            float best = 1e9f;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                if (mode == 0)
                    store_kernel<0><<<592, 224>>>(buf, bytes / 16, w);
                else
                    store_kernel<1><<<592, 224>>>(buf, bytes / 16, w);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const double written = mode == 0 ? (double)bytes : (double)bytes * 18 / 32;
            printf("mode %d warps/CTA %d: %.3f ms  %.0f GB/s written\n", mode, w, best, written / best / 1e6);
        }
    {
        const size_t canvases = bytes / (96 * 512);
        float best = 1e9f;
        for (int pix : {56, 64, 48, 32, 96, 60}) {
            best = 1e9f;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                float *o = reinterpret_cast<float *>(buf);
                switch (pix) {
                case 56: split_kernel<56><<<592, 224>>>(o, canvases); break;
                case 64: split_kernel<64><<<592, 224>>>(o, canvases); break;
                case 48: split_kernel<48><<<592, 224>>>(o, canvases); break;
                case 32: split_kernel<32><<<592, 224>>>(o, canvases); break;
                case 96: split_kernel<96><<<592, 224>>>(o, canvases); break;
                default: split_kernel<60><<<592, 224>>>(o, canvases); break;
                }
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("split pattern, %d pixel floats + %d padding floats per row (padding warp + 5 pixel warps): %.3f ms  %.0f GB/s written\n",
                   pix, 128 - pix, best, (double)canvases * 96 * 512 / best / 1e6);
        }
        for (int vec : {2, 4}) {
            best = 1e9f;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                if (vec == 2)
                    split_vec_kernel<40, 2><<<592, 224>>>(reinterpret_cast<float *>(buf), canvases);
                else
                    split_vec_kernel<40, 4><<<592, 224>>>(reinterpret_cast<float *>(buf), canvases);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("split pattern, 40 pixel floats written %d per lane (%d-byte stores) + 88 padding floats: %.3f ms  %.0f GB/s written\n",
                   vec, 4 * vec, best, (double)canvases * 96 * 512 / best / 1e6);
        }
        {
            best = 1e9f;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                split_kernel<40><<<592, 224>>>(reinterpret_cast<float *>(buf), canvases);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("split pattern, 40 pixel floats written 1 per lane (4-byte stores) + 88 padding floats: %.3f ms  %.0f GB/s written\n",
                   best, (double)canvases * 96 * 512 / best / 1e6);
        }
        cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 512);
        for (int piece : {96 * 512, 24 * 512, 8 * 512})
            for (int depth : {1, 2, 4}) {
                best = 1e9f;
                for (int rep = 0; rep < 4; rep++) {
                    cudaEventRecord(e0);
                    bulk_kernel<<<592, 224, piece>>>(reinterpret_cast<float *>(buf), canvases, piece, depth);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (ms < best) best = ms;
                }
                printf("TMA bulk stores of %d bytes, %d in flight per CTA: %.3f ms  %.0f GB/s written\n", piece, depth, best,
                       (double)canvases * 96 * 512 / best / 1e6);
            }
    }
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        PFN_cuTensorMapEncodeTiled encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
        const size_t canvases = bytes / (96 * 512);
        cudaFuncSetAttribute(tensor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024);
        for (int pix : {40, 64, 128, 32, 96})
            for (int depth : {1, 2}) {
                Maps2 maps;
                const cuuint64_t gdim[2] = {128, (cuuint64_t)canvases * 96};
                const cuuint64_t gstride[1] = {512};
                const cuuint32_t estride[2] = {1, 1};
                const cuuint32_t bp[2] = {(cuuint32_t)pix, 96}, bt[2] = {(cuuint32_t)(pix < 128 ? 128 - pix : 4), 8};
                CUresult r1 = encode(&maps.pix, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, gdim, gstride, bp, estride,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                CUresult r2 = encode(&maps.tail, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, gdim, gstride, bt, estride,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
                    printf("encode failed %d %d\n", (int)r1, (int)r2);
                    continue;
                }
                const size_t smem = (size_t)(96 * pix + 8 * (128 - pix)) * 4;
                float best = 1e9f;
                for (int rep = 0; rep < 4; rep++) {
                    cudaEventRecord(e0);
                    tensor_kernel<<<592, 224, smem>>>(maps, canvases, pix, depth);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (ms < best) best = ms;
                }
                printf("TMA tensor stores: pixel box %d x 96 + 12 tail boxes %d x 8 per canvas, %d in flight: %.3f ms  %.0f GB/s written (%s)\n",
                       pix, 128 - pix, depth, best, (double)canvases * 96 * 512 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
                if (depth == 1 && pix < 128) {
                    best = 1e9f;
                    for (int rep = 0; rep < 4; rep++) {
                        cudaEventRecord(e0);
                        mixed_kernel<<<592, 224, 8 * (128 - pix) * 4>>>(maps, reinterpret_cast<float *>(buf), canvases, pix);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                        float ms;
                        cudaEventElapsedTime(&ms, e0, e1);
                        if (ms < best) best = ms;
                    }
                    printf("mixed: %d pixel floats per row by 4-byte stores of five warps + the tails as TMA tensor stores, same canvas: %.3f ms  %.0f GB/s written (%s)\n",
                           pix, best, (double)canvases * 96 * 512 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
                }
            }
    }
    float best = 1e9f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        cudaMemsetAsync(buf, 1, bytes);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("cudaMemset: %.3f ms  %.0f GB/s\n", best, bytes / best / 1e6);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
