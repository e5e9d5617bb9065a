// Micro-benchmark (profiles/micro): how fast can W warps per CTA (4 CTAs / SM, 148 SMs) stream float4 stores to HBM?
// Answers whether ONE padding warp per CTA of crop_resize_pad_kernel can be write-bandwidth bound by itself.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu && ./store_bw
#include <cstdio>
#include <cuda_runtime.h>

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
