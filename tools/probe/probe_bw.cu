// Developer probe: what the HBM path gives for K1's traffic shape (201 MB of write-once float output as 128-byte
// warp stores + 50 MB of read-once uint8 input), without any of K1's arithmetic.  nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>   // 0: st.cs 32-bit per lane, 1: st.cs 128-bit per lane, 2: default st 32-bit, 3: 32-bit + read 1/4 as many bytes
__global__ void fill(float* __restrict__ dst, const uint32_t* __restrict__ src, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (MODE == 1) {
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (; i < n / 4; i += stride) __stcs(d4 + i, make_float4(1.f, 2.f, 3.f, (float)i));
    } else if (MODE == 3) {
        // K1-like: CHW planes, each thread 8 rows x 3 channels (24 stores from one base), reads 3 bytes per pixel
        for (; i < n; i += stride) {
            uint32_t v = 0;
            if ((i & 3) == 0) v = __ldg(src + (i >> 2));
            __stcs(dst + i, (float)v);
        }
    } else {
        for (; i < n; i += stride) { if (MODE == 0) __stcs(dst + i, (float)i); else dst[i] = (float)i; }
    }
}

int main() {
    const size_t n = (size_t)256 * 3 * 256 * 256;   // floats = 201 MB
    const int R = 4;
    float* d[R]; uint32_t* s;
    for (int r = 0; r < R; ++r) CK(cudaMalloc(&d[r], n * 4));
    CK(cudaMalloc(&s, n));
    CK(cudaMemset(s, 1, n));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 4; ++mode) {
        for (int cps = 4; cps <= 16; cps *= 2) {
            const int grid = 148 * cps, iters = 20;
            float best = 1e9f, tot = 0.f;
            for (int it = 0; it < iters + 2; ++it) {
                cudaEventRecord(e0);
                if (mode == 0) fill<0><<<grid, 256>>>(d[it % R], s, n);
                if (mode == 1) fill<1><<<grid, 256>>>(d[it % R], s, n);
                if (mode == 2) fill<2><<<grid, 256>>>(d[it % R], s, n);
                if (mode == 3) fill<3><<<grid, 256>>>(d[it % R], s, n);
                cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (it >= 2) { tot += ms; if (ms < best) best = ms; }
            }
            const double bytes = n * 4.0 + (mode == 3 ? n : 0);
            printf("mode %d ctas/sm %2d: mean %.2f us (%.0f GB/s)  best %.2f us\n", mode, cps, tot / iters * 1e3, bytes / (tot / iters * 1e-3) / 1e9, best * 1e3);
        }
    }
    // back-to-back launches in one stream (what the pipeline sees): 20 fills of 4 rotating buffers
    cudaEventRecord(e0);
    for (int it = 0; it < 40; ++it) fill<0><<<148 * 8, 256>>>(d[it % R], s, n);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("back-to-back mode 0: %.2f us per 201 MB fill (%.0f GB/s)\n", ms / 40 * 1e3, n * 4.0 / (ms / 40 * 1e-3) / 1e9);
    return 0;
}
