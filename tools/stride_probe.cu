// Measurement aid: read bandwidth of "class-row" access patterns on a [N, C, HW] float tensor.
// Each block owns `run` consecutive hw positions of one image and walks all C rows, `batch` rows in flight.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
template <int BATCH>
__global__ void rows_kernel(const float4* __restrict__ x, float* __restrict__ out, int N, int C, int hw4, int run4,
                            int blocks_per_image, int order) {
    int b = blockIdx.x;
    int n, r;
    if (order == 0) { n = b / blocks_per_image; r = b % blocks_per_image; }      // image-major
    else { r = b / N; n = b % N; }                                               // image-minor
    int i4 = r * run4 + threadIdx.x;
    if (threadIdx.x >= run4 || i4 >= hw4) return;
    const float4* p = x + (size_t)n * C * hw4 + i4;
    float acc = 0.f;
    for (int c0 = 0; c0 < C; c0 += BATCH) {
        float4 v[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) v[j] = __ldcs(p + (size_t)(c0 + j) * hw4);
#pragma unroll
        for (int j = 0; j < BATCH; ++j) acc += v[j].x + v[j].y + v[j].z + v[j].w;
    }
    if (acc == 12345.678f) out[0] = acc;
}
__global__ void flat_kernel(const float4* __restrict__ x, float* __restrict__ out, size_t n4) {
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x * 4 + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x * 4) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (i + j * blockDim.x < n4) ? __ldcs(x + i + j * blockDim.x) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += v[j].x + v[j].y + v[j].z + v[j].w;
    }
    if (acc == 12345.678f) out[0] = acc;
}
int main() {
    const int N = 16, C = 80, hw = 16800, hw4 = hw / 4;
    size_t n4 = (size_t)N * C * hw4;
    float4* x; float* out; char* fl;
    cudaMalloc(&x, n4 * 16); cudaMalloc(&out, 4); cudaMalloc(&fl, 256 << 20);
    cudaMemset(x, 0, n4 * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](auto launch) {
        float best = 1e9;
        for (int it = 0; it < 6; ++it) {
            cudaMemsetAsync(fl, 1, 256 << 20);
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
        }
        return best * 1000.f;
    };
    double bytes = (double)n4 * 16;
    float us = timeit([&] { flat_kernel<<<148 * 5, 256>>>(x, out, n4); });
    printf("flat read                         %7.1f us %7.0f GB/s\n", us, bytes / us / 1e3);
    for (int order = 0; order < 2; ++order)
        for (int threads : {128, 256, 1024})
            for (int batch : {4, 8, 16}) {
                int run4 = threads;
                int bpi = (hw4 + run4 - 1) / run4;
                float t;
                if (batch == 4) t = timeit([&] { rows_kernel<4><<<N * bpi, threads>>>(x, out, N, C, hw4, run4, bpi, order); });
                else if (batch == 8) t = timeit([&] { rows_kernel<8><<<N * bpi, threads>>>(x, out, N, C, hw4, run4, bpi, order); });
                else t = timeit([&] { rows_kernel<16><<<N * bpi, threads>>>(x, out, N, C, hw4, run4, bpi, order); });
                printf("rows order=%d piece=%5dB batch=%2d blocks=%5d  %7.1f us %7.0f GB/s\n", order, threads * 16, batch, N * bpi, t, bytes / t / 1e3);
            }
    return 0;
}
