// Measurement aid: is a lone warp's float64 chain bound by latency or by the issue rate of the FP64 pipe?  One warp
// (and W warps of one block, i.e. one per sub-partition for W <= 4, several per sub-partition above) run K independent
// DFMA chains each; the probe prints cycles per DFMA issued by a warp and per sub-partition.
//   nvcc -arch=sm_100a -O3 -o probes/fp64_rate_probe tools/fp64_rate_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define REP 128

template <int K>
__global__ void rate_kernel(double seed, double* out, long long* cycles) {
    double x[K];
#pragma unroll
    for (int k = 0; k < K; ++k) x[k] = seed + threadIdx.x * 1e-9 + k;
    const double y = seed * 0.5;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int outer = 0; outer < 16; ++outer) {
#pragma unroll
        for (int i = 0; i < REP; ++i) {
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = fma(x[k], y, 1e-3);
        }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cycles[threadIdx.x >> 5] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += x[k];
    out[threadIdx.x] = s;
}

template <int K>
void run(int warps, double* out, long long* cyc) {
    for (int rep = 0; rep < 2; ++rep) {
        rate_kernel<K><<<1, 32 * warps>>>(1.000001, out, cyc);
        cudaDeviceSynchronize();
    }
    long long h[32];
    cudaMemcpy(h, cyc, sizeof(long long) * warps, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    const double per_warp = (double)mx / (16.0 * REP * K);
    const int per_smsp = (warps + 3) / 4;
    printf("warps %2d  chains/warp %d : %6.2f cycles per DFMA of a warp, %6.2f per DFMA of a sub-partition\n", warps, K,
           per_warp, per_warp / per_smsp);
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1024 * sizeof(double));
    cudaMalloc(&cyc, 32 * sizeof(long long));
    run<1>(1, out, cyc); run<2>(1, out, cyc); run<4>(1, out, cyc); run<8>(1, out, cyc);
    run<1>(4, out, cyc); run<4>(4, out, cyc);
    run<1>(8, out, cyc); run<2>(8, out, cyc); run<4>(8, out, cyc);
    run<1>(16, out, cyc); run<1>(32, out, cyc); run<4>(32, out, cyc);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
