// Measurement aid: dependent-issue latency (cycles) of the instructions on the EM critical path of
// select_gmm_kernel, one warp on one SM.   nvcc -arch=sm_100a -O3 -o probes/lat_probe tools/lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define REP 256

template <int OP>
__global__ void lat_kernel(double seed, double* out, long long* cycles) {
    double x = seed + threadIdx.x * 1e-9, y = seed * 0.5;
    float xf = (float)seed + threadIdx.x * 1e-6f;
    long long t0 = clock64();
#pragma unroll 1
    for (int outer = 0; outer < 16; ++outer) {
#pragma unroll
        for (int i = 0; i < REP; ++i) {
            if (OP == 0) x = fma(x, y, 1e-3);                              // DFMA
            if (OP == 1) x = x + y;                                        // DADD
            if (OP == 2) x = x * y;                                        // DMUL
            if (OP == 3) asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(x));
            if (OP == 4) x = __shfl_xor_sync(0xffffffffu, x, 1);           // 64-bit shuffle (2 SHFL)
            if (OP == 5) xf = fmaf(xf, 0.999f, 1e-3f);                     // FFMA
            if (OP == 6) xf = __fsqrt_rn(xf + 1.0f);                       // IEEE sqrt
            if (OP == 7) xf = __fdiv_rn(1.0f, xf + 1.0f);                  // IEEE div
            if (OP == 8) x = (double)(float)x;                             // F2F round trip
            if (OP == 9) x = rint(x * 1.0000001);                          // DMUL + rint
            if (OP == 10) x = __shfl_xor_sync(0xffffffffu, x, 1) + x;      // one reduction round
            if (OP == 11) xf = __shfl_xor_sync(0xffffffffu, xf, 1);        // 32-bit shuffle
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    out[threadIdx.x] = x + (double)xf;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 32 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    const char* names[] = {"DFMA", "DADD", "DMUL", "MUFU.RCP64H", "SHFL64", "FFMA", "fsqrt_rn(+FADD)",
                           "fdiv_rn(+FADD)", "F2F.F32.F64+F2F.F64.F32", "DMUL+rint", "SHFL64+DADD", "SHFL32"};
#define RUN(OP)                                                                         \
    for (int rep = 0; rep < 2; ++rep) {                                                 \
        lat_kernel<OP><<<1, 32>>>(1.000001, out, cyc);                                  \
        cudaDeviceSynchronize();                                                        \
    }                                                                                   \
    {                                                                                   \
        long long h;                                                                    \
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);                         \
        printf("%-28s %.1f cycles\n", names[OP], (double)h / (16.0 * REP));             \
    }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11)
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
