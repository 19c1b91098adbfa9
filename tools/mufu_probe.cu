// Measurement aid: MUFU (ex2 / lg2 / rcp) and FFMA throughput per SM on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probes/mufu_probe tools/mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = seed + 0.001f * (threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
            if (OP == 1) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
            if (OP == 3) v[j] = fmaf(v[j], v[j], 0.5f);
            if (OP == 4) {   // the focal mix: ex2, add, lg2, rcp
                float e, l, r;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v[j]));
                float u = 1.0f + e;
                asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u));
                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(u));
                v[j] = l * r - 1.0f;
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    if (s == 123.456f) out[0] = s;
}
int main() {
    float* out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, blocks = 148 * 8;
    const char* names[] = {"ex2", "lg2", "rcp", "ffma", "ex2+lg2+rcp(+3 fp32)"};
    for (int op = 0; op < 5; ++op) {
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (op == 0) k<0><<<blocks, 256>>>(out, iters, -0.5f);
            if (op == 1) k<1><<<blocks, 256>>>(out, iters, 1.5f);
            if (op == 2) k<2><<<blocks, 256>>>(out, iters, 1.5f);
            if (op == 3) k<3><<<blocks, 256>>>(out, iters, 0.5f);
            if (op == 4) k<4><<<blocks, 256>>>(out, iters, -0.5f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double ops = (double)blocks * 256 * iters * 8 * (op == 4 ? 3 : 1);
        printf("%-22s %8.3f ms  %6.1f %s/clk/SM (at 1965 MHz)\n", names[op], best, ops / (best * 1e-3) / 148 / 1.965e9,
               op == 4 ? "MUFU" : "ops");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
