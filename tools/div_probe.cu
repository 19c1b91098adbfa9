// Verification aid: the branch-free float32 division used in the IoU hot loop (reciprocal seed + FMA
// refinement + remainder corrections, no special-case path) must equal __fdiv_rn bit for bit on the operand
// ranges of box IoUs (normal, positive).   nvcc -arch=sm_100a -O3 -o probes/div_probe tools/div_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float div_rn_normal(float a, float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    r = fmaf(fmaf(-b, r, 1.0f), r, r);
    float q = a * r;
    q = fmaf(fmaf(-b, q, a), r, q);
    q = fmaf(fmaf(-b, q, a), r, q);
    return q;
}
__device__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__global__ void k(unsigned long long* bad, unsigned long long n_per_thread, int mode) {
    unsigned long long id = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long local = 0;
    for (unsigned long long i = 0; i < n_per_thread; ++i) {
        unsigned h1 = hash((unsigned)(id * n_per_thread + i) * 2u + 1u), h2 = hash(h1 + 0x9e3779b9u);
        float a, b;
        if (mode == 0) {            // IoU-like: inter in [1, 2^22], union >= inter
            a = 1.0f + (float)(h1 & 0x3fffff) * (1.0f / 4.0f);
            b = a + (float)(h2 & 0xffffff) * (1.0f / 16.0f);
        } else if (mode == 1) {     // arbitrary normal mantissas, exponents within +-20
            a = __uint_as_float((h1 & 0x007fffffu) | ((107u + (h1 >> 26)) << 23));
            b = __uint_as_float((h2 & 0x007fffffu) | ((107u + (h2 >> 26)) << 23));
        } else {                    // integer-valued areas (the +1 convention on integer pixel boxes)
            a = (float)(1 + (h1 & 0xfffff));
            b = (float)(1 + (h1 & 0xfffff) + (h2 & 0xfffff));
        }
        if (__float_as_uint(div_rn_normal(a, b)) != __float_as_uint(__fdiv_rn(a, b))) ++local;
    }
    if (local) atomicAdd(bad, local);
}
int main() {
    unsigned long long* bad; cudaMalloc(&bad, 8);
    for (int mode = 0; mode < 3; ++mode) {
        cudaMemset(bad, 0, 8);
        k<<<148 * 8, 256>>>(bad, 4096, mode);
        unsigned long long h = 0; cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost);
        printf("mode %d: %llu mismatches of %llu\n", mode, h, 148ull * 8 * 256 * 4096);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
