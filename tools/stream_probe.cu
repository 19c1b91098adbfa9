// Measurement aid (not part of the product): what does a flat gated read of the post-processor's logits (464 MB at 64
// images) achieve on B200 straight after the bench's 256 MiB flush write?  Variants: (A) grid-stride float4 loads with
// U loads in flight per thread; (B) 1-D bulk asynchronous copies (cp.async.bulk + mbarrier) into a shared-memory ring
// drained by the block's warps.  Every variant does the candidate kernel's gate (max over the 16 values a thread
// holds, compare) and counts the passing elements, so that nothing is optimised away.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/stream_probe tools/stream_probe.cu && /tmp/stream_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int U>
__global__ void __launch_bounds__(1024) flat_kernel(const float4* __restrict__ src, size_t n4, float gate,
                                                    unsigned long long* __restrict__ out) {
    unsigned cnt = 0;
    const size_t T = blockDim.x;
    const size_t stride = (size_t)gridDim.x * T * U;
    for (size_t base = (size_t)blockIdx.x * T * U + threadIdx.x; base < n4; base += stride) {
        float4 x[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const size_t i = base + (size_t)j * T;
            x[j] = i < n4 ? __ldcs(src + i) : make_float4(-1e30f, -1e30f, -1e30f, -1e30f);
        }
        float mx = -1e30f;
#pragma unroll
        for (int j = 0; j < U; ++j) mx = fmaxf(mx, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
        if (mx > gate) {
#pragma unroll
            for (int j = 0; j < U; ++j)
                cnt += (x[j].x > gate) + (x[j].y > gate) + (x[j].z > gate) + (x[j].w > gate);
        }
    }
    if (cnt) atomicAdd(out, (unsigned long long)cnt);
}

// load flavours: 0 = __ldcs (ld.global.cs), 1 = __ldg (ld.global.nc), 2 = ld.global.nc.L1::no_allocate.L2::256B,
// 3 = ld.global.nc.L2::256B, 4 = plain ld.global
template <int F>
__device__ __forceinline__ float4 load_f(const float4* p) {
    float4 v;
    if (F == 0) return __ldcs(p);
    if (F == 1) return __ldg(p);
    if (F == 2)
        asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (F == 3)
        asm volatile("ld.global.nc.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (F == 4)
        asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int U, int F>
__global__ void __launch_bounds__(1024) flavour_kernel(const float4* __restrict__ src, size_t n4, float gate,
                                                       unsigned long long* __restrict__ out) {
    unsigned cnt = 0;
    const size_t T = blockDim.x;
    const size_t stride = (size_t)gridDim.x * T * U;
    size_t base = (size_t)blockIdx.x * T * U + threadIdx.x;
    const float4 pad = make_float4(-1e30f, -1e30f, -1e30f, -1e30f);
    float4 x[U], nx[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const size_t i = base + (size_t)j * T;
        x[j] = i < n4 ? load_f<F>(src + i) : pad;
    }
    for (; base < n4; base += stride) {
        const size_t nb = base + stride;
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const size_t i = nb + (size_t)j * T;
            nx[j] = i < n4 ? load_f<F>(src + i) : pad;
        }
        float mx = -1e30f;
#pragma unroll
        for (int j = 0; j < U; ++j) mx = fmaxf(mx, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
        if (mx > gate) {
#pragma unroll
            for (int j = 0; j < U; ++j)
                cnt += (x[j].x > gate) + (x[j].y > gate) + (x[j].z > gate) + (x[j].w > gate);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) x[j] = nx[j];
    }
    if (cnt) atomicAdd(out, (unsigned long long)cnt);
}

// prefetch-one-chunk-ahead variant (the structure of post_candidates_kernel without its queue)
template <int U>
__global__ void __launch_bounds__(1024) ahead_kernel(const float4* __restrict__ src, size_t n4, float gate,
                                                     unsigned long long* __restrict__ out) {
    unsigned cnt = 0;
    const size_t T = blockDim.x;
    const size_t stride = (size_t)gridDim.x * T * U;
    size_t base = (size_t)blockIdx.x * T * U + threadIdx.x;
    float4 x[U], nx[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const size_t i = base + (size_t)j * T;
        x[j] = i < n4 ? __ldcs(src + i) : make_float4(-1e30f, -1e30f, -1e30f, -1e30f);
    }
    for (; base < n4; base += stride) {
        const size_t nb = base + stride;
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const size_t i = nb + (size_t)j * T;
            nx[j] = i < n4 ? __ldcs(src + i) : make_float4(-1e30f, -1e30f, -1e30f, -1e30f);
        }
        float mx = -1e30f;
#pragma unroll
        for (int j = 0; j < U; ++j) mx = fmaxf(mx, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
        if (mx > gate) {
#pragma unroll
            for (int j = 0; j < U; ++j)
                cnt += (x[j].x > gate) + (x[j].y > gate) + (x[j].z > gate) + (x[j].w > gate);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) x[j] = nx[j];
    }
    if (cnt) atomicAdd(out, (unsigned long long)cnt);
}

// ---- bulk-copy ring -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}\n" ::"r"(
            smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int STAGES, int STAGE_BYTES>
__global__ void __launch_bounds__(256) ring_kernel(const float4* __restrict__ src, size_t n4, float gate,
                                                    unsigned long long* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[STAGES], empty[STAGES];
    constexpr int V = STAGE_BYTES / 16;                 // float4 per stage
    const size_t n_chunks = (n4 + V - 1) / V;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 256 / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned cnt = 0;
    const int lane = threadIdx.x & 31;
    // chunks of this block: blockIdx.x, + gridDim.x, ...
    size_t issue = blockIdx.x;
    int issued = 0;
    auto issue_one = [&]() {
        if (issue < n_chunks) {
            const int s = issued % STAGES;
            const size_t first = issue * V;
            const unsigned bytes = (unsigned)(min((size_t)V, n4 - first) * 16);
            mbar_expect_tx(&full[s], bytes);
            bulk_load(smem + (size_t)s * STAGE_BYTES, src + first, bytes, &full[s]);
        }
        issue += gridDim.x;
        ++issued;
    };
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES; ++s) issue_one();
    int k = 0;
    for (size_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x, ++k) {
        const int s = k % STAGES;
        const unsigned par = (k / STAGES) & 1;
        mbar_wait(&full[s], par);
        const size_t first = ch * V;
        const int valid = (int)min((size_t)V, n4 - first);
        const float4* p = reinterpret_cast<const float4*>(smem + (size_t)s * STAGE_BYTES);
        float mx = -1e30f;
        float4 x[V / 256];
#pragma unroll
        for (int j = 0; j < V / 256; ++j) {
            const int i = j * 256 + threadIdx.x;
            x[j] = i < valid ? p[i] : make_float4(-1e30f, -1e30f, -1e30f, -1e30f);
            mx = fmaxf(mx, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
        }
        if (mx > gate) {
#pragma unroll
            for (int j = 0; j < V / 256; ++j)
                cnt += (x[j].x > gate) + (x[j].y > gate) + (x[j].z > gate) + (x[j].w > gate);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (threadIdx.x == 0) {                          // refill the stage once all warps have let go of it
            mbar_wait(&empty[s], par);
            issue_one();
        }
    }
    if (cnt) atomicAdd(out, (unsigned long long)cnt);
}

int main() {
    const size_t n = (size_t)64 * 80 * 22400;            // floats: C4, 64 images
    const size_t n4 = n / 4;
    float* d;
    CK(cudaMalloc(&d, n * 4));
    std::vector<float> h(1 << 20);
    srand(1);
    for (auto& v : h) v = (rand() % 1000 < 3) ? 0.5f : -6.0f + ((rand() % 2001) / 1000.0f);     // 0.3 % above the gate
    for (size_t o = 0; o < n; o += h.size())
        CK(cudaMemcpy(d + o, h.data(), std::min(h.size(), n - o) * 4, cudaMemcpyHostToDevice));
    unsigned char* flush;
    CK(cudaMalloc(&flush, 256u << 20));
    unsigned long long* out;
    CK(cudaMalloc(&out, 8));
    const float gate = -2.9454f;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto timeit = [&](const char* name, auto launch) {
        std::vector<float> ts;
        unsigned long long hits = 0;
        for (int r = 0; r < 12; ++r) {
            CK(cudaMemsetAsync(flush, r, 256u << 20));
            CK(cudaMemsetAsync(out, 0, 8));
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            ts.push_back(ms * 1000.f);
            CK(cudaMemcpy(&hits, out, 8, cudaMemcpyDeviceToHost));
        }
        std::sort(ts.begin(), ts.end());
        printf("%-44s %8.1f us  %7.0f GB/s   (gated %.3f %%)\n", name, ts[ts.size() / 2], n * 4 / ts[ts.size() / 2] / 1e3,
               100.0 * hits / n);
    };
    const float4* s4 = reinterpret_cast<const float4*>(d);
    printf("bytes %.1f MB, after a 256 MiB flush write\n", n * 4 / 1e6);
#define FLAT(U, B, T) timeit("flat U=" #U " blocks/SM=" #B " threads=" #T, [&] { flat_kernel<U><<<148 * B, T>>>(s4, n4, gate, out); })
    FLAT(4, 4, 256); FLAT(4, 2, 256); FLAT(8, 2, 256); FLAT(16, 2, 256); FLAT(16, 1, 256); FLAT(8, 1, 512); FLAT(16, 1, 512);
    FLAT(4, 1, 1024); FLAT(8, 1, 1024); FLAT(16, 1, 1024); FLAT(4, 2, 512); FLAT(8, 2, 512); FLAT(4, 2, 1024);
#define AHEAD(U, B, T) timeit("one chunk ahead U=" #U " blocks/SM=" #B " threads=" #T, [&] { ahead_kernel<U><<<148 * B, T>>>(s4, n4, gate, out); })
    AHEAD(4, 4, 256); AHEAD(4, 2, 256); AHEAD(8, 2, 256); AHEAD(8, 1, 256); AHEAD(4, 1, 512); AHEAD(8, 1, 512); AHEAD(4, 1, 1024);
    AHEAD(8, 1, 1024); AHEAD(4, 2, 512); AHEAD(4, 3, 256);
#define FLAV(F) timeit("one chunk ahead U=4 blocks/SM=4 threads=256 flavour " #F, [&] { flavour_kernel<4, F><<<148 * 4, 256>>>(s4, n4, gate, out); })
    FLAV(0); FLAV(1); FLAV(2); FLAV(3); FLAV(4);
#define RING(S, KB, B)                                                                                            \
    do {                                                                                                          \
        CK(cudaFuncSetAttribute(ring_kernel<S, KB * 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, S * KB * 1024)); \
        timeit("bulk ring stages=" #S " x " #KB " KB blocks/SM=" #B,                                            \
               [&] { ring_kernel<S, KB * 1024><<<148 * B, 256, S * KB * 1024>>>(s4, n4, gate, out); });           \
    } while (0)
    RING(6, 16, 2); RING(12, 16, 1); RING(6, 32, 1); RING(3, 64, 1); RING(4, 48, 1); RING(3, 32, 2); RING(5, 40, 1); RING(13, 16, 1);
    return 0;
}
