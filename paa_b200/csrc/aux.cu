// Operators on either side of the PAA hot path: the anchor grid the heads are scored against,
// its visibility field, and the pairwise IoU that the evaluation / test-time-augmentation code of the
// reference calls on BoxLists.
//   AnchorGenerator.grid_anchors      paa_core/modeling/rpn/anchor_generator.py:73-95
//   AnchorGenerator.add_visibility_to anchor_generator.py:97-110
//   boxlist_iou                       paa_core/structures/boxlist_ops.py:81-116
#include "kernels.h"

namespace paa {

// anchors[(y*W + x)*a + k] = cell[k] + (x*stride, y*stride, x*stride, y*stride): one thread per anchor.
__global__ void __launch_bounds__(256)
grid_anchors_kernel(const float* __restrict__ cell, int a, int H, int W, float stride, float4* __restrict__ out) {
    const size_t total = (size_t)H * W * a;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % (size_t)a);
        const size_t loc = i / (size_t)a;
        const int x = (int)(loc % (size_t)W), y = (int)(loc / (size_t)W);
        // torch.arange(0, W*stride, step=stride, dtype=float32) evaluates start + i*step in float32
        const float sx = __fmul_rn((float)x, stride), sy = __fmul_rn((float)y, stride);
        const float4 c = ldg4(cell + (size_t)k * 4);
        out[i] = make_float4(__fadd_rn(sx, c.x), __fadd_rn(sy, c.y), __fadd_rn(sx, c.z), __fadd_rn(sy, c.w));
    }
}

__global__ void __launch_bounds__(256)
anchor_visibility_kernel(const float4* __restrict__ anchors, size_t n, float img_w, float img_h, float thresh,
                         unsigned char* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned char v = 1;
        if (thresh >= 0.0f) {
            const float4 b = anchors[i];
            v = (b.x >= -thresh) && (b.y >= -thresh) && (b.z < __fadd_rn(img_w, thresh)) &&
                (b.w < __fadd_rn(img_h, thresh));
        }
        out[i] = v;
    }
}

// out[i, j] = IoU(+1 convention) of boxes1[i] and boxes2[j] in the reference's float32 operation order.
// One block computes a 16 x 256 tile: the 16 row boxes live in shared memory, a thread owns one column.
constexpr int kIouRows = 16;
__global__ void __launch_bounds__(256)
boxlist_iou_kernel(const float* __restrict__ b1, int n1, const float* __restrict__ b2, int n2,
                   float* __restrict__ out) {
    __shared__ float4 s_box[kIouRows];
    __shared__ float s_area[kIouRows];
    const int r0 = blockIdx.y * kIouRows;
    if (threadIdx.x < kIouRows && r0 + threadIdx.x < n1) {
        const float4 b = ldg4(b1 + (size_t)(r0 + threadIdx.x) * 4);
        s_box[threadIdx.x] = b;
        s_area[threadIdx.x] = area_plus1(b);
    }
    __syncthreads();
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n2) return;
    const float4 c = ldg4(b2 + (size_t)j * 4);
    const float area_c = area_plus1(c);
    const int rows = min(kIouRows, n1 - r0);
    for (int r = 0; r < rows; ++r)
        out[(size_t)(r0 + r) * n2 + j] = iou_plus1(s_box[r], s_area[r], c, area_c);
}

}  // namespace paa

using namespace paa;

namespace paa {

__global__ void selftest_roots_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    out[i] = sqrt_rn_normal(v);
    out[(size_t)n + i] = __fsqrt_rn(v);
    out[2 * (size_t)n + i] = div_rn_normal(1.0f, v);
    out[3 * (size_t)n + i] = __fdiv_rn(1.0f, v);
}

int launch_selftest_roots(const float* x, int n, float* out, cudaStream_t stream) {
    if (n == 0) return 0;
    selftest_roots_kernel<<<(n + 255) / 256, 256, 0, stream>>>(x, n, out);
    PAA_LAUNCH_CHECK("selftest_roots_kernel");
    return 0;
}

}  // namespace paa

extern "C" {

int paa_grid_anchors(const float* cell_anchors, int anchors_per_loc, int grid_h, int grid_w, float stride,
                     float* out_anchors, void* stream_) {
    if (!cell_anchors || !out_anchors || anchors_per_loc < 1 || grid_h < 0 || grid_w < 0) {
        set_error("bad arguments to paa_grid_anchors");
        return PAA_ERR_BAD_ARGUMENT;
    }
    const size_t total = (size_t)grid_h * grid_w * anchors_per_loc;
    if (total == 0) return 0;
    if (reinterpret_cast<uintptr_t>(out_anchors) & 15u) {
        set_error("paa_grid_anchors: out_anchors must be 16-byte aligned");
        return PAA_ERR_BAD_ARGUMENT;
    }
    const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    grid_anchors_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        cell_anchors, anchors_per_loc, grid_h, grid_w, stride, reinterpret_cast<float4*>(out_anchors));
    PAA_LAUNCH_CHECK("grid_anchors_kernel");
    return 0;
}

int paa_anchor_visibility(const float* anchors, int64_t n, float image_w, float image_h, float straddle_thresh,
                          uint8_t* out, void* stream_) {
    if (n < 0 || (n > 0 && (!anchors || !out))) {
        set_error("bad arguments to paa_anchor_visibility");
        return PAA_ERR_BAD_ARGUMENT;
    }
    if (n == 0) return 0;
    if (reinterpret_cast<uintptr_t>(anchors) & 15u) {
        set_error("paa_anchor_visibility: anchors must be 16-byte aligned");
        return PAA_ERR_BAD_ARGUMENT;
    }
    const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    anchor_visibility_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<const float4*>(anchors), (size_t)n, image_w, image_h, straddle_thresh, out);
    PAA_LAUNCH_CHECK("anchor_visibility_kernel");
    return 0;
}

int paa_boxlist_iou(const float* boxes1, int n1, const float* boxes2, int n2, float* out, void* stream_) {
    if (n1 < 0 || n2 < 0 || (n1 > 0 && !boxes1) || (n2 > 0 && !boxes2) || ((size_t)n1 * n2 > 0 && !out)) {
        set_error("bad arguments to paa_boxlist_iou");
        return PAA_ERR_BAD_ARGUMENT;
    }
    if (n1 == 0 || n2 == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(boxes1) & 15u) || (reinterpret_cast<uintptr_t>(boxes2) & 15u)) {
        set_error("paa_boxlist_iou: box arrays must be 16-byte aligned");
        return PAA_ERR_BAD_ARGUMENT;
    }
    const dim3 grid((unsigned)((n2 + 255) / 256), (unsigned)((n1 + kIouRows - 1) / kIouRows));
    if (grid.y > 65535u) {
        set_error("paa_boxlist_iou: n1=%d too large (<= %d)", n1, 65535 * kIouRows);
        return PAA_ERR_UNSUPPORTED;
    }
    boxlist_iou_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(boxes1, n1, boxes2, n2, out);
    PAA_LAUNCH_CHECK("boxlist_iou_kernel");
    return 0;
}

}  // extern "C"
