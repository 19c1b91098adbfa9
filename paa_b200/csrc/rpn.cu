// The plain RPN training loss (paa_core/modeling/rpn/loss.py:98-137) over a GIVEN sample of anchors: binary
// cross-entropy with logits on the sampled anchors' objectness (mean) and smooth-L1 (beta) on the sampled positives'
// regression outputs against BoxCoder targets computed on the fly (box_coder.py:22-50), summed / number of sampled
// anchors; gradients for the sampled entries, zeros elsewhere.  IoU matching and the Matcher labelling in front of it
// are the RetinaNet path's kernels (paa_retinanet_assign with unit labels); the balanced sampler between the two is
// torch.randperm on the host mirror's side, like the reference's.
#include "kernels.h"

namespace paa {

constexpr int kRpnThreads = 1024;

struct RpnSample {
    const long long* idx;     // [n_pos + n_neg] image * A + anchor, positives first
    int n_pos, n_neg;
};

__device__ __forceinline__ float smooth_l1(float x, float t, float beta, float scale, float* grad) {
    const float diff = __fsub_rn(x, t);
    const float n = fabsf(diff);
    if (n < beta) {                                  // layers/smooth_l1_loss.py:10-15
        *grad = __fdiv_rn(diff, beta) * scale;
        return __fdiv_rn(__fmul_rn(__fmul_rn(0.5f, n), n), beta);
    }
    *grad = (diff > 0.0f ? 1.0f : -1.0f) * scale;
    return __fsub_rn(n, __fmul_rn(0.5f, beta));
}

// One block: the sample is a few hundred anchors per image.  Partial sums in double, folded in a fixed order
// (thread-strided accumulation, warp shuffles, warp 0): the losses are reproducible run to run.
__global__ void __launch_bounds__(kRpnThreads)
rpn_loss_kernel(const Geometry geo, const GtOffsets go, const float* __restrict__ gt_boxes,
                const int* __restrict__ matched, const RpnSample smp, const float wx, const float wy, const float ww,
                const float wh, const float beta, const float* __restrict__ gout, float* __restrict__ losses) {
    __shared__ double s_part[kRpnThreads / PAA_WARP][2];
    const int total = smp.n_pos + smp.n_neg;
    const float g_obj = gout ? gout[0] : 1.0f, g_box = gout ? gout[1] : 1.0f;
    const float inv = total > 0 ? 1.0f / (float)total : 0.0f;
    double bce_sum = 0.0, box_sum = 0.0;
    for (int k = threadIdx.x; k < total; k += kRpnThreads) {
        const long long flat = smp.idx[k];
        const int n = (int)(flat / geo.A), a = (int)(flat - (long long)n * geo.A);
        const int l = anchor_level(geo, a);
        const LevelView& lv = geo.lv[l];
        const int i = a - lv.a_off;
        const bool positive = k < smp.n_pos;
        // objectness: F.binary_cross_entropy_with_logits = max(x, 0) - x * y + log(1 + exp(-|x|))
        const size_t o_off = head_offset(geo, lv, n, i, 0, 1);
        const float x = __ldg(lv.cls + o_off);
        const float y = positive ? 1.0f : 0.0f;
        const float e = expf(-fabsf(x));
        bce_sum += (double)(fmaxf(x, 0.0f) - x * y + log1pf(e));
        if (lv.g_cls) {
            const float sig = x >= 0.0f ? 1.0f / (1.0f + e) : e / (1.0f + e);
            lv.g_cls[o_off] = (sig - y) * inv * g_obj;
        }
        if (positive) {
            const int m = max(matched[flat], 0);          // rpn/loss.py:51: matched_idxs.clamp(min=0)
            const float4 anc = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
            const float4 gt = ldg4(gt_boxes + (size_t)(go.v[n] + m) * 4);
            const float4 t = encode_box_legacy(gt, anc, wx, wy, ww, wh);
            const float* rp = lv.reg + head_offset(geo, lv, n, i, 0, 4);
            const float4 d = load_channels4(rp, head_cstride(geo, lv));
            float4 gd;
            float s = 0.f;
            s += smooth_l1(d.x, t.x, beta, inv * g_box, &gd.x);
            s += smooth_l1(d.y, t.y, beta, inv * g_box, &gd.y);
            s += smooth_l1(d.z, t.z, beta, inv * g_box, &gd.z);
            s += smooth_l1(d.w, t.w, beta, inv * g_box, &gd.w);
            box_sum += (double)s;
            if (lv.g_reg) store_channels4(lv.g_reg + head_offset(geo, lv, n, i, 0, 4), head_cstride(geo, lv), gd);
        }
    }
    bce_sum = warp_sum(bce_sum);
    box_sum = warp_sum(box_sum);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_part[warp][0] = bce_sum;
        s_part[warp][1] = box_sum;
    }
    __syncthreads();
    if (warp == 0) {
        double a0 = s_part[lane][0], a1 = s_part[lane][1];
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        if (lane == 0) {
            // an empty sample: the reference's mean over no element is NaN, its 0 / 0 box loss as well
            losses[0] = total > 0 ? (float)(a0 / (double)total) : __int_as_float(0x7fc00000);
            losses[1] = total > 0 ? (float)(a1 / (double)total) : __int_as_float(0x7fc00000);
        }
    }
}

int launch_rpn_loss(const Geometry& geo, const GtOffsets& go, const float* gt_boxes, const int* matched,
                    const long long* sampled, int n_pos, int n_neg, const float* weights, float beta,
                    const float* grad_losses, float* losses, cudaStream_t stream) {
    // gradients: zeros everywhere but on the sampled anchors
    for (int l = 0; l < geo.num_levels; ++l) {
        const LevelView& lv = geo.lv[l];
        const size_t per = (size_t)geo.num_images * lv.n_anchor * sizeof(float);
        if (lv.g_cls) PAA_CUDA_CHECK(cudaMemsetAsync(lv.g_cls, 0, per, stream));
        if (lv.g_reg) PAA_CUDA_CHECK(cudaMemsetAsync(lv.g_reg, 0, per * 4, stream));
    }
    RpnSample smp;
    smp.idx = sampled;
    smp.n_pos = n_pos;
    smp.n_neg = n_neg;
    rpn_loss_kernel<<<1, kRpnThreads, 0, stream>>>(geo, go, gt_boxes, matched, smp, weights[0], weights[1], weights[2],
                                                   weights[3], beta, grad_losses, losses);
    PAA_LAUNCH_CHECK("rpn_loss_kernel");
    return 0;
}

}  // namespace paa
