// Final PAA losses and their gradients on sm_100a, one streaming pass over the head outputs.
//
// Replaces loss.py:317-358 of the reference plus the autograd backward of the returned losses:
//   cls  = sum over all anchors x classes of sigmoid focal loss against the PAA labels
//          (layers/sigmoid_focal_loss.py:40-52, csrc/cuda/SigmoidFocalLoss_cuda.cu:20-101) / num_pos_avg
//   reg  = sum over positives of GIoU loss (loss.py:46-87) weighted by the detached IoU between the
//          decoded prediction and the decoded target (loss.py:331-341) / sum_iou_avg * REG_LOSS_WEIGHT
//   iou  = BCE-with-logits(iou_pred[pos], IoU) summed / num_pos_avg * IOU_LOSS_WEIGHT
// The kernel reads each logit once (NCHW, coalesced over consecutive anchors), and writes the
// gradient of the same element in the same pass, already divided by the all-reduced normalisers
// that paa_assign left in device memory (no host round trip for .item(), loss.py:321,338).
#include "kernels.h"

namespace paa {

constexpr int kFinalThreads = PAA_TILE;

// upper bound of blocks the final kernel launches (sizes the partial-sum buffer)
int loss_grid_blocks(int num_images, int tiles_per_image) { (void)num_images; (void)tiles_per_image; return 148 * 16; }

struct GradScales {
    float cls, reg, bce;     // d(total)/d(sum) factors
    bool weighted;           // GIoU loss weighted by IoU (loss.py:83: only if the weights sum > 0)
};

__device__ __forceinline__ GradScales make_scales(const LossScalars& sc, const double* __restrict__ norm,
                                                  const double* __restrict__ local_norm,
                                                  const float* __restrict__ gout) {
    // num_pos is an exact integer in float (< 2^24); the sums are float32 in the reference as well
    const float world = (float)sc.world_size;
    const float num_pos_avg = fmaxf((float)norm[0] / world, 1.0f);     // loss.py:322
    const float g0 = gout ? gout[0] : 1.0f, g1 = gout ? gout[1] : 1.0f, g2 = gout ? gout[2] : 1.0f;
    GradScales s;
    s.cls = g0 / num_pos_avg;
    if (sc.use_iou_pred) {
        const float reg_norm = (float)norm[1] / world;                 // loss.py:338,354
        s.reg = g1 * sc.reg_loss_weight / reg_norm;
        s.bce = g2 * sc.iou_loss_weight / num_pos_avg;
        s.weighted = local_norm[1] > 0.0;
    } else {
        s.reg = g1 * sc.reg_loss_weight / num_pos_avg;
        s.bce = 0.0f;
        s.weighted = false;
    }
    return s;
}

// torch's sub-gradient for max(a, b) / min(a, b): the whole gradient to the selected operand, half
// to each on a tie.
__device__ __forceinline__ float pick_first(float a, float b, bool want_max) {
    if (a == b) return 0.5f;
    return ((a > b) == want_max) ? 1.0f : 0.0f;
}

// GIoU loss of (decoded prediction p, target t) and d(loss)/d(regression deltas).
__device__ __forceinline__ float giou_loss_and_grad(float4 d, const AnchorFrame& f, float4 t, float4* grad_d) {
    float pw, ph;
    bool pass_w, pass_h;
    const float4 p = decode_box(d, f, &pw, &ph, &pass_w, &pass_h);
    const float loss = giou_loss_boxes(p, t);
    // forward intermediates (plain float math is fine for the gradient)
    const float px1 = p.x, py1 = p.y;
    const float px2 = fmaxf(p.x, p.z), py2 = fmaxf(p.y, p.w);
    const float p_area = (px2 - px1) * (py2 - py1);
    const float t_area = (t.z - t.x) * (t.w - t.y);
    const float ix1 = fmaxf(px1, t.x), iy1 = fmaxf(py1, t.y), ix2 = fminf(px2, t.z), iy2 = fminf(py2, t.w);
    const bool overlap = (iy2 > iy1) && (ix2 > ix1);
    const float inter = overlap ? (ix2 - ix1) * (iy2 - iy1) : 0.0f;
    const float ex1 = fminf(px1, t.x), ey1 = fminf(py1, t.y), ex2 = fmaxf(px2, t.z), ey2 = fmaxf(py2, t.w);
    const float enc = (ex2 - ex1) * (ey2 - ey1) + 1e-7f;
    const float uni = p_area + t_area - inter + 1e-7f;
    // backward of loss = 1 - (inter/uni - (enc - uni)/enc)
    const float g_giou = -1.0f;
    const float g_uni = g_giou * (1.0f / enc) + g_giou * (-inter / (uni * uni));
    const float g_enc = g_giou * (-uni / (enc * enc));
    const float g_inter = g_giou / uni - g_uni;
    const float g_parea = g_uni;
    float gx1 = 0.f, gy1 = 0.f, gx2m = 0.f, gy2m = 0.f;    // grads of px1, py1, px2(max'ed), py2(max'ed)
    gx2m += g_parea * (py2 - py1);
    gx1 -= g_parea * (py2 - py1);
    gy2m += g_parea * (px2 - px1);
    gy1 -= g_parea * (px2 - px1);
    if (overlap) {
        const float gw = g_inter * (iy2 - iy1), gh = g_inter * (ix2 - ix1);
        gx2m += gw * pick_first(px2, t.z, false);
        gx1 -= gw * pick_first(px1, t.x, true);
        gy2m += gh * pick_first(py2, t.w, false);
        gy1 -= gh * pick_first(py1, t.y, true);
    }
    {
        const float gw = g_enc * (ey2 - ey1), gh = g_enc * (ex2 - ex1);
        gx2m += gw * pick_first(px2, t.z, true);
        gx1 -= gw * pick_first(px1, t.x, false);
        gy2m += gh * pick_first(py2, t.w, true);
        gy1 -= gh * pick_first(py1, t.y, false);
    }
    // px2 = max(x1, x2raw), py2 = max(y1, y2raw)  (loss.py:52-53)
    const float sx = pick_first(p.x, p.z, true), sy = pick_first(p.y, p.w, true);
    const float gx2 = gx2m * (1.0f - sx), gy2 = gy2m * (1.0f - sy);
    gx1 += gx2m * sx;
    gy1 += gy2m * sy;
    // decode backward (atss.py:78-96)
    const float g_pcx = gx1 + gx2, g_pcy = gy1 + gy2;
    const float g_pw = 0.5f * (gx2 - gx1), g_ph = 0.5f * (gy2 - gy1);
    grad_d->x = g_pcx * f.w / 10.0f;
    grad_d->y = g_pcy * f.h / 10.0f;
    grad_d->z = pass_w ? g_pw * pw / 5.0f : 0.0f;
    grad_d->w = pass_h ? g_ph * ph / 5.0f : 0.0f;
    return loss;
}

// Work item = 128 consecutive anchors of one level of one image x one chunk of kClsChunk classes;
// one thread = one anchor.  Blocks are persistent (grid = SM count x resident blocks) and walk the
// items round-robin, so partial sums are reduced once per block and the item -> block map is static
// (bit-reproducible sums).  Every logit of a chunk is read once (a warp reads one 128-byte line per
// class) and its gradient written once.  All classes are first treated as negatives; the single
// positive class of a positive anchor is patched afterwards.  The chunk-0 item of a tile also handles
// the regression / IoU-prediction losses and gradients of its anchors.
constexpr int kClsChunk = 16;
constexpr int kClsBatch = 8;
constexpr int kFinalBlocksPerSM = 8;

int loss_class_chunks(int C) { return (C + kClsChunk - 1) / kClsChunk; }

// negative-class focal term without its (1-alpha) factor, and the gradient with `k` = (1-alpha) * scale
template <bool kG2>
__device__ __forceinline__ void neg_term_grad(float x, float gamma, float k, float* sum, float* grad) {
    const SigmoidParts s = sigmoid_parts(x);
    const float nlogq = fmaxf(x, 0.0f) + s.l1p;
    const float mod = kG2 ? s.p * s.p : __powf(s.p, gamma);
    *sum = fmaf(mod, nlogq, *sum);
    const float gq = kG2 ? s.q + s.q : gamma * s.q;
    *grad = (mod * fmaf(gq, nlogq, s.p)) * k;
}

template <bool kGrads, bool kG2>
__global__ void __launch_bounds__(kFinalThreads, kFinalBlocksPerSM)
final_loss_kernel(const Geometry geo, const GtOffsets go, const float* __restrict__ gt_boxes,
                  const LossScalars sc, const int* __restrict__ paa_label, const int* __restrict__ matched,
                  const double* __restrict__ norm, const double* __restrict__ local_norm,
                  const float* __restrict__ gout, double* __restrict__ block_part, const int n_chunks,
                  const int n_items) {
    __shared__ double s_part[kFinalThreads / PAA_WARP][3];
    const GradScales gs = make_scales(sc, norm, local_norm, gout);
    const float alpha = sc.alpha, gamma = sc.gamma, oma = 1.0f - sc.alpha;
    const float kneg = oma * gs.cls;
    float neg_sum = 0.f, fix_sum = 0.f, reg_sum = 0.f, bce_sum = 0.f;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int chunk = item % n_chunks;
        const int bt = item / n_chunks;
        const int n = bt / geo.tiles_per_image;
        const int tile = bt - n * geo.tiles_per_image;
        int first;
        const int l = tile_level(geo, tile, &first);
        const LevelView& lv = geo.lv[l];
        const int i = first + threadIdx.x;
        if (i >= lv.n_anchor) continue;
        const int c_begin = chunk * kClsChunk;
        const int c_end = min(geo.C, c_begin + kClsChunk);
        const size_t flat = (size_t)n * geo.A + lv.a_off + i;
        const int label = __ldg(paa_label + flat);
        const size_t off = head_offset(n, i, c_begin, geo.C, geo.apl, lv.hw);
        const float* __restrict__ cls = lv.cls + off;
        float* __restrict__ gcls = lv.g_cls ? lv.g_cls + off : nullptr;
        const unsigned stride = (unsigned)lv.hw;
        const bool write = kGrads && gcls != nullptr;
        if (c_end - c_begin == kClsChunk) {
#pragma unroll
            for (int b0 = 0; b0 < kClsChunk; b0 += kClsBatch) {
                float x[kClsBatch];
#pragma unroll
                for (int j = 0; j < kClsBatch; ++j) x[j] = __ldg(cls + (unsigned)(b0 + j) * stride);
#pragma unroll
                for (int j = 0; j < kClsBatch; ++j) {
                    float g;
                    neg_term_grad<kG2>(x[j], gamma, kneg, &neg_sum, &g);
                    if (write) gcls[(unsigned)(b0 + j) * stride] = g;
                }
            }
        } else {
            for (int c = 0; c < c_end - c_begin; ++c) {
                float g;
                neg_term_grad<kG2>(__ldg(cls + (unsigned)c * stride), gamma, kneg, &neg_sum, &g);
                if (write) gcls[(unsigned)c * stride] = g;
            }
        }
        if (label > 0 && label - 1 >= c_begin && label - 1 < c_end) {
            const unsigned po = (unsigned)(label - 1 - c_begin) * stride;
            const float xp = __ldg(cls + po);
            const SigmoidParts sp = sigmoid_parts(xp);
            float tn, gn, tp, gp;
            focal_negative(xp, sp, gamma, kG2, oma, &tn, &gn);
            focal_positive(xp, sp, gamma, kG2, alpha, &tp, &gp);
            fix_sum += tp - tn;
            if (write) gcls[po] = gp * gs.cls;
        }

        if (chunk == 0) {
            float4 gd = make_float4(0.f, 0.f, 0.f, 0.f);
            float giou_g = 0.f;
            if (label > 0) {
                const int m = matched[flat];
                const float4 a = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
                const AnchorFrame f = anchor_frame(a);
                const float* rp = lv.reg + head_offset(n, i, 0, 4, geo.apl, lv.hw);
                const float4 d = make_float4(__ldg(rp), __ldg(rp + lv.hw), __ldg(rp + 2 * (size_t)lv.hw),
                                             __ldg(rp + 3 * (size_t)lv.hw));
                const float4 gt = ldg4(gt_boxes + (size_t)(go.v[n] + m) * 4);
                const float4 tgt = decode_box(encode_box(gt, f), f);
                float w = 1.0f;
                if (sc.use_iou_pred) {
                    const float4 pred = decode_box(d, f);
                    const float q = iou_plus1(tgt, area_plus1(tgt), pred, area_plus1(pred));
                    const float xi = __ldg(lv.iou + head_offset(n, i, 0, 1, geo.apl, lv.hw));
                    const float ei = expf(-fabsf(xi));
                    bce_sum += fmaxf(xi, 0.0f) - xi * q + log1pf(ei);
                    const float sig = (xi >= 0.0f) ? 1.0f / (1.0f + ei) : ei / (1.0f + ei);
                    giou_g = (sig - q) * gs.bce;
                    if (gs.weighted) w = q;
                }
                float4 gdd;
                const float gl = giou_loss_and_grad(d, f, tgt, &gdd);
                reg_sum += gl * w;
                const float k = w * gs.reg;
                gd = make_float4(gdd.x * k, gdd.y * k, gdd.z * k, gdd.w * k);
            }
            if (kGrads) {
                if (lv.g_reg) {
                    float* gr = lv.g_reg + head_offset(n, i, 0, 4, geo.apl, lv.hw);
                    gr[0] = gd.x;
                    gr[lv.hw] = gd.y;
                    gr[2 * (size_t)lv.hw] = gd.z;
                    gr[3 * (size_t)lv.hw] = gd.w;
                }
                if (lv.g_iou) lv.g_iou[head_offset(n, i, 0, 1, geo.apl, lv.hw)] = giou_g;
            }
        }
    }
    // block partial sums (double, fixed order)
    const float cls_sum = fmaf(oma, neg_sum, fix_sum);
    double a0 = warp_sum((double)cls_sum), a1 = warp_sum((double)reg_sum), a2 = warp_sum((double)bce_sum);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_part[warp][0] = a0;
        s_part[warp][1] = a1;
        s_part[warp][2] = a2;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kFinalThreads / PAA_WARP; ++w) t += s_part[w][threadIdx.x];
        block_part[(size_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
}

// Folds the per-block partials in a fixed order and applies the normalisers (loss.py:354-358).
__global__ void __launch_bounds__(256)
finish_loss_kernel(const double* __restrict__ block_part, int blocks, const LossScalars sc,
                   const double* __restrict__ norm, float* __restrict__ losses) {
    __shared__ double s[8][3];
    double a[3] = {0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < blocks; b += 256) {
        a[0] += block_part[(size_t)b * 3 + 0];
        a[1] += block_part[(size_t)b * 3 + 1];
        a[2] += block_part[(size_t)b * 3 + 2];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a[k] = warp_sum(a[k]);
        if (lane == 0) s[warp][k] = a[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[3] = {0.0, 0.0, 0.0};
        for (int w = 0; w < 8; ++w)
            for (int k = 0; k < 3; ++k) t[k] += s[w][k];
        const double world = (double)sc.world_size;
        const float num_pos_avg = (float)fmax(norm[0] / world, 1.0);
        losses[0] = (float)t[0] / num_pos_avg;
        if (sc.use_iou_pred) {
            const float reg_norm = (float)(norm[1] / world);
            losses[1] = (float)t[1] / reg_norm * sc.reg_loss_weight;
            losses[2] = (float)t[2] / num_pos_avg * sc.iou_loss_weight;
        } else {
            losses[1] = (float)t[1] / num_pos_avg * sc.reg_loss_weight;
            losses[2] = 0.0f;
        }
    }
}

int launch_final_loss(const Geometry& geo, const GtOffsets& go, const float* gt_boxes,
                      const LossScalars& sc, const LossWorkspace& ws, const double* normalisers,
                      const float* grad_losses, float* losses, bool write_grads, cudaStream_t stream) {
    const int n_chunks = loss_class_chunks(geo.C);
    const int n_items = geo.num_images * geo.tiles_per_image * n_chunks;
    int grid = 148 * kFinalBlocksPerSM;
    if (grid > n_items) grid = n_items;
    {
        KernelTimer timer(PAA_KERNEL_FINAL_LOSS, stream);
        const bool g2 = (sc.gamma == 2.0f);
#define PAA_FINAL(G, T)                                                                                \
    final_loss_kernel<G, T><<<grid, kFinalThreads, 0, stream>>>(geo, go, gt_boxes, sc, ws.paa_label,   \
        ws.matched, normalisers, ws.local_norm, grad_losses, ws.block_part, n_chunks, n_items)
        if (write_grads) {
            if (g2) PAA_FINAL(true, true); else PAA_FINAL(true, false);
        } else {
            if (g2) PAA_FINAL(false, true); else PAA_FINAL(false, false);
        }
#undef PAA_FINAL
    }
    PAA_LAUNCH_CHECK("final_loss_kernel");
    finish_loss_kernel<<<1, 256, 0, stream>>>(ws.block_part, grid, sc, normalisers, losses);
    PAA_LAUNCH_CHECK("finish_loss_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// grad *= new/old per loss, for a backward() whose upstream gradients differ from the ones the
// fused forward assumed.
// ---------------------------------------------------------------------------------------------
struct RescaleJob {
    float* ptr[3 * PAA_MAX_LEVELS];
    unsigned long long count[3 * PAA_MAX_LEVELS];
    int which[3 * PAA_MAX_LEVELS];
    int n;
};

// One launch for all gradient tensors; returns at once when every ratio is 1 (the usual
// `sum(losses).backward()`).
__global__ void __launch_bounds__(256)
rescale_kernel(const RescaleJob job, const float* __restrict__ old_g, const float* __restrict__ new_g) {
    float r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = new_g[k] / old_g[k];
    if (r[0] == 1.0f && r[1] == 1.0f && r[2] == 1.0f) return;
    for (int j = 0; j < job.n; ++j) {
        const float rr = r[job.which[j]];
        if (rr == 1.0f) continue;
        float* p = job.ptr[j];
        const size_t count = job.count[j];
        for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < count;
             k += (size_t)gridDim.x * blockDim.x)
            p[k] *= rr;
    }
}

int launch_rescale_grads(const Geometry& geo, const float* old_g, const float* new_g, cudaStream_t stream) {
    RescaleJob job;
    job.n = 0;
    for (int l = 0; l < geo.num_levels; ++l) {
        const LevelView& lv = geo.lv[l];
        const size_t per = (size_t)geo.num_images * geo.apl * lv.hw;
        float* ptr[3] = {lv.g_cls, lv.g_reg, lv.g_iou};
        const size_t cnt[3] = {per * geo.C, per * 4, per};
        for (int k = 0; k < 3; ++k) {
            if (!ptr[k]) continue;
            job.ptr[job.n] = ptr[k];
            job.count[job.n] = cnt[k];
            job.which[job.n] = k;
            ++job.n;
        }
    }
    if (job.n == 0) return 0;
    rescale_kernel<<<148 * 4, 256, 0, stream>>>(job, old_g, new_g);
    PAA_LAUNCH_CHECK("rescale_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Stand-alone sigmoid focal loss on [n, C] row-major logits: the `_C.sigmoid_focalloss_forward`
// / `_backward` entry points (csrc/SigmoidFocalLoss.h:10-41).
// ---------------------------------------------------------------------------------------------
__global__ void focal_rowmajor_kernel(const float* __restrict__ logits, const int* __restrict__ targets,
                                      const float* __restrict__ d_losses, size_t total, int C, float gamma,
                                      float alpha, float* __restrict__ out, int backward) {
    const bool g2 = (gamma == 2.0f);
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total;
         k += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(k / C), c = (int)(k - (size_t)row * C);
        const int t = __ldg(targets + row);
        const float x = __ldg(logits + k);
        const float e = expf(-fabsf(x));
        const float l1p = log1pf(e);
        const float inv = 1.0f / (1.0f + e);
        const float pr = (x >= 0.0f) ? inv : e * inv;
        const float qr = (x >= 0.0f) ? e * inv : inv;
        float v = 0.0f;
        if (t == c + 1) {
            const float nlogp = fmaxf(-x, 0.0f) + l1p;
            const float mod = g2 ? qr * qr : powf(qr, gamma);
            v = backward ? -alpha * mod * (qr + gamma * pr * nlogp) : alpha * mod * nlogp;
        } else if (t >= 0) {
            const float nlogq = fmaxf(x, 0.0f) + l1p;
            const float mod = g2 ? pr * pr : powf(pr, gamma);
            v = backward ? (1.0f - alpha) * mod * (gamma * qr * nlogq + pr) : (1.0f - alpha) * mod * nlogq;
        }
        if (backward) v *= __ldg(d_losses + k);
        out[k] = v;
    }
}

static int focal_grid(size_t total) {
    size_t g = (total + 255) / 256;
    const size_t cap = 148 * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

int launch_focal_forward(const float* logits, const int* targets, int n, int C, float gamma, float alpha,
                         float* losses, cudaStream_t stream) {
    const size_t total = (size_t)n * C;
    if (total == 0) return 0;
    focal_rowmajor_kernel<<<focal_grid(total), 256, 0, stream>>>(logits, targets, nullptr, total, C, gamma,
                                                                alpha, losses, 0);
    PAA_LAUNCH_CHECK("focal_rowmajor_kernel(forward)");
    return 0;
}

int launch_focal_backward(const float* logits, const int* targets, const float* d_losses, int n, int C,
                          float gamma, float alpha, float* d_logits, cudaStream_t stream) {
    const size_t total = (size_t)n * C;
    if (total == 0) return 0;
    focal_rowmajor_kernel<<<focal_grid(total), 256, 0, stream>>>(logits, targets, d_losses, total, C, gamma,
                                                                alpha, d_logits, 1);
    PAA_LAUNCH_CHECK("focal_rowmajor_kernel(backward)");
    return 0;
}

}  // namespace paa
