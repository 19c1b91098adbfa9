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

// upper bound of blocks the final kernel launches (sizes the partial-sum buffer): one per (128-anchor run,
// class chunk); 64 chunks cover C <= 1024
int loss_grid_blocks(int num_images, int tiles_per_image) { return num_images * tiles_per_image * 64; }

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

// Work item = a run of consecutive anchors of one level of one image x one chunk of kClsChunk classes.
// On levels whose H*W is a multiple of 4 a thread owns 4 consecutive anchors and moves float4s (one
// address computation and one LDG.128 / STG.128 per 4 logits); other levels (the small coarse ones) use
// one anchor per thread.  One block per item, largest items first; every block leaves its partial sums
// in its own slot, so the final fold is order-fixed and bit-reproducible.  Every logit of a chunk is
// read once and its gradient written once.  All classes are first treated as negatives; the single
// positive class of a positive anchor is patched afterwards.  The chunk-0 item of a run also handles the
// regression / IoU-prediction losses and gradients of its anchors.
constexpr int kClsChunk = 16;
constexpr int kVecBatch = 4;       // classes in flight per thread on the float4 path (4 x 16 B)
constexpr int kClsBatch = 8;       // classes in flight per thread on the scalar path
constexpr int kFinalBlocksPerSM = 6;

int loss_class_chunks(int C) { return (C + kClsChunk - 1) / kClsChunk; }

struct FinalPlan {
    int item_off[PAA_MAX_LEVELS + 1];   // per image: first item of each level (items = runs x chunks)
    int vec[PAA_MAX_LEVELS];            // 1: float4 path
    int n_chunks;
    int items_per_image;
};

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

struct FinalCtx {
    float alpha, gamma, oma, kneg;
    GradScales gs;
};

// swaps the (already written) negative-class result of the labelled class for the positive-class one
template <bool kGrads, bool kG2>
__device__ __forceinline__ void patch_positive(const float* __restrict__ cls, float* __restrict__ gcls,
                                               unsigned elem_off, const FinalCtx& cx, float* fix_sum) {
    const float xp = __ldg(cls + elem_off);
    const SigmoidParts sp = sigmoid_parts(xp);
    float tn, gn, tp, gp;
    focal_negative(xp, sp, cx.gamma, kG2, cx.oma, &tn, &gn);
    focal_positive(xp, sp, cx.gamma, kG2, cx.alpha, &tp, &gp);
    *fix_sum += tp - tn;
    if (kGrads && gcls) gcls[elem_off] = gp * cx.gs.cls;
}

// regression + IoU-prediction losses / gradients of one positive anchor (loss.py:328-349)
__device__ __forceinline__ void positive_box_terms(const Geometry& geo, const LevelView& lv, const GtOffsets& go,
                                                   const float* __restrict__ gt_boxes, const LossScalars& sc,
                                                   const FinalCtx& cx, int n, int i, int m, float4 d, float xi,
                                                   float* reg_sum, float* bce_sum, float4* gd, float* gi) {
    const float4 a = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
    const AnchorFrame f = anchor_frame(a);
    const float4 gt = ldg4(gt_boxes + (size_t)(go.v[n] + m) * 4);
    const float4 tgt = decode_box(encode_box(gt, f), f);
    float w = 1.0f;
    if (sc.use_iou_pred) {
        const float4 pred = decode_box(d, f);
        const float q = iou_plus1(tgt, area_plus1(tgt), pred, area_plus1(pred));
        const float ei = expf(-fabsf(xi));
        *bce_sum += fmaxf(xi, 0.0f) - xi * q + log1pf(ei);
        const float sig = (xi >= 0.0f) ? 1.0f / (1.0f + ei) : ei / (1.0f + ei);
        *gi = (sig - q) * cx.gs.bce;
        if (cx.gs.weighted) w = q;
    }
    float4 gdd;
    const float gl = giou_loss_and_grad(d, f, tgt, &gdd);
    *reg_sum += gl * w;
    const float k = w * cx.gs.reg;
    *gd = make_float4(gdd.x * k, gdd.y * k, gdd.z * k, gdd.w * k);
}

template <bool kGrads, bool kG2>
__global__ void __launch_bounds__(kFinalThreads, kFinalBlocksPerSM)
final_loss_kernel(const Geometry geo, const GtOffsets go, const FinalPlan plan,
                  const float* __restrict__ gt_boxes, const LossScalars sc,
                  const int* __restrict__ paa_label, const int* __restrict__ matched,
                  const double* __restrict__ norm, const double* __restrict__ local_norm,
                  const float* __restrict__ gout, double* __restrict__ block_part) {
    __shared__ double s_part[kFinalThreads / PAA_WARP][3];
    FinalCtx cx;
    cx.gs = make_scales(sc, norm, local_norm, gout);
    cx.alpha = sc.alpha;
    cx.gamma = sc.gamma;
    cx.oma = 1.0f - sc.alpha;
    cx.kneg = cx.oma * cx.gs.cls;
    float neg_sum = 0.f, fix_sum = 0.f, reg_sum = 0.f, bce_sum = 0.f;

    // item -> (image, level, run, chunk); all images' large items come first
    const int q = blockIdx.x / geo.num_images;
    const int n = blockIdx.x - q * geo.num_images;
    int l = 0;
#pragma unroll 1
    for (int k = 1; k < geo.num_levels; ++k)
        if (q >= plan.item_off[k]) l = k;
    const int local = q - plan.item_off[l];
    const int chunk = local % plan.n_chunks;
    const int run = local / plan.n_chunks;
    const LevelView& lv = geo.lv[l];
    const int c_begin = chunk * kClsChunk;
    const int c_end = min(geo.C, c_begin + kClsChunk);
    const unsigned stride = (unsigned)lv.hw;
    const size_t img_flat = (size_t)n * geo.A + lv.a_off;

    if (plan.vec[l]) {
        // ---- float4 path: 4 consecutive anchors per thread (apl == 1, hw % 4 == 0) ----------------
        const int i0 = (run * kFinalThreads + threadIdx.x) * 4;
        if (i0 < lv.n_anchor) {
            const size_t off = ((size_t)n * geo.C + c_begin) * stride + i0;
            const float4* __restrict__ cls4 = reinterpret_cast<const float4*>(lv.cls + off);
            float4* __restrict__ g4 = lv.g_cls ? reinterpret_cast<float4*>(lv.g_cls + off) : nullptr;
            const unsigned stride4 = stride >> 2;
            const bool write = kGrads && g4 != nullptr;
            const int nc = c_end - c_begin;
            for (int b0 = 0; b0 < nc; b0 += kVecBatch) {
                float4 x[kVecBatch];
#pragma unroll
                for (int j = 0; j < kVecBatch; ++j)
                    x[j] = (b0 + j < nc) ? __ldg(cls4 + (unsigned)(b0 + j) * stride4)
                                         : make_float4(-100.f, -100.f, -100.f, -100.f);
#pragma unroll
                for (int j = 0; j < kVecBatch; ++j) {
                    float4 g;
                    neg_term_grad<kG2>(x[j].x, cx.gamma, cx.kneg, &neg_sum, &g.x);
                    neg_term_grad<kG2>(x[j].y, cx.gamma, cx.kneg, &neg_sum, &g.y);
                    neg_term_grad<kG2>(x[j].z, cx.gamma, cx.kneg, &neg_sum, &g.z);
                    neg_term_grad<kG2>(x[j].w, cx.gamma, cx.kneg, &neg_sum, &g.w);
                    if (write && b0 + j < nc) g4[(unsigned)(b0 + j) * stride4] = g;
                }
            }
            int label[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) label[k] = __ldg(paa_label + img_flat + i0 + k);
            const float* cls = lv.cls + off;
            float* gcls = lv.g_cls ? lv.g_cls + off : nullptr;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (label[k] > 0 && label[k] - 1 >= c_begin && label[k] - 1 < c_end)
                    patch_positive<kGrads, kG2>(cls, gcls, (unsigned)(label[k] - 1 - c_begin) * stride + k, cx,
                                                &fix_sum);
            if (chunk == 0) {
                float4 gx = make_float4(0.f, 0.f, 0.f, 0.f), gy = gx, gw = gx, gh = gx, gi = gx;
                if (label[0] > 0 || label[1] > 0 || label[2] > 0 || label[3] > 0) {
                    const float* rp = lv.reg + (size_t)n * 4 * stride + i0;
                    const float4 dx = __ldg(reinterpret_cast<const float4*>(rp));
                    const float4 dy = __ldg(reinterpret_cast<const float4*>(rp + stride));
                    const float4 dw = __ldg(reinterpret_cast<const float4*>(rp + 2 * (size_t)stride));
                    const float4 dh = __ldg(reinterpret_cast<const float4*>(rp + 3 * (size_t)stride));
                    float4 xi = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (sc.use_iou_pred) xi = __ldg(reinterpret_cast<const float4*>(lv.iou + (size_t)n * stride + i0));
                    const float dxs[4] = {dx.x, dx.y, dx.z, dx.w}, dys[4] = {dy.x, dy.y, dy.z, dy.w};
                    const float dws[4] = {dw.x, dw.y, dw.z, dw.w}, dhs[4] = {dh.x, dh.y, dh.z, dh.w};
                    const float xis[4] = {xi.x, xi.y, xi.z, xi.w};
                    float ox[4] = {0.f, 0.f, 0.f, 0.f}, oy[4] = {0.f, 0.f, 0.f, 0.f}, ow[4] = {0.f, 0.f, 0.f, 0.f},
                          oh[4] = {0.f, 0.f, 0.f, 0.f}, oi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (label[k] > 0) {
                            float4 gd = make_float4(0.f, 0.f, 0.f, 0.f);
                            float g1 = 0.f;
                            positive_box_terms(geo, lv, go, gt_boxes, sc, cx, n, i0 + k,
                                               matched[img_flat + i0 + k],
                                               make_float4(dxs[k], dys[k], dws[k], dhs[k]), xis[k], &reg_sum,
                                               &bce_sum, &gd, &g1);
                            ox[k] = gd.x;
                            oy[k] = gd.y;
                            ow[k] = gd.z;
                            oh[k] = gd.w;
                            oi[k] = g1;
                        }
                    }
                    gx = make_float4(ox[0], ox[1], ox[2], ox[3]);
                    gy = make_float4(oy[0], oy[1], oy[2], oy[3]);
                    gw = make_float4(ow[0], ow[1], ow[2], ow[3]);
                    gh = make_float4(oh[0], oh[1], oh[2], oh[3]);
                    gi = make_float4(oi[0], oi[1], oi[2], oi[3]);
                }
                if (kGrads) {
                    if (lv.g_reg) {
                        float* gr = lv.g_reg + (size_t)n * 4 * stride + i0;
                        *reinterpret_cast<float4*>(gr) = gx;
                        *reinterpret_cast<float4*>(gr + stride) = gy;
                        *reinterpret_cast<float4*>(gr + 2 * (size_t)stride) = gw;
                        *reinterpret_cast<float4*>(gr + 3 * (size_t)stride) = gh;
                    }
                    if (lv.g_iou) *reinterpret_cast<float4*>(lv.g_iou + (size_t)n * stride + i0) = gi;
                }
            }
        }
    } else {
        // ---- scalar path: one anchor per thread ----------------------------------------------------
        const int i = run * kFinalThreads + threadIdx.x;
        if (i < lv.n_anchor) {
            const size_t flat = img_flat + i;
            const int label = __ldg(paa_label + flat);
            const size_t off = head_offset(n, i, c_begin, geo.C, geo.apl, lv.hw);
            const float* __restrict__ cls = lv.cls + off;
            float* __restrict__ gcls = lv.g_cls ? lv.g_cls + off : nullptr;
            const bool write = kGrads && gcls != nullptr;
            const int nc = c_end - c_begin;
            for (int b0 = 0; b0 < nc; b0 += kClsBatch) {
                float x[kClsBatch];
#pragma unroll
                for (int j = 0; j < kClsBatch; ++j)
                    x[j] = (b0 + j < nc) ? __ldg(cls + (unsigned)(b0 + j) * stride) : -100.0f;
#pragma unroll
                for (int j = 0; j < kClsBatch; ++j) {
                    float g;
                    neg_term_grad<kG2>(x[j], cx.gamma, cx.kneg, &neg_sum, &g);
                    if (write && b0 + j < nc) gcls[(unsigned)(b0 + j) * stride] = g;
                }
            }
            if (label > 0 && label - 1 >= c_begin && label - 1 < c_end)
                patch_positive<kGrads, kG2>(cls, gcls, (unsigned)(label - 1 - c_begin) * stride, cx, &fix_sum);
            if (chunk == 0) {
                float4 gd = make_float4(0.f, 0.f, 0.f, 0.f);
                float gi = 0.f;
                if (label > 0) {
                    const float* rp = lv.reg + head_offset(n, i, 0, 4, geo.apl, lv.hw);
                    const float4 d = make_float4(__ldg(rp), __ldg(rp + lv.hw), __ldg(rp + 2 * (size_t)lv.hw),
                                                 __ldg(rp + 3 * (size_t)lv.hw));
                    const float xi = sc.use_iou_pred ? __ldg(lv.iou + head_offset(n, i, 0, 1, geo.apl, lv.hw)) : 0.f;
                    positive_box_terms(geo, lv, go, gt_boxes, sc, cx, n, i, matched[flat], d, xi, &reg_sum, &bce_sum,
                                       &gd, &gi);
                }
                if (kGrads) {
                    if (lv.g_reg) {
                        float* gr = lv.g_reg + head_offset(n, i, 0, 4, geo.apl, lv.hw);
                        gr[0] = gd.x;
                        gr[lv.hw] = gd.y;
                        gr[2 * (size_t)lv.hw] = gd.z;
                        gr[3 * (size_t)lv.hw] = gd.w;
                    }
                    if (lv.g_iou) lv.g_iou[head_offset(n, i, 0, 1, geo.apl, lv.hw)] = gi;
                }
            }
        }
    }
    // block partial sums (double, fixed order)
    const float cls_sum = fmaf(cx.oma, neg_sum, fix_sum);
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
    FinalPlan plan;
    plan.n_chunks = loss_class_chunks(geo.C);
    int items = 0;
    for (int l = 0; l < geo.num_levels; ++l) {
        const LevelView& lv = geo.lv[l];
        // float4 path: one anchor per location, rows of 4 anchors never straddle a class plane, and every
        // tensor of the level is 16-byte aligned
        auto aligned = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
        const bool vec = geo.apl == 1 && (lv.hw % 4) == 0 && aligned(lv.cls) && aligned(lv.reg) && aligned(lv.iou) &&
                         aligned(lv.g_cls) && aligned(lv.g_reg) && aligned(lv.g_iou);
        plan.vec[l] = vec ? 1 : 0;
        plan.item_off[l] = items;
        const int per_run = vec ? kFinalThreads * 4 : kFinalThreads;
        items += ((lv.n_anchor + per_run - 1) / per_run) * plan.n_chunks;
    }
    for (int l = geo.num_levels; l <= PAA_MAX_LEVELS; ++l) plan.item_off[l] = items;
    plan.items_per_image = items;
    const int grid = geo.num_images * items;
    {
        KernelTimer timer(PAA_KERNEL_FINAL_LOSS, stream);
        const bool g2 = (sc.gamma == 2.0f);
#define PAA_FINAL(G, T)                                                                                  \
    final_loss_kernel<G, T><<<grid, kFinalThreads, 0, stream>>>(geo, go, plan, gt_boxes, sc, ws.paa_label, \
        ws.matched, normalisers, ws.local_norm, grad_losses, ws.block_part)
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
