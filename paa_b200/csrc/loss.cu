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
#include <cstdlib>

#include "kernels.h"

namespace paa {


// partial-sum slots: the bulk kernel's persistent blocks, then one per 128-anchor tile
int loss_grid_blocks(int num_images, int tiles_per_image) {
    const int tiles = num_images * tiles_per_image;
    return 148 * 5 + (tiles > 148 * 8 ? tiles : 148 * 8);       // bulk blocks, then tiles or positive_list blocks
}

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
    if (sc.flavour == PAA_LOSS_RETINANET) {
        // retinanet/loss.py:70,79: reg / max(1, num_pos * BBOX_REG_WEIGHT), cls / (num_pos + N), this rank's counts
        const float npos = (float)norm[0];
        s.cls = g0 / (npos + (float)sc.num_images);
        s.reg = g1 / fmaxf(1.0f, npos * sc.reg_norm_weight);
        s.bce = 0.0f;
        s.weighted = false;
        return s;
    }
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

// The final pass is split by what the work depends on:
//
//   bulk_focal_kernel      every logit of every anchor as a NEGATIVE class: loss term + gradient.  This is
//                          79/80 of the arithmetic and all of the traffic, and it does not care which anchor
//                          or class an element belongs to -- so each level's [N, C, H, W] tensor is streamed
//                          as one flat array, front to back, float4 per lane, 4 float4 in flight per thread
//                          (the access pattern of a memcpy: whole DRAM pages, no strides).
//   positive_terms_kernel  one thread per anchor: the single positive class of every PAA-positive anchor is
//                          patched (loss term and gradient element), its regression / IoU-prediction losses
//                          and gradients are computed, and the regression / IoU gradients of all other
//                          anchors are zeroed.  Reads 4 B and writes 20 B per anchor.
//
// Both leave per-block partial sums in fixed slots; finish_loss_kernel folds them in a fixed order.
constexpr int kBulkThreads = 256;
constexpr int kBulkVecs = 4;                                   // float4 per thread per iteration
constexpr int kBulkChunk = kBulkThreads * kBulkVecs;           // float4 per block iteration (16 KB)
constexpr int kBulkBlocksPerSM = 5;
constexpr int kBulkMaxBlocks = 148 * kBulkBlocksPerSM;

struct BulkPlan {
    const float* src[PAA_MAX_LEVELS];
    float* dst[PAA_MAX_LEVELS];
    unsigned long long count[PAA_MAX_LEVELS];      // floats in the level's tensor
    unsigned chunk_off[PAA_MAX_LEVELS + 1];        // first chunk of each level in the virtual concatenation
    int n;
    // kIgnore only (RetinaNet): one bit per anchor, set = the focal loss ignores the anchor.  Level l's bits are
    // ordered [image][anchor slot][location] -- the order of the level's NCHW planes -- so element e of the
    // level's tensor belongs to bit (e / hw / C) * hw + e % hw.
    const unsigned* ign[PAA_MAX_LEVELS];
    unsigned hw[PAA_MAX_LEVELS];
    unsigned long long magic_hw[PAA_MAX_LEVELS];   // ceil(2^64 / hw) (hw / 4 on the float4 path)
    unsigned long long magic_c;                    // ceil(2^64 / C)
    unsigned C;
    unsigned char vec4[PAA_MAX_LEVELS];            // hw % 4 == 0: the four elements of a float4 share a plane
    int l2_prefetch;                               // percentage of the chunks pulled towards the L2 before the dependency wait
};

// n / d for n < 2^32 with magic = ceil(2^64 / d), d >= 2: exact (error term n / 2^64 < 1 / d).
__device__ __forceinline__ unsigned div_magic(unsigned n, unsigned long long magic) {
    return (unsigned)__umul64hi((unsigned long long)n, magic);
}

// The ignore bit of element e of level l on the per-element path (magic_hw = ceil(2^64 / hw)).
__device__ __forceinline__ unsigned ignore_bit(const BulkPlan& plan, int l, unsigned e) {
    const unsigned hw = plan.hw[l];
    const unsigned plane = div_magic(e, plan.magic_hw[l]);
    const unsigned bit = div_magic(plane, plan.magic_c) * hw + (e - plane * hw);
    return (__ldg(plan.ign[l] + (bit >> 5)) >> (bit & 31u)) & 1u;
}

// The ignore bits of the float4 at float4-index e4 of level l (bit k = element k).
__device__ __forceinline__ unsigned ignore_nibble(const BulkPlan& plan, int l, unsigned long long e4) {
    const unsigned hw = plan.hw[l];
    const unsigned* __restrict__ bits = plan.ign[l];
    if (plan.vec4[l]) {
        const unsigned hw4 = hw >> 2;
        const unsigned plane = div_magic((unsigned)e4, plan.magic_hw[l]);
        const unsigned loc4 = (unsigned)e4 - plane * hw4;
        const unsigned bit = div_magic(plane, plan.magic_c) * hw + (loc4 << 2);
        return (__ldg(bits + (bit >> 5)) >> (bit & 31u)) & 0xfu;
    }
    unsigned nib = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) nib |= ignore_bit(plan, l, (unsigned)(e4 << 2) + k) << k;
    return nib;
}

// negative-class focal term without its (1-alpha) factor, and the gradient with `k` = (1-alpha) * scale
template <bool kG2>
__device__ __forceinline__ void neg_term_grad(float x, float gamma, float k, float* sum, float* grad) {
    const SigmoidLean s = sigmoid_lean(x);
    const float mod = kG2 ? s.p * s.p : __powf(s.p, gamma);
    *sum = fmaf(mod, s.sp, *sum);
    const float gq = kG2 ? s.q + s.q : gamma * s.q;
    *grad = (mod * fmaf(gq, s.sp, s.p)) * k;
}

// element of an ignored anchor: no loss term, zero gradient
__device__ __forceinline__ float ignore_select(unsigned nib, int k, float v) { return (nib >> k) & 1u ? 0.0f : v; }

template <bool kG2>
__device__ __forceinline__ void neg_terms4(float4 x, float gamma, float kneg, unsigned nib, float* neg_sum, float4* g) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    neg_term_grad<kG2>(x.x, gamma, kneg, &t[0], &g->x);
    neg_term_grad<kG2>(x.y, gamma, kneg, &t[1], &g->y);
    neg_term_grad<kG2>(x.z, gamma, kneg, &t[2], &g->z);
    neg_term_grad<kG2>(x.w, gamma, kneg, &t[3], &g->w);
    *neg_sum += ignore_select(nib, 0, t[0]);
    *neg_sum += ignore_select(nib, 1, t[1]);
    *neg_sum += ignore_select(nib, 2, t[2]);
    *neg_sum += ignore_select(nib, 3, t[3]);
    g->x = ignore_select(nib, 0, g->x);
    g->y = ignore_select(nib, 1, g->y);
    g->z = ignore_select(nib, 2, g->z);
    g->w = ignore_select(nib, 3, g->w);
}

template <bool kGrads, bool kG2, bool kIgnore = false>
__global__ void __launch_bounds__(kBulkThreads, kBulkBlocksPerSM)
bulk_focal_kernel(const BulkPlan plan, const LossScalars sc, const double* __restrict__ norm,
                  const double* __restrict__ local_norm, const float* __restrict__ gout,
                  double* __restrict__ block_part) {
    __shared__ double s_part[kBulkThreads / PAA_WARP];
    const float gamma = sc.gamma, oma = 1.0f - sc.alpha;
    float neg_sum = 0.f;
    PAA_TRACE_SCOPE(4);
    const unsigned n_chunks = plan.chunk_off[plan.n];
    unsigned ch = blockIdx.x;
    // Programmatic dependent launch: this grid may become resident while the kernel that produces the normalisers
    // (select_gmm_kernel's slowest fits, the fold kernels) is still running.  The logits do not depend on it, so
    // the block's first chunk is fetched before the dependency wait; everything scaled by 1 / num_pos comes after.
    float4 x0[kBulkVecs];
    bool first_full = false;
    int l0 = 0;
    unsigned long long base0 = 0;
    if (ch < n_chunks) {
#pragma unroll 1
        for (int k = 1; k < plan.n; ++k)
            if (ch >= plan.chunk_off[k]) l0 = k;
        base0 = (unsigned long long)(ch - plan.chunk_off[l0]) * kBulkChunk;
        first_full = base0 + kBulkChunk <= (plan.count[l0] >> 2);
        if (first_full) {
            const float4* __restrict__ src4 = reinterpret_cast<const float4*>(plan.src[l0]);
#pragma unroll
            for (int j = 0; j < kBulkVecs; ++j) x0[j] = __ldcs(src4 + base0 + j * kBulkThreads + threadIdx.x);
        }
    }
    // ... and the block's later chunks are pulled towards the L2 (bulk prefetch, one instruction per 16 KB chunk,
    // issued by one thread: the instruction takes warp-uniform operands) while the memory system idles behind the
    // last EM fits.  Measured: the blocks sit at the wait for ~38 us of select_gmm_kernel's 54; the launch overlap is
    // worth ~2 us of the step, the prefetch another ~1-2 us (profiles/README.md).
    if (plan.l2_prefetch) {
        // no more than ~96 MB: what the 126 MB L2 can hold until the pass gets there (a 1 GB RetinaNet stream
        // prefetched whole just evicts itself and is read twice)
        unsigned pf_end = (unsigned)(((unsigned long long)n_chunks * (unsigned)plan.l2_prefetch) / 100u);
        pf_end = pf_end < 6144u ? pf_end : 6144u;
        for (unsigned c = blockIdx.x + gridDim.x; c < pf_end && threadIdx.x == 0; c += gridDim.x) {
            int l = 0;
#pragma unroll 1
            for (int k = 1; k < plan.n; ++k)
                if (c >= plan.chunk_off[k]) l = k;
            const unsigned long long first = (unsigned long long)(c - plan.chunk_off[l]) * kBulkChunk * 4ull;   // floats
            const unsigned long long left = plan.count[l] - first;
            const unsigned bytes = (unsigned)((left < (unsigned long long)kBulkChunk * 4ull ? left : kBulkChunk * 4ull) * 4ull) & ~15u;
            if (bytes)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(plan.src[l] + first), "r"(bytes) : "memory");
        }
    }
    pdl_wait();
    PAA_TRACE_WAITED();
    pdl_launch_dependents();        // after the wait: at most one future kernel sits resident behind the running one
    const GradScales gs = make_scales(sc, norm, local_norm, gout);
    const float kneg = oma * gs.cls;
    if (first_full) {
        float4* __restrict__ dst4 = reinterpret_cast<float4*>(plan.dst[l0]);
        const bool write = kGrads && dst4 != nullptr;
#pragma unroll
        for (int j = 0; j < kBulkVecs; ++j) {
            float4 g;
            if (kIgnore) {
                neg_terms4<kG2>(x0[j], gamma, kneg, ignore_nibble(plan, l0, base0 + j * kBulkThreads + threadIdx.x),
                                &neg_sum, &g);
            } else {
                neg_term_grad<kG2>(x0[j].x, gamma, kneg, &neg_sum, &g.x);
                neg_term_grad<kG2>(x0[j].y, gamma, kneg, &neg_sum, &g.y);
                neg_term_grad<kG2>(x0[j].z, gamma, kneg, &neg_sum, &g.z);
                neg_term_grad<kG2>(x0[j].w, gamma, kneg, &neg_sum, &g.w);
            }
            if (write) __stcs(dst4 + base0 + j * kBulkThreads + threadIdx.x, g);
        }
        ch += gridDim.x;
    }
    for (; ch < n_chunks; ch += gridDim.x) {
        int l = 0;
#pragma unroll 1
        for (int k = 1; k < plan.n; ++k)
            if (ch >= plan.chunk_off[k]) l = k;
        const unsigned long long count = plan.count[l];
        const unsigned long long n4 = count >> 2;                                  // whole float4s
        const unsigned long long base4 = (unsigned long long)(ch - plan.chunk_off[l]) * kBulkChunk;
        const float4* __restrict__ src4 = reinterpret_cast<const float4*>(plan.src[l]);
        float4* __restrict__ dst4 = reinterpret_cast<float4*>(plan.dst[l]);
        const bool write = kGrads && dst4 != nullptr;
        if (base4 + kBulkChunk <= n4) {
            float4 x[kBulkVecs];
#pragma unroll
            for (int j = 0; j < kBulkVecs; ++j) x[j] = __ldcs(src4 + base4 + j * kBulkThreads + threadIdx.x);
            if (kIgnore) {
                unsigned nib[kBulkVecs];
#pragma unroll
                for (int j = 0; j < kBulkVecs; ++j) nib[j] = ignore_nibble(plan, l, base4 + j * kBulkThreads + threadIdx.x);
#pragma unroll
                for (int j = 0; j < kBulkVecs; ++j) {
                    float4 g;
                    neg_terms4<kG2>(x[j], gamma, kneg, nib[j], &neg_sum, &g);
                    if (write) __stcs(dst4 + base4 + j * kBulkThreads + threadIdx.x, g);
                }
            } else {
#pragma unroll
                for (int j = 0; j < kBulkVecs; ++j) {
                    float4 g;
                    neg_term_grad<kG2>(x[j].x, gamma, kneg, &neg_sum, &g.x);
                    neg_term_grad<kG2>(x[j].y, gamma, kneg, &neg_sum, &g.y);
                    neg_term_grad<kG2>(x[j].z, gamma, kneg, &neg_sum, &g.z);
                    neg_term_grad<kG2>(x[j].w, gamma, kneg, &neg_sum, &g.w);
                    if (write) __stcs(dst4 + base4 + j * kBulkThreads + threadIdx.x, g);
                }
            }
        } else {
            // last chunk of a level: guarded float4s, then the (count % 4) scalar tail
            for (int j = 0; j < kBulkVecs; ++j) {
                const unsigned long long i4 = base4 + j * kBulkThreads + threadIdx.x;
                if (i4 < n4) {
                    const float4 x = __ldcs(src4 + i4);
                    float4 g;
                    if (kIgnore) {
                        neg_terms4<kG2>(x, gamma, kneg, ignore_nibble(plan, l, i4), &neg_sum, &g);
                    } else {
                        neg_term_grad<kG2>(x.x, gamma, kneg, &neg_sum, &g.x);
                        neg_term_grad<kG2>(x.y, gamma, kneg, &neg_sum, &g.y);
                        neg_term_grad<kG2>(x.z, gamma, kneg, &neg_sum, &g.z);
                        neg_term_grad<kG2>(x.w, gamma, kneg, &neg_sum, &g.w);
                    }
                    if (write) __stcs(dst4 + i4, g);
                }
            }
            const unsigned long long tail = (n4 << 2) + threadIdx.x;
            if (threadIdx.x < 4 && tail < count) {
                float g, t = 0.f;
                neg_term_grad<kG2>(plan.src[l][tail], gamma, kneg, &t, &g);
                if (kIgnore) {
                    const unsigned nib = ignore_bit(plan, l, (unsigned)tail);   // count % 4 != 0: per-element path
                    t = ignore_select(nib, 0, t);
                    g = ignore_select(nib, 0, g);
                }
                neg_sum += t;
                if (write) plan.dst[l][tail] = g;
            }
        }
    }
    double a = warp_sum((double)(oma * neg_sum));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_part[warp] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kBulkThreads / PAA_WARP; ++w) t += s_part[w];
        block_part[(size_t)blockIdx.x * 3 + 0] = t;
        block_part[(size_t)blockIdx.x * 3 + 1] = 0.0;
        block_part[(size_t)blockIdx.x * 3 + 2] = 0.0;
    }
}

// Zero regression / IoU-prediction gradients of the non-positive anchors: thread `tid` of `nthr` takes anchors tid,
// tid + nthr, ... four labels requested per trip (under the bulk pass's traffic a dependent round trip takes
// microseconds).  The positives' elements are written by positive_list_kernel's threads: disjoint sets.
__device__ __forceinline__ void zero_one_anchor(const Geometry& geo, unsigned t) {
    const int n = (int)(t / (unsigned)geo.A), a = (int)(t - (unsigned)n * (unsigned)geo.A);
    const int l = anchor_level(geo, a);
    const LevelView& lv = geo.lv[l];
    const int i = a - lv.a_off;
    if (lv.g_reg) {
        float* gr = lv.g_reg + head_offset(geo, lv, n, i, 0, 4);
        store_channels4(gr, head_cstride(geo, lv), make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    }
    if (lv.g_iou) lv.g_iou[head_offset(geo, lv, n, i, 0, 1)] = 0.0f;
}

// labels of anchors tid + u * nthr, u < 4 (1 past the end) -> their zeroes
__device__ __forceinline__ void zero_from_labels(const Geometry& geo, const int (&lab)[4], unsigned tid, unsigned nthr) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (lab[u] <= 0) zero_one_anchor(geo, tid + (unsigned)u * nthr);
}

__device__ __forceinline__ void zero_nonpositive_grads(const Geometry& geo, const int* __restrict__ paa_label,
                                                       unsigned first, unsigned nthr, unsigned total) {
    for (unsigned t0 = first; t0 < total; t0 += 4u * nthr) {
        int lab[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned t = t0 + (unsigned)u * nthr;
            lab[u] = t < total ? __ldcg(paa_label + t) : 1;
        }
        zero_from_labels(geo, lab, t0, nthr);
    }
}

// ---------------------------------------------------------------------------------------------
// bulk_focal_early_kernel: the bulk pass of the PAA step, with part of its work done in select_gmm_kernel's shadow.
//
// The gradients are scaled by 1 / num_pos, which exists only after the slowest EM fit -- for the last ~30 us of
// select_gmm_kernel most SMs hold no fit any more and the memory system idles.  An earlier attempt let every resident
// bulk block work on unscaled gradients during that time and lost what it gained: a fit is one dependent instruction
// chain, and every other ready warp on its SM sub-partition stretches it (profiles/r2_step_trace.txt).  Here a block
// works early only while NO fit lives on its SM (select_gmm_kernel counts the fits per SM in ticket[kCtlSmLive + sm]):
//   before the dependency wait   thread 0 polls {normalisers ready, fits on this SM}; on a free SM the block claims
//                                the next chunk from a global counter, computes the loss terms and stores the
//                                UNSCALED gradients (plain stores: they stay in the L2), up to kEarlyMax chunks;
//   after the wait               it scales its own early chunks in place (the same thread re-reads what it stored;
//                                (h * 1) * k == h * k bit for bit), then claims the remaining chunks one by one.
// Chunks are handed out dynamically, so the order in which a block sums loss terms changes from run to run; the sum
// is therefore taken in fixed point (2^-30 units, exact integer additions: any order gives the same total).  The
// flags are hints only -- what orders memory is the dependency wait.
// ---------------------------------------------------------------------------------------------
constexpr int kEarlyMax = 10;
constexpr unsigned kClaimStop = 0xffffffffu;
constexpr float kFixScale = 1073741824.0f;            // 2^30

struct EarlyCtl {
    unsigned* ctl;          // LossWorkspace::ticket
    unsigned limit;         // chunks [0, limit) may be processed early
    int max_early;          // ... at most this many by one block (<= kEarlyMax): a block scales its early chunks itself
                            //     after the wait, so what a few early-free SMs may take is bounded by twice the fair share
};

__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// one whole chunk: loss terms of this thread's 16 logits (returned), gradients h * k stored
template <bool kG2, bool kKeepInL2>
__device__ __forceinline__ float bulk_chunk(const float4* __restrict__ src4, float4* __restrict__ dst4,
                                            unsigned long long base4, float gamma, float k) {
    float4 x[kBulkVecs];
#pragma unroll
    for (int j = 0; j < kBulkVecs; ++j) x[j] = __ldcs(src4 + base4 + j * kBulkThreads + threadIdx.x);
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < kBulkVecs; ++j) {
        float4 g;
        neg_term_grad<kG2>(x[j].x, gamma, k, &t, &g.x);
        neg_term_grad<kG2>(x[j].y, gamma, k, &t, &g.y);
        neg_term_grad<kG2>(x[j].z, gamma, k, &t, &g.z);
        neg_term_grad<kG2>(x[j].w, gamma, k, &t, &g.w);
        if (kKeepInL2) dst4[base4 + j * kBulkThreads + threadIdx.x] = g;
        else __stcs(dst4 + base4 + j * kBulkThreads + threadIdx.x, g);
    }
    return t;
}

__device__ __forceinline__ int bulk_level_of(const BulkPlan& plan, unsigned ch) {
    int l = 0;
#pragma unroll 1
    for (int k = 1; k < plan.n; ++k)
        if (ch >= plan.chunk_off[k]) l = k;
    return l;
}

template <bool kG2>
__global__ void __launch_bounds__(kBulkThreads, kBulkBlocksPerSM)
bulk_focal_early_kernel(const BulkPlan plan, const LossScalars sc, const double* __restrict__ norm,
                        const double* __restrict__ local_norm, const float* __restrict__ gout,
                        double* __restrict__ block_part, const EarlyCtl ec, const Geometry geo,
                        const int* __restrict__ paa_label) {
    __shared__ long long s_part[kBulkThreads / PAA_WARP];
    __shared__ unsigned s_claim[2];
    __shared__ unsigned s_mine[kEarlyMax];
    const float gamma = sc.gamma, oma = 1.0f - sc.alpha;
    long long acc = 0;
    PAA_TRACE_SCOPE(4);
    const unsigned n_chunks = plan.chunk_off[plan.n];
    unsigned* const counter = ec.ctl + kCtlChunk;
    if (plan.l2_prefetch && threadIdx.x == 0) {
        unsigned pf_end = (unsigned)(((unsigned long long)n_chunks * (unsigned)plan.l2_prefetch) / 100u);
        pf_end = pf_end < 6144u ? pf_end : 6144u;
        for (unsigned c = blockIdx.x; c < pf_end; c += gridDim.x) {
            const int l = bulk_level_of(plan, c);
            const unsigned long long first = (unsigned long long)(c - plan.chunk_off[l]) * kBulkChunk * 4ull;   // floats
            const unsigned long long left = plan.count[l] - first;
            const unsigned bytes = (unsigned)((left < (unsigned long long)kBulkChunk * 4ull ? left : kBulkChunk * 4ull) * 4ull) & ~15u;
            if (bytes)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(plan.src[l] + first), "r"(bytes) : "memory");
        }
    }
    // ---- before the wait: unscaled chunks while no fit lives on this SM ----
    int par = 0, n_mine = 0;
    unsigned carried = kClaimStop;          // a chunk claimed early that has to wait for the normalisers
    if (ec.limit) {
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        const unsigned* live = ec.ctl + kCtlSmLive + (smid & (kCtlMaxSms - 1));
        for (;;) {
            if (threadIdx.x == 0) {
                unsigned c = kClaimStop;
                for (int polls = 0; polls < 4096; ++polls) {        // bounded: a hint must never hang the step
                    if (ld_relaxed(ec.ctl + kCtlReady)) break;
                    if (ld_relaxed(live) == 0u) {
                        c = atomicAdd(counter, 1u);
                        break;
                    }
                    __nanosleep(250);
                }
                s_claim[par] = c;
            }
            __syncthreads();
            const unsigned c = s_claim[par];
            par ^= 1;
            if (c == kClaimStop) break;
            if (c >= ec.limit) {
                carried = c;
                break;
            }
            const int l = bulk_level_of(plan, c);
            const unsigned long long base4 = (unsigned long long)(c - plan.chunk_off[l]) * kBulkChunk;
            if (base4 + kBulkChunk > (plan.count[l] >> 2)) {          // a level's last, partial chunk
                carried = c;
                break;
            }
            const float t = bulk_chunk<kG2, true>(reinterpret_cast<const float4*>(plan.src[l]),
                                                  reinterpret_cast<float4*>(plan.dst[l]), base4, gamma, 1.0f);
            acc += __float2ll_rn(t * kFixScale);
            if (threadIdx.x == 0) s_mine[n_mine] = c;
            if (++n_mine >= ec.max_early) break;
        }
    }
    pdl_wait();
    PAA_TRACE_WAITED();
    pdl_launch_dependents();
    const GradScales gs = make_scales(sc, norm, local_norm, gout);
    const float kneg = oma * gs.cls;
    __syncthreads();                        // s_mine
    // ---- zero regression / IoU gradients of the non-positive anchors (7 MB on C2).  Here and not in
    // positive_list_kernel: that kernel then needs one thread per (GT, slot) only, few enough blocks to be resident
    // beside this pass from its start, and its loads under this pass's traffic take microseconds each ----
    // The first four labels per thread are requested here and used after the block's first pieces of work.
    const unsigned z_tid = blockIdx.x * kBulkThreads + threadIdx.x, z_nthr = gridDim.x * kBulkThreads;
    const unsigned z_total = paa_label ? (unsigned)geo.num_images * (unsigned)geo.A : 0u;
    int z_lab[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const unsigned t = z_tid + (unsigned)u * z_nthr;
        z_lab[u] = t < z_total ? __ldcg(paa_label + t) : 1;
    }
    bool z_pending = true;
    // ---- the early chunks get their scale (L2 hits: this thread re-reads exactly what it stored) ----
    for (int m = 0; m < n_mine; ++m) {
        const unsigned c = s_mine[m];
        const int l = bulk_level_of(plan, c);
        float4* __restrict__ dst4 = reinterpret_cast<float4*>(plan.dst[l]) +
                                    (unsigned long long)(c - plan.chunk_off[l]) * kBulkChunk + threadIdx.x;
        float4 g[kBulkVecs];
#pragma unroll
        for (int j = 0; j < kBulkVecs; ++j) g[j] = __ldcg(dst4 + j * kBulkThreads);
#pragma unroll
        for (int j = 0; j < kBulkVecs; ++j) {
            g[j].x *= kneg;
            g[j].y *= kneg;
            g[j].z *= kneg;
            g[j].w *= kneg;
            __stcs(dst4 + j * kBulkThreads, g[j]);
        }
    }
    // ---- the remaining chunks, claimed one ahead ----
    if (threadIdx.x == 0) s_claim[par] = carried != kClaimStop ? carried : atomicAdd(counter, 1u);
    __syncthreads();
    for (;;) {
        const unsigned c = s_claim[par];
        if (c >= n_chunks) break;
        unsigned next = 0;
        if (threadIdx.x == 0) next = atomicAdd(counter, 1u);
        const int l = bulk_level_of(plan, c);
        const unsigned long long count = plan.count[l];
        const unsigned long long n4 = count >> 2;                                  // whole float4s
        const unsigned long long base4 = (unsigned long long)(c - plan.chunk_off[l]) * kBulkChunk;
        const float4* __restrict__ src4 = reinterpret_cast<const float4*>(plan.src[l]);
        float4* __restrict__ dst4 = reinterpret_cast<float4*>(plan.dst[l]);
        float t = 0.f;
        if (base4 + kBulkChunk <= n4) {
            t = bulk_chunk<kG2, false>(src4, dst4, base4, gamma, kneg);
        } else {
            for (int j = 0; j < kBulkVecs; ++j) {
                const unsigned long long i4 = base4 + j * kBulkThreads + threadIdx.x;
                if (i4 < n4) {
                    const float4 x = __ldcs(src4 + i4);
                    float4 g;
                    neg_term_grad<kG2>(x.x, gamma, kneg, &t, &g.x);
                    neg_term_grad<kG2>(x.y, gamma, kneg, &t, &g.y);
                    neg_term_grad<kG2>(x.z, gamma, kneg, &t, &g.z);
                    neg_term_grad<kG2>(x.w, gamma, kneg, &t, &g.w);
                    __stcs(dst4 + i4, g);
                }
            }
            const unsigned long long tail = (n4 << 2) + threadIdx.x;
            if (threadIdx.x < 4 && tail < count) {
                float g;
                neg_term_grad<kG2>(plan.src[l][tail], gamma, kneg, &t, &g);
                plan.dst[l][tail] = g;
            }
        }
        acc += __float2ll_rn(t * kFixScale);
        if (z_pending) {
            zero_from_labels(geo, z_lab, z_tid, z_nthr);
            z_pending = false;
        }
        if (threadIdx.x == 0) s_claim[par ^ 1] = next;
        __syncthreads();
        par ^= 1;
    }
    if (z_pending) zero_from_labels(geo, z_lab, z_tid, z_nthr);
    if (z_total > 4u * z_nthr) zero_nonpositive_grads(geo, paa_label, z_tid + 4u * z_nthr, z_nthr, z_total);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(PAA_FULL, acc, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
#pragma unroll
        for (int w = 0; w < kBulkThreads / PAA_WARP; ++w) t += s_part[w];
        block_part[(size_t)blockIdx.x * 3 + 0] = (double)oma * ((double)t * (1.0 / (double)kFixScale));
        block_part[(size_t)blockIdx.x * 3 + 1] = 0.0;
        block_part[(size_t)blockIdx.x * 3 + 2] = 0.0;
    }
}

struct FinalCtx {
    float alpha, gamma, oma, kneg;
    GradScales gs;
};

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
        float q;
        if (sc.flavour == PAA_LOSS_ATSS) {
            q = centerness_target(tgt, f);                       // atss/loss.py:233-245
        } else {
            const float4 pred = decode_box(d, f);
            q = iou_plus1(tgt, area_plus1(tgt), pred, area_plus1(pred));
        }
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

// FCOS positive (fcos/loss.py:253-273): IOULoss (layers/iou_loss.py:12-51) between the predicted and the target
// distances weighted by the centerness target, BCE of the centerness logit against the same target.
__device__ __forceinline__ void positive_fcos_terms(const Geometry& geo, const LevelView& lv, const GtOffsets& go,
                                                    const float* __restrict__ gt_boxes, const LossScalars& sc,
                                                    const FinalCtx& cx, int n, int l, int i, int m, float4 p, float xi,
                                                    float* reg_sum, float* bce_sum, float4* gd, float* gi) {
    const float4 pt = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
    const float4 gt = ldg4(gt_boxes + (size_t)(go.v[n] + m) * 4);
    const float4 t = fcos_ltrb(pt.x, pt.y, gt, sc.fcos_norm != 0, sc.fcos_stride[l]);
    const float q = fcos_centerness(t);
    const float ei = expf(-fabsf(xi));
    *bce_sum += fmaxf(xi, 0.0f) - xi * q + log1pf(ei);
    const float sig = (xi >= 0.0f) ? 1.0f / (1.0f + ei) : ei / (1.0f + ei);
    *gi = (sig - q) * cx.gs.bce;
    const float w = cx.gs.weighted ? q : 1.0f;
    // forward (argument order of iou_loss.py: left, top, right, bottom)
    const float t_area = (t.x + t.z) * (t.y + t.w), p_area = (p.x + p.z) * (p.y + p.w);
    const float wi = fminf(p.x, t.x) + fminf(p.z, t.z), hi = fminf(p.w, t.w) + fminf(p.y, t.y);
    const float gw = fmaxf(p.x, t.x) + fmaxf(p.z, t.z), gh = fmaxf(p.w, t.w) + fmaxf(p.y, t.y);
    const float ac = gw * gh + 1e-7f;
    const float inter = wi * hi, uni = t_area + p_area - inter;
    const float iou = (inter + 1.0f) / (uni + 1.0f);
    float loss, g_iou, g_uni = 0.f, g_ac = 0.f;
    if (sc.fcos_iou_type == PAA_IOU_LOSS_IOU) {
        loss = -logf(iou);
        g_iou = -1.0f / iou;
    } else if (sc.fcos_iou_type == PAA_IOU_LOSS_LINEAR) {
        loss = 1.0f - iou;
        g_iou = -1.0f;
    } else {
        loss = 1.0f - (iou - (ac - uni) / ac);          // 1 - iou + 1 - uni / ac
        g_iou = -1.0f;
        g_uni = -1.0f / ac;
        g_ac = uni / (ac * ac);
    }
    *reg_sum += loss * w;
    // backward
    g_uni += g_iou * (-(inter + 1.0f) / ((uni + 1.0f) * (uni + 1.0f)));
    const float g_inter = g_iou / (uni + 1.0f) - g_uni;
    const float g_wi = g_inter * hi, g_hi = g_inter * wi, g_gw = g_ac * gh, g_gh = g_ac * gw;
    const float k = w * cx.gs.reg;
    gd->x = (g_uni * (p.y + p.w) + g_wi * pick_first(p.x, t.x, false) + g_gw * pick_first(p.x, t.x, true)) * k;
    gd->z = (g_uni * (p.y + p.w) + g_wi * pick_first(p.z, t.z, false) + g_gw * pick_first(p.z, t.z, true)) * k;
    gd->y = (g_uni * (p.x + p.z) + g_hi * pick_first(p.y, t.y, false) + g_gh * pick_first(p.y, t.y, true)) * k;
    gd->w = (g_uni * (p.x + p.z) + g_hi * pick_first(p.w, t.w, false) + g_gh * pick_first(p.w, t.w, true)) * k;
}

// Smooth-L1 regression loss / gradient of one RetinaNet positive (layers/smooth_l1_loss.py:6-17 on
// box_coder.py:22-50 targets).
__device__ __forceinline__ float smooth_l1_term(float x, float t, float beta, float k, float* g) {
    const float d = x - t, n = fabsf(d);
    if (n < beta) {
        *g = (d / beta) * k;
        return 0.5f * n * n / beta;
    }
    *g = (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)) * k;
    return n - 0.5f * beta;
}

__device__ __forceinline__ void positive_smooth_l1(const Geometry& geo, const LevelView& lv, const GtOffsets& go,
                                                   const float* __restrict__ gt_boxes, const LossScalars& sc,
                                                   const FinalCtx& cx, int n, int i, int m, float4 d,
                                                   float* reg_sum, float4* gd) {
    const float4 a = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
    const float4 gt = ldg4(gt_boxes + (size_t)(go.v[n] + m) * 4);
    const float4 t = encode_box_legacy(gt, a, sc.code_w[0], sc.code_w[1], sc.code_w[2], sc.code_w[3]);
    float s = 0.f;
    s += smooth_l1_term(d.x, t.x, sc.beta, cx.gs.reg, &gd->x);
    s += smooth_l1_term(d.y, t.y, sc.beta, cx.gs.reg, &gd->y);
    s += smooth_l1_term(d.z, t.z, sc.beta, cx.gs.reg, &gd->z);
    s += smooth_l1_term(d.w, t.w, sc.beta, cx.gs.reg, &gd->w);
    *reg_sum += s;
}

// One thread per anchor, 128-anchor tiles (same tiling as the assignment kernels); a block takes
// `tiles_per_block` consecutive tiles (1 unless the call has more than kMaxTileBlocks tiles, so that the fold of
// the partials stays short).  Runs after bulk_focal_kernel on the same stream: it overwrites the labelled
// class's gradient element.  `zero_fill`: this kernel also zeroes the regression / IoU gradients of the
// non-positive anchors (coalesced for one anchor per location; with several the launcher clears them up front).
constexpr int kMaxTileBlocks = 8192;

template <bool kGrads, bool kG2>
__global__ void __launch_bounds__(PAA_TILE)
positive_terms_kernel(const Geometry geo, const GtOffsets* __restrict__ gop, const float* __restrict__ gt_boxes,
                      const LossScalars sc, const int* __restrict__ paa_label, const int* __restrict__ matched,
                      const double* __restrict__ norm, const double* __restrict__ local_norm,
                      const float* __restrict__ gout, double* __restrict__ block_part, int tiles_total,
                      int tiles_per_block, bool zero_fill, bool patch_ignored) {
    __shared__ double s_part[PAA_TILE / PAA_WARP][3];
    __shared__ int s_ign[PAA_TILE];                 // ignored anchors of the tile (index within the level)
    pdl_wait();
    pdl_launch_dependents();
    const GtOffsets& go = *gop;
    __shared__ int s_wcnt[PAA_TILE / PAA_WARP];
    float fix_sum = 0.f, reg_sum = 0.f, bce_sum = 0.f, ign_sum = 0.f;
    const bool may_ignore = patch_ignored;                          // block-uniform
    const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5;
    const int t_end = min((int)(blockIdx.x + 1) * tiles_per_block, tiles_total);
    for (int t = blockIdx.x * tiles_per_block; t < t_end; ++t) {
        const int n = t / geo.tiles_per_image;
        const int tile = t - n * geo.tiles_per_image;
        int first;
        const int l = tile_level(geo, tile, &first);
        const LevelView& lv = geo.lv[l];
        const int i = first + threadIdx.x;
        const bool valid = i < lv.n_anchor;
        const size_t flat = (size_t)n * geo.A + lv.a_off + (valid ? i : 0);
        const int label = valid ? __ldg(paa_label + flat) : 0;
        float4 gd = make_float4(0.f, 0.f, 0.f, 0.f);
        float gi = 0.f;
        if (label > 0) {
            FinalCtx cx;
            cx.gs = make_scales(sc, norm, local_norm, gout);
            cx.alpha = sc.alpha;
            cx.gamma = sc.gamma;
            cx.oma = 1.0f - sc.alpha;
            cx.kneg = cx.oma * cx.gs.cls;
            // classification: swap the negative-class result of the labelled class for the positive one
            const size_t off = head_offset(geo, lv, n, i, label - 1, geo.C);
            const float xp = __ldg(lv.cls + off);
            const SigmoidParts sp = sigmoid_parts(xp);
            float tn_acc = 0.f, g_unused;
            neg_term_grad<kG2>(xp, cx.gamma, cx.kneg, &tn_acc, &g_unused);      // exactly what the bulk pass added
            float tn, gn, tp, gp;
            focal_negative(xp, sp, cx.gamma, kG2, cx.oma, &tn, &gn);
            focal_positive(xp, sp, cx.gamma, kG2, cx.alpha, &tp, &gp);
            (void)tn;
            (void)gn;
            fix_sum += tp - cx.oma * tn_acc;
            if (kGrads && lv.g_cls) lv.g_cls[off] = gp * cx.gs.cls;
            // box regression + IoU prediction
            const float* rp = lv.reg + head_offset(geo, lv, n, i, 0, 4);
            const float4 d = load_channels4(rp, head_cstride(geo, lv));
            const float xi = sc.use_iou_pred ? __ldg(lv.iou + head_offset(geo, lv, n, i, 0, 1)) : 0.f;
            if (sc.flavour == PAA_LOSS_RETINANET)
                positive_smooth_l1(geo, lv, go, gt_boxes, sc, cx, n, i, matched[flat], d, &reg_sum, &gd);
            else if (sc.flavour == PAA_LOSS_FCOS)
                positive_fcos_terms(geo, lv, go, gt_boxes, sc, cx, n, l, i, matched[flat], d, xi, &reg_sum, &bce_sum,
                                    &gd, &gi);
            else
                positive_box_terms(geo, lv, go, gt_boxes, sc, cx, n, i, matched[flat], d, xi, &reg_sum, &bce_sum,
                                   &gd, &gi);
        }
        if (kGrads && valid && (zero_fill || label > 0)) {
            if (lv.g_reg) {
                float* gr = lv.g_reg + head_offset(geo, lv, n, i, 0, 4);
                store_channels4(gr, head_cstride(geo, lv), make_float4(gd.x, gd.y, gd.z, gd.w));
            }
            if (lv.g_iou) lv.g_iou[head_offset(geo, lv, n, i, 0, 1)] = gi;
        }
        if (may_ignore) {
            // Ignored anchors (RetinaNet, between the Matcher thresholds) when the bulk pass could not skip them
            // itself: the focal loss ignores all of their classes (sigmoid_focal_loss.py:50,
            // SigmoidFocalLoss_cuda.cu:44) -- take back what the bulk pass added.  Their C elements are hw floats apart, so the whole block shares them: the tile's ignored
            // anchors are listed in anchor order (deterministic) and the (class, anchor) pairs dealt out to all
            // threads as independent loads, instead of one thread walking C dependent cache misses.
            const unsigned bal = __ballot_sync(PAA_FULL, label < 0);
            if (lane_ == 0) s_wcnt[warp_] = __popc(bal);
            __syncthreads();
            int base = 0, total = 0;
#pragma unroll
            for (int w = 0; w < PAA_TILE / PAA_WARP; ++w) {
                if (w < warp_) base += s_wcnt[w];
                total += s_wcnt[w];
            }
            if (label < 0) s_ign[base + __popc(bal & ((1u << lane_) - 1u))] = i;
            __syncthreads();
            const int items = total * geo.C;
            for (int it = threadIdx.x; it < items; it += PAA_TILE) {
                const int c = it / total, k = it - c * total;
                const size_t off = head_offset(geo, lv, n, s_ign[k], c, geo.C);
                float g_unused;
                neg_term_grad<kG2>(__ldg(lv.cls + off), sc.gamma, 0.0f, &ign_sum, &g_unused);
                if (kGrads && lv.g_cls) lv.g_cls[off] = 0.0f;
            }
        }
    }
    fix_sum -= (1.0f - sc.alpha) * ign_sum;
    double a0 = warp_sum((double)fix_sum), a1 = warp_sum((double)reg_sum), a2 = warp_sum((double)bce_sum);
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
        for (int w = 0; w < PAA_TILE / PAA_WARP; ++w) t += s_part[w][threadIdx.x];
        block_part[(size_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
}

// Folds the per-block partials in a fixed order and applies the normalisers (loss.py:354-358).
// One block of 512 threads; every thread issues its (up to four) independent loads before adding, so the
// fold costs a couple of memory round trips instead of one per 256 partials.
constexpr int kFinishThreads = 512;

__device__ __forceinline__ void fold_partials(const double* __restrict__ part, int blocks, double (&a)[3]) {
    for (int b0 = threadIdx.x; b0 < blocks; b0 += 4 * kFinishThreads) {
        double v[4][3];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int b = b0 + u * kFinishThreads;
#pragma unroll
            for (int k = 0; k < 3; ++k) v[u][k] = (b < blocks) ? __ldcg(part + (size_t)b * 3 + k) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 3; ++k) a[k] += v[u][k];
    }
}

// loss.py:354-358 (and the other flavours' normalisations) applied to the folded sums {cls, reg, iou-pred / centerness}
__device__ __forceinline__ void write_losses(const double (&t)[3], const LossScalars& sc, const double* __restrict__ norm,
                                             float* __restrict__ losses) {
    const double world = (double)sc.world_size;
    const float num_pos_avg = (float)fmax(norm[0] / world, 1.0);
    if (sc.flavour == PAA_LOSS_RETINANET) {
        const float npos = (float)norm[0];
        losses[0] = (float)t[0] / (npos + (float)sc.num_images);
        losses[1] = (float)t[1] / fmaxf(1.0f, npos * sc.reg_norm_weight);
        losses[2] = 0.0f;
        return;
    }
    losses[0] = (float)t[0] / num_pos_avg;
    if (sc.use_iou_pred) {
        const float reg_norm = (float)(norm[1] / world);
        losses[1] = (float)t[1] / reg_norm * sc.reg_loss_weight;
        // a rank without positives returns an empty sum instead (atss/loss.py:274-277, fcos/loss.py:274-277)
        if (sc.flavour != PAA_LOSS_PAA && t[1] == 0.0) losses[1] = 0.0f;
        losses[2] = (float)t[2] / num_pos_avg * sc.iou_loss_weight;
    } else {
        losses[1] = (float)t[1] / num_pos_avg * sc.reg_loss_weight;
        losses[2] = 0.0f;
    }
}

__global__ void __launch_bounds__(kFinishThreads)
finish_loss_kernel(const double* __restrict__ part_a, int blocks_a, const double* __restrict__ part_b,
                   int blocks_b, const LossScalars sc, const double* __restrict__ norm,
                   float* __restrict__ losses) {
    __shared__ double s[kFinishThreads / PAA_WARP][3];
    pdl_wait();
    double a[3] = {0.0, 0.0, 0.0};
    fold_partials(part_a, blocks_a, a);
    fold_partials(part_b, blocks_b, a);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a[k] = warp_sum(a[k]);
        if (lane == 0) s[warp][k] = a[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[3] = {0.0, 0.0, 0.0};
        for (int w = 0; w < kFinishThreads / PAA_WARP; ++w)
            for (int k = 0; k < 3; ++k) t[k] += s[w][k];
        write_losses(t, sc, norm, losses);
    }
}

// ---------------------------------------------------------------------------------------------
// PAA's positives, list-driven.  select_gmm_kernel leaves every GT's positive prefix as a short list of
// anchors; one thread per (GT, slot) patches the labelled class (loss term and gradient element), computes the
// GIoU / IoU-prediction losses and gradients -- the ~8 positives per GT of a batch instead of a scan of all
// N x A anchors for the 2 % that are positive (the scan kernel, still used by the other flavours, spent its time
// at the block barrier behind the few warps that held a positive: 17 us for the C2 batch).  The same launch
// zeroes the regression / IoU-prediction gradients of the non-positive anchors (one thread per anchor), and
// its last block folds all partial sums and writes the three losses (no separate finish launch).
// ---------------------------------------------------------------------------------------------
constexpr int kPosThreads = 128;
constexpr int kPosMaxBlocks = 148 * 8;

template <bool kGrads, bool kG2>
__global__ void __launch_bounds__(kPosThreads, 8)
positive_list_kernel(const Geometry geo, const GtOffsets* __restrict__ gop, const int* __restrict__ gt_image,
                     const float* __restrict__ gt_boxes, const int64_t* __restrict__ gt_labels, const LossScalars sc,
                     const int* __restrict__ pos_list, const int* __restrict__ part_npos, int cap,
                     const int* __restrict__ paa_label, bool zero_here, const double* __restrict__ norm,
                     const double* __restrict__ local_norm, const float* __restrict__ gout,
                     const double* __restrict__ bulk_part, int bulk_blocks, double* __restrict__ block_part,
                     unsigned* __restrict__ ticket, float* __restrict__ losses) {
    __shared__ double s_part[kPosThreads / PAA_WARP][3];
    __shared__ bool s_last;
    PAA_TRACE_SCOPE(5);
    // Only two things here depend on bulk_focal_kernel, the programmatic predecessor: the labelled class's gradient
    // element (bulk_focal wrote the negative-class value there) and the fold of its partial sums.  Everything else
    // reads what select_gmm_kernel and its predecessors left -- complete and visible before this grid's first block
    // starts, because every bulk_focal block calls launch_dependents AFTER its own dependency wait -- and is done
    // before the wait, in the shadow of the bulk pass: the zero fill, the positives' losses and regression / IoU
    // gradients, the block's partial sums and its ticket.  (Reads of that data go to the L2: ld.cg.)  After the
    // wait: one store per positive and the last block's fold.  13 -> ~5 us behind bulk_focal's end on C2.
    const GtOffsets& go = *gop;
    const int num_gt = go.v[geo.num_images];
    const unsigned tid = blockIdx.x * kPosThreads + threadIdx.x, nthr = gridDim.x * kPosThreads;
    // (a) zero gradients of the non-positive anchors (unless the bulk pass did it): the positives' are written below
    if (kGrads && zero_here) zero_nonpositive_grads(geo, paa_label, tid, nthr, (unsigned)geo.num_images * (unsigned)geo.A);
    // (b) the positives
    float fix_sum = 0.f, reg_sum = 0.f, bce_sum = 0.f;
    const unsigned items = (unsigned)num_gt * (unsigned)cap;
    const double nrm[2] = {__ldcg(norm), __ldcg(norm + 1)}, lnrm[2] = {0.0, __ldcg(local_norm + 1)};
    FinalCtx cx;
    cx.gs = make_scales(sc, nrm, lnrm, gout);
    cx.alpha = sc.alpha;
    cx.gamma = sc.gamma;
    cx.oma = 1.0f - sc.alpha;
    cx.kneg = cx.oma * cx.gs.cls;
    float* patch_dst[2] = {nullptr, nullptr};     // the thread's first two positives: their class gradient elements,
    float patch_val[2] = {0.f, 0.f};              // stored after the wait
    for (unsigned t = tid; t < items; t += nthr) {
        const int gi = (int)(t / (unsigned)cap), j = (int)(t - (unsigned)gi * (unsigned)cap);
        // one memory round trip for everything that only depends on (GT, slot): the prefix length, the GT's image and
        // class (= the label select_gmm_kernel wrote for its positives), the slot's anchor (stale past the prefix)
        const int npos = __ldcg(part_npos + gi);
        const int n = __ldcg(gt_image + gi);
        const int a = __ldcg(pos_list + (size_t)gi * cap + j);
        const int label = (int)__ldg(gt_labels + gi);
        if (j >= npos) continue;
        const int l = anchor_level(geo, a);
        const LevelView& lv = geo.lv[l];
        const int i = a - lv.a_off;
        // classification: swap the negative-class result of the labelled class for the positive one
        const size_t off = head_offset(geo, lv, n, i, label - 1, geo.C);
        const float xp = __ldg(lv.cls + off);
        const SigmoidParts sp = sigmoid_parts(xp);
        float tn_acc = 0.f, g_unused;
        neg_term_grad<kG2>(xp, cx.gamma, cx.kneg, &tn_acc, &g_unused);      // exactly what the bulk pass added
        float tp, gp;
        focal_positive(xp, sp, cx.gamma, kG2, cx.alpha, &tp, &gp);
        fix_sum += tp - cx.oma * tn_acc;
        if (kGrads && lv.g_cls) {
            if (t == tid) {
                patch_dst[0] = lv.g_cls + off;
                patch_val[0] = gp * cx.gs.cls;
            } else if (t == tid + nthr) {
                patch_dst[1] = lv.g_cls + off;
                patch_val[1] = gp * cx.gs.cls;
            }
        }
        // box regression + IoU prediction
        const float* rp = lv.reg + head_offset(geo, lv, n, i, 0, 4);
        const float4 d = load_channels4(rp, head_cstride(geo, lv));
        const float xi = sc.use_iou_pred ? __ldg(lv.iou + head_offset(geo, lv, n, i, 0, 1)) : 0.f;
        float4 gd = make_float4(0.f, 0.f, 0.f, 0.f);
        float gi_ = 0.f;
        positive_box_terms(geo, lv, go, gt_boxes, sc, cx, n, i, gi - go.v[n], d, xi, &reg_sum, &bce_sum, &gd, &gi_);
        if (kGrads) {
            if (lv.g_reg) {
                float* gr = lv.g_reg + head_offset(geo, lv, n, i, 0, 4);
                store_channels4(gr, head_cstride(geo, lv), make_float4(gd.x, gd.y, gd.z, gd.w));
            }
            if (lv.g_iou) lv.g_iou[head_offset(geo, lv, n, i, 0, 1)] = gi_;
        }
    }
    double a0 = warp_sum((double)fix_sum), a1 = warp_sum((double)reg_sum), a2 = warp_sum((double)bce_sum);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_part[warp][0] = a0;
        s_part[warp][1] = a1;
        s_part[warp][2] = a2;
    }
    __syncthreads();
    // thread 0 publishes the block's partial sums and takes a ticket; the last block to do so folds every partial
    // sum in a fixed order and writes the losses
    if (threadIdx.x == 0) {
        double t[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int w = 0; w < kPosThreads / PAA_WARP; ++w)
#pragma unroll
            for (int k = 0; k < 3; ++k) t[k] += s_part[w][k];
#pragma unroll
        for (int k = 0; k < 3; ++k) block_part[(size_t)blockIdx.x * 3 + k] = t[k];
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    // the block that took the last ticket folds this grid's partial sums, still in front of the wait ...
    __shared__ double s_own[kPosThreads / PAA_WARP][3];
    if (s_last) {
        double acc[3] = {0.0, 0.0, 0.0};
        __threadfence();
        for (int b0 = threadIdx.x; b0 < (int)gridDim.x; b0 += 4 * kPosThreads) {
            double v[4][3];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int b = b0 + u * kPosThreads;
#pragma unroll
                for (int k = 0; k < 3; ++k) v[u][k] = b < (int)gridDim.x ? __ldcg(block_part + (size_t)b * 3 + k) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < 3; ++k) acc[k] += v[u][k];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            acc[k] = warp_sum(acc[k]);
            if (lane == 0) s_own[warp][k] = acc[k];
        }
    }
    PAA_TRACE_POINT(7);                 // when the blocks are done with the work in front of the wait
    pdl_wait();
    PAA_TRACE_WAITED();
    pdl_launch_dependents();
    // (c) behind the bulk pass: the labelled class's gradient element of every positive
    if (kGrads) {
        if (patch_dst[0]) *patch_dst[0] = patch_val[0];
        if (patch_dst[1]) *patch_dst[1] = patch_val[1];
        for (unsigned t = tid + 2u * nthr; t < items; t += nthr) {      // calls with more (GT, slot) items than that
            const int gi = (int)(t / (unsigned)cap), j = (int)(t - (unsigned)gi * (unsigned)cap);
            const int npos = __ldcg(part_npos + gi);
            const int n = __ldcg(gt_image + gi);
            const int a = __ldcg(pos_list + (size_t)gi * cap + j);
            const int label = (int)__ldg(gt_labels + gi);
            if (j >= npos) continue;
            const LevelView& lv = geo.lv[anchor_level(geo, a)];
            if (!lv.g_cls) continue;
            const size_t off = head_offset(geo, lv, n, a - lv.a_off, label - 1, geo.C);
            const float xp = __ldg(lv.cls + off);
            float tp, gp;
            focal_positive(xp, sigmoid_parts(xp), cx.gamma, kG2, cx.alpha, &tp, &gp);
            lv.g_cls[off] = gp * cx.gs.cls;
        }
    }
    if (!s_last) return;
    // ... and the bulk pass's behind it (its blocks leave their sum in slot 0; eight loads in flight per thread)
    double bulk = 0.0;
    for (int b0 = threadIdx.x; b0 < bulk_blocks; b0 += 8 * kPosThreads) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int b = b0 + u * kPosThreads;
            v[u] = b < bulk_blocks ? __ldcg(bulk_part + (size_t)b * 3) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) bulk += v[u];
    }
    bulk = warp_sum(bulk);
    if (lane == 0) s_part[warp][0] = bulk;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[3] = {0.0, 0.0, 0.0};
        for (int w = 0; w < kPosThreads / PAA_WARP; ++w) {
            t[0] += s_part[w][0];
            for (int k = 0; k < 3; ++k) t[k] += s_own[w][k];
        }
        write_losses(t, sc, norm, losses);
        *ticket = 0u;                   // paa_loss may be called again on the same assignment
        ticket[kCtlChunk - 1] = 0u;     // ... and so may the bulk pass's chunk counter (ticket = LossWorkspace::ticket + 1)
    }
}

#ifdef PAA_TRACE
PAA_TRACE_SETTER(trace_set_loss)
#endif

int launch_final_loss(const Geometry& geo, const float* gt_boxes, const int64_t* gt_labels,
                      const LossScalars& sc, const LossWorkspace& ws, const double* normalisers,
                      const float* grad_losses, float* losses, bool write_grads, cudaStream_t stream) {
    const GtOffsets* gop = ws.go;
    BulkPlan plan;
    plan.n = geo.num_levels;
    unsigned chunks = 0;
    for (int l = 0; l < geo.num_levels; ++l) {
        const LevelView& lv = geo.lv[l];
        plan.src[l] = lv.cls;
        plan.dst[l] = lv.g_cls;
        plan.count[l] = (unsigned long long)geo.num_images * geo.apl * geo.C * lv.hw;
        plan.chunk_off[l] = chunks;
        const unsigned long long n4 = (plan.count[l] + 3) / 4;
        chunks += (unsigned)((n4 + kBulkChunk - 1) / kBulkChunk);
        if ((reinterpret_cast<uintptr_t>(lv.cls) & 15u) || (reinterpret_cast<uintptr_t>(lv.g_cls) & 15u)) {
            set_error("level %d: box_cls / grad_box_cls must be 16-byte aligned", l);
            return PAA_ERR_BAD_ARGUMENT;
        }
    }
    for (int l = geo.num_levels; l <= PAA_MAX_LEVELS; ++l) plan.chunk_off[l] = chunks;
    for (int l = geo.num_levels; l < PAA_MAX_LEVELS; ++l) {
        plan.src[l] = nullptr;
        plan.dst[l] = nullptr;
        plan.count[l] = 0;
    }
    // RetinaNet / ATSS 'IoU': the labelling kernel left one bit per ignored anchor (retina.cu); the bulk pass skips those
    // elements in stream, which saves positive_terms_kernel C scattered reads and writes per ignored anchor.
    const bool has_ignored = sc.flavour == PAA_LOSS_RETINANET ||
                             (sc.flavour == PAA_LOSS_ATSS && sc.atss_type == PAA_ATSS_POSITIVE_IOU);
    bool bulk_ignores = has_ignored && geo.C >= 2;
    if (getenv("PAA_RETINA_PATCH")) bulk_ignores = false;      // test hook: force the per-anchor fallback
    if (geo.nhwc) bulk_ignores = false;      // the ignore bits are in NCHW plane order: channels-last calls patch per anchor
    for (int l = 0; l < geo.num_levels && bulk_ignores; ++l)
        if (plan.count[l] >= (1ull << 32) || geo.lv[l].hw < 2 || geo.lv[l].hw == 4) bulk_ignores = false;
    plan.C = (unsigned)geo.C;
    plan.magic_c = geo.C >= 2 ? ~0ull / (unsigned)geo.C + 1ull : 0ull;
    for (int l = 0; l < PAA_MAX_LEVELS; ++l) {
        plan.ign[l] = nullptr;
        plan.hw[l] = 1;
        plan.magic_hw[l] = 0ull;
        plan.vec4[l] = 0;
        if (l >= geo.num_levels || !bulk_ignores) continue;
        const LevelView& lv = geo.lv[l];
        plan.ign[l] = reinterpret_cast<const unsigned*>(ws.tile_gtmask) + (size_t)geo.num_images * lv.tile_off * 4;
        plan.hw[l] = (unsigned)lv.hw;
        plan.vec4[l] = (lv.hw & 3) == 0 ? 1 : 0;
        const unsigned d = plan.vec4[l] ? (unsigned)lv.hw >> 2 : (unsigned)lv.hw;
        plan.magic_hw[l] = ~0ull / d + 1ull;       // ceil(2^64 / d) for d >= 2 that is not a power of two, and
                                                   // 2^64 / d exactly when it is
    }
    plan.l2_prefetch = getenv("PAA_L2_PREFETCH_PCT") ? atoi(getenv("PAA_L2_PREFETCH_PCT")) : 100;
    int bulk_grid = kBulkMaxBlocks;
    if ((unsigned)bulk_grid > chunks) bulk_grid = (int)chunks;
    const int tiles_total = geo.num_images * geo.tiles_per_image;
    const int tiles_per_block = (tiles_total + kMaxTileBlocks - 1) / kMaxTileBlocks;
    const int tile_grid = (tiles_total + tiles_per_block - 1) / tiles_per_block;
    // several anchors per location: an anchor's regression channels are hw floats apart from its neighbour's, so
    // zeroing them anchor by anchor scatters 4-byte writes; clear the tensors up front and write positives only
    const bool zero_fill = geo.apl == 1 || geo.nhwc;      // (channels-last: an anchor's channels are contiguous)
    if (write_grads && !zero_fill) {
        for (int l = 0; l < geo.num_levels; ++l) {
            const LevelView& lv = geo.lv[l];
            const size_t per = (size_t)geo.num_images * geo.apl * lv.hw * sizeof(float);
            if (lv.g_reg) PAA_CUDA_CHECK(cudaMemsetAsync(lv.g_reg, 0, per * 4, stream));
            if (lv.g_iou) PAA_CUDA_CHECK(cudaMemsetAsync(lv.g_iou, 0, per, stream));
        }
    }
    const bool g2 = (sc.gamma == 2.0f);
    bool bulk_zeroes = false;        // bulk_focal_early_kernel zeroed the non-positive anchors' regression / IoU gradients
    double* bulk_part = ws.block_part;
    double* tile_part = ws.block_part + (size_t)kBulkMaxBlocks * 3;
    {
        KernelTimer timer(PAA_KERNEL_FINAL_LOSS, stream);
        // events around the launch serialise it behind its predecessor: no overlap window, so no point in pulling
        // chunks towards the L2 ahead of the stream (it only adds work to the kernel timed alone)
        if (timer.slot_ >= 0) plan.l2_prefetch = 0;
        // launched as a programmatic dependent of the previous kernel on the stream (see the kernel's prologue)
#define PAA_BULK(G, T, I)                                                                                    \
    PAA_PDL_LAUNCH((bulk_focal_kernel<G, T, I>), bulk_grid, kBulkThreads, stream, plan, sc, normalisers,     \
                   ws.local_norm, grad_losses, bulk_part)
        // PAA with one fused select_gmm launch in front: bulk_focal_early_kernel (dynamic chunks, fixed-point loss
        // sum), part of the pass in the shadow of the last EM fits.  PAA_BULK_EARLY_PCT = share of the chunks that may
        // be taken early (0 = none; negative = the static kernel, for comparisons).  The two-launch form keeps no
        // per-SM fit count: static kernel.  (Several ranks: norm_wait_kernel in between is a programmatic dependent
        // launch that lets this one start at once; behind an NCCL all-reduce there is no shadow and no early chunk.)
        // Four blocks per SM instead of five: the registers left over let positive_list_kernel's first blocks
        // become resident (and do their work that does not depend on this pass) while the pass runs.
        bulk_zeroes = false;
        bool dynamic = sc.flavour == PAA_LOSS_PAA && geo.apl == 1 && write_grads && !bulk_ignores &&
                       !select_gmm_two_launch(sc);
        for (int l = 0; l < geo.num_levels; ++l)
            if (!plan.dst[l]) dynamic = false;
        int early_pct = 60;
        if (const char* e = getenv("PAA_BULK_EARLY_PCT")) early_pct = atoi(e);
        if (early_pct < 0) dynamic = false;
        // small calls: the claims and the in-place scaling cost what they hide (2 images of C2's size, 875 chunks: 72.6
        // against 72.4 us; from 4 images, 1750 chunks, on the dynamic form wins: 81.9 against 85.1 us)
        unsigned min_chunks = 1200u;
        if (const char* e = getenv("PAA_BULK_EARLY_MIN_CHUNKS")) min_chunks = (unsigned)atoi(e);      // test / measurement switch
        if (chunks < min_chunks) dynamic = false;
        EarlyCtl ec;
        ec.ctl = ws.ticket;
        ec.limit = (unsigned)((unsigned long long)chunks * (unsigned)(early_pct < 0 ? 0 : (early_pct > 100 ? 100 : early_pct)) / 100u);
        ec.max_early = kEarlyMax;
        if (dynamic) {
            int bps = 4;
            if (const char* e = getenv("PAA_BULK_BPS")) bps = atoi(e);          // measurement switch
            if (bps >= 1 && bps <= kBulkBlocksPerSM && (unsigned)(148 * bps) < chunks) bulk_grid = 148 * bps;
            const int fair = (int)((ec.limit + (unsigned)bulk_grid - 1u) / (unsigned)bulk_grid);
            ec.max_early = 2 * fair < kEarlyMax ? (2 * fair < 1 ? 1 : 2 * fair) : kEarlyMax;
            if (g2) PAA_PDL_LAUNCH((bulk_focal_early_kernel<true>), bulk_grid, kBulkThreads, stream, plan, sc, normalisers,
                                   ws.local_norm, grad_losses, bulk_part, ec, geo, (const int*)ws.paa_label);
            else PAA_PDL_LAUNCH((bulk_focal_early_kernel<false>), bulk_grid, kBulkThreads, stream, plan, sc, normalisers,
                                ws.local_norm, grad_losses, bulk_part, ec, geo, (const int*)ws.paa_label);
            bulk_zeroes = true;
        } else if (bulk_ignores) {
            if (write_grads) {
                if (g2) PAA_BULK(true, true, true); else PAA_BULK(true, false, true);
            } else {
                if (g2) PAA_BULK(false, true, true); else PAA_BULK(false, false, true);
            }
        } else if (write_grads) {
            if (g2) PAA_BULK(true, true, false); else PAA_BULK(true, false, false);
        } else {
            if (g2) PAA_BULK(false, true, false); else PAA_BULK(false, false, false);
        }
#undef PAA_BULK
    }
    if (sc.flavour == PAA_LOSS_PAA && geo.apl == 1) {
        // PAA: the positives come as per-GT lists from select_gmm_kernel; the same launch zeroes the other anchors'
        // regression / IoU gradients and its last block writes the losses
        const int cap = geo.num_levels * sc.topk;
        const long long work_a = bulk_zeroes ? 0 : (long long)geo.num_images * geo.A, work_b = (long long)sc.gt_capacity * cap;
        long long blocks = ((work_a > work_b ? work_a : work_b) + kPosThreads - 1) / kPosThreads;
        int pos_grid = (int)(blocks < 1 ? 1 : (blocks > kPosMaxBlocks ? kPosMaxBlocks : blocks));
        // beside bulk_focal_early_kernel's four blocks per SM two of this kernel's are resident: no more than that,
        // so that all of them work in the bulk pass's shadow (a thread keeps two positives' class elements for later)
        if (bulk_zeroes && pos_grid > 148 * 2) pos_grid = 148 * 2;
        if (const char* e = getenv("PAA_POS_BLOCKS")) {             // measurement switch
            const int pb = atoi(e);
            if (pb >= 1 && pb < pos_grid) pos_grid = pb;
        }
        KernelTimer timer(PAA_KERNEL_POSITIVE_TERMS, stream);
#define PAA_POSL(G, T)                                                                                       \
    PAA_PDL_LAUNCH((positive_list_kernel<G, T>), pos_grid, kPosThreads, stream, geo, gop, ws.gt_image, gt_boxes, gt_labels, sc, \
        ws.pos_list, ws.part_npos, cap, ws.paa_label, !bulk_zeroes, normalisers, ws.local_norm, grad_losses, bulk_part, bulk_grid, \
        tile_part, ws.ticket + 1, losses)
        if (write_grads) {
            if (g2) PAA_POSL(true, true); else PAA_POSL(true, false);
        } else {
            if (g2) PAA_POSL(false, true); else PAA_POSL(false, false);
        }
#undef PAA_POSL
        return 0;
    }
#define PAA_POS(G, T)                                                                                        \
    PAA_PDL_LAUNCH((positive_terms_kernel<G, T>), tile_grid, PAA_TILE, stream, geo, gop, gt_boxes, sc,       \
        ws.paa_label, ws.matched, normalisers, ws.local_norm, grad_losses, tile_part, tiles_total,           \
        tiles_per_block, zero_fill, has_ignored && !bulk_ignores)
    {
        KernelTimer timer(PAA_KERNEL_POSITIVE_TERMS, stream);
        if (write_grads) {
            if (g2) PAA_POS(true, true); else PAA_POS(true, false);
        } else {
            if (g2) PAA_POS(false, true); else PAA_POS(false, false);
        }
    }
#undef PAA_POS
    {
        KernelTimer timer(PAA_KERNEL_FINISH_LOSS, stream);
        PAA_PDL_LAUNCH(finish_loss_kernel, 1, kFinishThreads, stream, bulk_part, bulk_grid, tile_part, tile_grid, sc,
                       normalisers, losses);
    }
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
