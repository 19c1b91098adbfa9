// PAA anchor assignment on sm_100a: IoU matching, anchor scores, per-GT top-k + GMM.
//
// Replaces, for the images of one rank and without ever materialising the [G, A] IoU matrix:
//   prepare_iou_based_targets   paa_core/modeling/rpn/paa/loss.py:89-126
//     boxlist_iou               paa_core/structures/boxlist_ops.py:81-116
//     Matcher(thr, thr, True)   paa_core/modeling/matcher.py:42-113
//   anchor scores               loss.py:293-306 (focal sum + GIoU on IoU-positive anchors)
//   compute_paa                 loss.py:128-236 incl. sklearn GaussianMixture (loss.py:197-203;
//                               scikit-learn 1.9.0 semantics restated in oracle/gmm_oracle.py)
//   normalisers                 loss.py:320-322,331-333,338 (this rank's partial sums)
//
// Data layout: head tensors are consumed in place as NCHW (one anchor per location makes the class
// stride H*W, so a warp reading 32 consecutive anchors of one class is one 128-byte line); all
// per-anchor intermediates are flat [N*A] arrays in the caller's workspace.
#include "kernels.h"
#include "fastmath64.cuh"

namespace paa {

// ---------------------------------------------------------------------------------------------
// K1: every anchor's best GT (first maximum) and every GT's maximal IoU.
// One block = one tile of 128 consecutive anchors of one level of one image.  GT boxes are staged
// in shared memory in chunks; a warp only evaluates the GTs whose box intersects the warp's
// bounding box (non-intersecting pairs have IoU exactly +0 and can neither raise a maximum nor
// win the first-maximum rule against the initial (0, GT 0)).
// ---------------------------------------------------------------------------------------------
constexpr int kGtChunk = 256;

__global__ void __launch_bounds__(PAA_TILE)
iou_best_kernel(const Geometry geo, const GtOffsets go, const float* __restrict__ gt_boxes,
                unsigned* __restrict__ gtmax, uint2* __restrict__ best) {
    __shared__ float4 s_gt[kGtChunk];
    __shared__ float s_area[kGtChunk];
    __shared__ unsigned s_max[kGtChunk];

    // heaviest blocks first: the coarse levels (last tiles of an image) intersect every GT, and the
    // cost of a block grows with the GT count of its image
    const int n = go.by_load[blockIdx.x % geo.num_images];
    const int tile = geo.tiles_per_image - 1 - blockIdx.x / geo.num_images;
    int first;
    const int l = tile_level(geo, tile, &first);
    const LevelView& lv = geo.lv[l];
    const int i = first + threadIdx.x;
    const bool valid = i < lv.n_anchor;
    const int lane = threadIdx.x & 31;

    float4 a = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
    if (valid) a = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
    const float area_a = area_plus1(a);
    const float wx1 = warp_min(a.x), wy1 = warp_min(a.y), wx2 = warp_max(a.z), wy2 = warp_max(a.w);

    const int gbase = go.v[n];
    const int G = go.v[n + 1] - gbase;
    float best_v = 0.0f;
    int best_g = 0;

    for (int c0 = 0; c0 < G; c0 += kGtChunk) {
        const int cnt = min(kGtChunk, G - c0);
        for (int t = threadIdx.x; t < cnt; t += PAA_TILE) {
            float4 b = ldg4(gt_boxes + (size_t)(gbase + c0 + t) * 4);
            s_gt[t] = b;
            s_area[t] = area_plus1(b);
            s_max[t] = 0u;
        }
        __syncthreads();
        for (int g0 = 0; g0 < cnt; g0 += PAA_WARP) {
            const int g = g0 + lane;
            bool hit = false;
            if (g < cnt) {
                float4 b = s_gt[g];
                float w = __fadd_rn(__fsub_rn(fminf(b.z, wx2), fmaxf(b.x, wx1)), 1.0f);
                float h = __fadd_rn(__fsub_rn(fminf(b.w, wy2), fmaxf(b.y, wy1)), 1.0f);
                hit = (w > 0.0f) && (h > 0.0f);
            }
            unsigned m = __ballot_sync(PAA_FULL, hit);
            while (m) {
                const int j = g0 + __ffs(m) - 1;
                m &= m - 1;
                float q = 0.0f;
                if (valid) q = iou_plus1(s_gt[j], s_area[j], a, area_a);
                if (q > best_v) {
                    best_v = q;
                    best_g = c0 + j;
                }
                unsigned wm = __reduce_max_sync(PAA_FULL, __float_as_uint(q));
                if (lane == 0 && wm != 0u) atomicMax(&s_max[j], wm);
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += PAA_TILE)
            if (s_max[t] != 0u) atomicMax(&gtmax[gbase + c0 + t], s_max[t]);
        __syncthreads();
    }
    if (valid) best[(size_t)n * geo.A + lv.a_off + i] = make_uint2(__float_as_uint(best_v), (unsigned)best_g);
}

int launch_iou_best(const Geometry& geo, const GtOffsets& go, const float* gt_boxes,
                    const LossWorkspace& ws, cudaStream_t stream) {
    int grid = geo.num_images * geo.tiles_per_image;
    KernelTimer timer(PAA_KERNEL_IOU_BEST, stream);
    iou_best_kernel<<<grid, PAA_TILE, 0, stream>>>(geo, go, gt_boxes, ws.gtmax, ws.best);
    PAA_LAUNCH_CHECK("iou_best_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K2: Matcher decision + IoU-based label + anchor score of IoU-positive anchors.
//   matched = argmax GT if max IoU >= thr, else -1, except that an anchor which is some GT's best
//   anchor (IoU == that GT's maximum, ties included) keeps its argmax GT (matcher.py:83-113).
//   Only GTs whose maximum is below thr can restore anything, so those are listed first.
//   score   = sum_c focal(logit_c | IoU label) + (1 - GIoU(decode(pred), decode(encode(gt))))
//   (loss.py:293-306; anchors without an IoU-positive label are never candidates, their 1e8 filler
//   is not materialised).
// ---------------------------------------------------------------------------------------------
// order-preserving map float -> unsigned (and back), so that (score, anchor) pairs compare as integers
__device__ __forceinline__ unsigned ordered_bits(float v) {
    unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// sum_c focal(x_c | label) for one anchor; `p` points at class 0, consecutive classes are `stride` apart.
// Loads are issued eight at a time before any of them is consumed.  Every class is first taken as a
// negative (lean math), then the labelled class's negative term is swapped for the positive one
// (accurate path for that single term).
__device__ __forceinline__ float neg_term_only(float x, float gamma, bool g2) {
    const SigmoidLean sl = sigmoid_lean(x);
    return focal_pow(sl.p, gamma, g2) * sl.sp;
}

__device__ __forceinline__ float focal_sum(const float* __restrict__ p, int stride, int C, int label,
                                           float gamma, float alpha) {
    const bool g2 = (gamma == 2.0f);
    const float oma = 1.0f - alpha;
    float neg = 0.0f;
    const unsigned st = (unsigned)stride;
    float xl = 0.0f;
    for (int c0 = 0; c0 < C; c0 += 8) {
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (c0 + j < C) ? __ldg(p + (unsigned)(c0 + j) * st) : -100.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {      // a padded logit of -100 contributes exactly 0
            neg += neg_term_only(x[j], gamma, g2);
            if (c0 + j + 1 == label) xl = x[j];
        }
    }
    const SigmoidParts sp = sigmoid_parts(xl);
    float tp, gp;
    focal_positive(xl, sp, gamma, g2, alpha, &tp, &gp);
    return fmaf(oma, neg - neg_term_only(xl, gamma, g2), tp);
}

__global__ void __launch_bounds__(PAA_TILE)
match_score_kernel(const Geometry geo, const GtOffsets go, const float* __restrict__ gt_boxes,
                   const int64_t* __restrict__ gt_labels, const unsigned* __restrict__ gtmax,
                   const uint2* __restrict__ best, const LossScalars sc, int* __restrict__ matched,
                   float* __restrict__ score, int* __restrict__ paa_label, uint4* __restrict__ tile_gtmask,
                   int* __restrict__ seg_count, unsigned long long* __restrict__ seg_pool,
                   const float* __restrict__ teacher_score, const LossDebug dbg) {
    __shared__ int s_lq[PAA_TILE];
    __shared__ int s_nlq;
    __shared__ unsigned s_mask[4];
    if (threadIdx.x < 4) s_mask[threadIdx.x] = 0u;

    const int n = blockIdx.x / geo.tiles_per_image;
    const int tile = blockIdx.x - n * geo.tiles_per_image;
    int first;
    const int l = tile_level(geo, tile, &first);
    const LevelView& lv = geo.lv[l];
    const int i = first + threadIdx.x;
    const bool valid = i < lv.n_anchor;
    const int gbase = go.v[n];
    const int G = go.v[n + 1] - gbase;
    const float thr = sc.iou_threshold;

    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    uint2 bv = make_uint2(0u, 0u);
    const size_t flat = (size_t)n * geo.A + lv.a_off + (valid ? i : 0);
    if (valid) {
        a = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
        bv = best[flat];
    }
    const float area_a = area_plus1(a);
    const float bval = __uint_as_float(bv.x);
    int m = (bval >= thr) ? (int)bv.y : -1;
    if (!(bval >= thr) && !(bval < thr)) m = (int)bv.y;   // NaN: neither below nor restored -> keeps argmax

    // low-quality GTs (max IoU below thr) restore their best anchors
    for (int c0 = 0; c0 < G; c0 += PAA_TILE) {
        if (threadIdx.x == 0) s_nlq = 0;
        __syncthreads();
        const int g = c0 + threadIdx.x;
        if (g < G) {
            const unsigned u = gtmax[gbase + g];
            if (__uint_as_float(u) < thr) s_lq[atomicAdd(&s_nlq, 1)] = g;
        }
        __syncthreads();
        const int nlq = s_nlq;
        if (valid && m < 0) {
            for (int k = 0; k < nlq; ++k) {
                const int gg = s_lq[k];
                const float4 b = ldg4(gt_boxes + (size_t)(gbase + gg) * 4);
                const float q = iou_plus1(b, area_plus1(b), a, area_a);
                if (q == __uint_as_float(gtmax[gbase + gg])) m = (int)bv.y;
            }
        }
        __syncthreads();
    }
    // which GTs (index mod 128) have matched anchors in this tile: lets the per-GT selection kernel
    // skip every tile that cannot contain one of its anchors, with no false negatives
    if (valid && m >= 0) atomicOr(&s_mask[(m & 127) >> 5], 1u << (m & 31));
    __syncthreads();
    if (threadIdx.x == 0)
        tile_gtmask[(size_t)n * geo.tiles_per_image + tile] = make_uint4(s_mask[0], s_mask[1], s_mask[2], s_mask[3]);
    if (!valid) return;

    int label = 0;
    if (m >= 0) label = (int)gt_labels[gbase + m];
    matched[flat] = m;
    paa_label[flat] = 0;
    if (dbg.matched_idx) dbg.matched_idx[flat] = m;
    if (dbg.iou_labels) dbg.iou_labels[flat] = label;

    float s = 1.0e8f;   // loss.py:15,301-306 filler, only visible through the debug output
    if (m >= 0 && label > 0) {
        const float* cls = lv.cls + head_offset(n, i, 0, geo.C, geo.apl, lv.hw);
        const float fsum = focal_sum(cls, lv.hw, geo.C, label, sc.gamma, sc.alpha);
        const float* rp = lv.reg + head_offset(n, i, 0, 4, geo.apl, lv.hw);
        const float4 d = make_float4(__ldg(rp), __ldg(rp + lv.hw), __ldg(rp + 2 * (size_t)lv.hw),
                                     __ldg(rp + 3 * (size_t)lv.hw));
        const AnchorFrame f = anchor_frame(a);
        const float4 pred = decode_box(d, f);
        const float4 gt = ldg4(gt_boxes + (size_t)(gbase + m) * 4);
        const float4 tgt = decode_box(encode_box(gt, f), f);
        s = __fadd_rn(fsum, giou_loss_boxes(pred, tgt));
        score[flat] = s;
        // hand the anchor to its (GT, level) candidate pool: the per-GT selection then reads one short
        // contiguous list instead of searching the image for its anchors
        const float key_score = teacher_score ? teacher_score[flat] : s;
        const int seg = (gbase + m) * geo.num_levels + l;
        const int slot = atomicAdd(&seg_count[seg], 1);
        if (slot < sc.seg_cap)
            seg_pool[(size_t)seg * kSegCap + slot] =
                ((unsigned long long)ordered_bits(key_score) << 32) | (unsigned)(lv.a_off + i);
    }
    if (dbg.combined_loss) dbg.combined_loss[flat] = s;
}

int launch_match_score(const Geometry& geo, const GtOffsets& go, const float* gt_boxes,
                       const int64_t* gt_labels, const LossScalars& sc, const LossWorkspace& ws,
                       const float* teacher_score, const LossDebug& dbg, cudaStream_t stream) {
    int grid = geo.num_images * geo.tiles_per_image;
    KernelTimer timer(PAA_KERNEL_MATCH_SCORE, stream);
    match_score_kernel<<<grid, PAA_TILE, 0, stream>>>(geo, go, gt_boxes, gt_labels, ws.gtmax, ws.best, sc,
                                                      ws.matched, ws.score, ws.paa_label, ws.tile_gtmask,
                                                      ws.seg_count, ws.seg_pool, teacher_score, dbg);
    PAA_LAUNCH_CHECK("match_score_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K3: one warp per GT -- per-level top-k candidates, sort, two-component GMM, labels, partial
// normalisers.  loss.py:151-236 + sklearn GaussianMixture.fit/predict/score_samples.
// ---------------------------------------------------------------------------------------------
constexpr int kSelWarps = 4;                // GTs per block
constexpr unsigned long long kEmptyKey = ~0ull;

// sklearn/utils/_array_api.py:1338-1366 for two entries.
__device__ __forceinline__ double logsumexp2(double a0, double a1) {
    if (a0 == a1) return (log1p(0.0) + 0.6931471805599453094) + a0;   // m = 2, masked sum = 0
    const double hi = fmax(a0, a1), lo = fmin(a0, a1);
    return (log1p(exp(lo - hi)) + 0.0) + hi;                          // m = 1, log(m) = 0
}
// Same, also returning the responsibilities exp(a_k - lse).  They are formed as 1/(1+s) and s/(1+s)
// with s = exp(lo - hi): one exp and one division instead of three exps; the results differ from
// sklearn's exp(a_k - lse) by a few 1e-16 relative, far below anything the float32 roundings of
// the next E-step can see.
__device__ __forceinline__ double logsumexp2_resp(double a0, double a1, double* r0, double* r1) {
    if (a0 == a1) {
        *r0 = 0.5;
        *r1 = 0.5;
        return (log1p(0.0) + 0.6931471805599453094) + a0;
    }
    const bool first_hi = a0 > a1;
    const double hi = first_hi ? a0 : a1, lo = first_hi ? a1 : a0;
    const double s = exp_nonpos(lo - hi);
    const double rh = div_fast(1.0, 1.0 + s), rl = s * rh;
    *r0 = first_hi ? rh : rl;
    *r1 = first_hi ? rl : rh;
    return (log1p_unit(s) + 0.0) + hi;
}

struct GmmState {
    double w0, w1, mu0, mu1;
    float pc0, pc1, var0, var1;
    bool first;                    // precisions still the float64 initial values (1.0)
};

// a_k = log N(x | mu_k, pc_k) + log w_k with the dtype flow of _estimate_log_gaussian_prob
// (sklearn/mixture/_gaussian_mixture.py:490-553): float32 product x*pc and float32 log-det once the
// precisions have been re-estimated, float32 buffer for the squared distance, float64 elsewhere.
__device__ __forceinline__ void weighted_log_prob(float x, const GmmState& s, double lw0, double lw1,
                                                  float ld0, float ld1, double* a0, double* a1) {
    const float LOG2PI = 1.8378770664093453f;
    double xs0, xs1;
    if (s.first) {
        xs0 = (double)x;
        xs1 = (double)x;
    } else {
        xs0 = (double)__fmul_rn(x, s.pc0);
        xs1 = (double)__fmul_rn(x, s.pc1);
    }
    const double y0 = xs0 - s.mu0 * (double)s.pc0;
    const double y1 = xs1 - s.mu1 * (double)s.pc1;
    const float q0 = __double2float_rn(y0 * y0);
    const float q1 = __double2float_rn(y1 * y1);
    const float lp0 = __fmul_rn(-0.5f, __fadd_rn(LOG2PI, q0));
    const float lp1 = __fmul_rn(-0.5f, __fadd_rn(LOG2PI, q1));
    if (s.first) {                    // float32 array + float64 log-det (log 1.0 = 0)
        *a0 = ((double)lp0 + 0.0) + lw0;
        *a1 = ((double)lp1 + 0.0) + lw1;
    } else {                          // float32 log-det, float32 add
        *a0 = (double)__fadd_rn(lp0, ld0) + lw0;
        *a1 = (double)__fadd_rn(lp1, ld1) + lw1;
    }
}

// Fits the mixture on the warp's n sorted samples (lane holds x[lane + 32*k]) and returns the length
// of the positive prefix (loss.py:206-217).  out8 (nullable, lane 0 writes) receives the parameters.
template <int SPL>
__device__ int gmm_positive_prefix(const float (&x)[SPL], int n, int lane, double* out8) {
    GmmState s;
    s.w0 = 0.5;
    s.w1 = 0.5;
    s.mu0 = (double)__shfl_sync(PAA_FULL, x[0], 0);                       // min (sorted ascending)
    {
        const int last = n - 1;
        float xl = 0.f;
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            float v = __shfl_sync(PAA_FULL, x[k], last & 31);
            if ((last >> 5) == k) xl = v;
        }
        s.mu1 = (double)xl;                                               // max
    }
    s.pc0 = s.pc1 = 1.0f;
    s.var0 = s.var1 = 1.0f;
    s.first = true;
    double lower = -INFINITY;
    int n_iter = 0;
    bool converged = false;
    const double EPS10 = 10.0 * 2.220446049250313e-16;
    for (int it = 1; it <= 100; ++it) {
        n_iter = it;
        const double lw0 = log_pos(s.w0), lw1 = log_pos(s.w1);
        const float ld0 = s.first ? 0.f : __double2float_rn(log_pos((double)s.pc0));
        const float ld1 = s.first ? 0.f : __double2float_rn(log_pos((double)s.pc1));
        // One reduction round per iteration: moments are taken about the PREVIOUS means (p0, p1), from
        // which the new mean and the variance about the new mean follow exactly:
        //   sum r x = A + p S,   sum r (x - mu)^2 = B - 2 d A + d^2 S   with d = mu - p.
        // (sklearn evaluates sum r (x - mu)^2 directly; the two agree to ~1e-14 relative, far below the
        // float32 rounding of the variance that follows.)
        const double p0 = s.mu0, p1 = s.mu1;
        double S0 = 0, S1 = 0, A0 = 0, A1 = 0, B0 = 0, B1 = 0, s_lpn = 0;
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            if (lane + 32 * k < n) {
                double a0, a1, r0, r1;
                weighted_log_prob(x[k], s, lw0, lw1, ld0, ld1, &a0, &a1);
                s_lpn += logsumexp2_resp(a0, a1, &r0, &r1);
                const double e0 = (double)x[k] - p0, e1 = (double)x[k] - p1;
                const double t0 = r0 * e0, t1 = r1 * e1;
                S0 += r0;
                S1 += r1;
                A0 += t0;
                A1 += t1;
                B0 = fma(t0, e0, B0);
                B1 = fma(t1, e1, B1);
            }
        }
        S0 = warp_sum(S0);
        S1 = warp_sum(S1);
        A0 = warp_sum(A0);
        A1 = warp_sum(A1);
        B0 = warp_sum(B0);
        B1 = warp_sum(B1);
        s_lpn = warp_sum(s_lpn);
        const double nk0 = S0 + EPS10, nk1 = S1 + EPS10;
        s.mu0 = div_fast(fma(p0, S0, A0), nk0);
        s.mu1 = div_fast(fma(p1, S1, A1), nk1);
        const double d0 = s.mu0 - p0, d1 = s.mu1 - p1;
        const double c0 = fma(d0, fma(d0, S0, -2.0 * A0), B0);
        const double c1 = fma(d1, fma(d1, S1, -2.0 * A1), B1);
        s.var0 = __fadd_rn(__double2float_rn(div_fast(fmax(c0, 0.0), nk0)), 1e-6f);
        s.var1 = __fadd_rn(__double2float_rn(div_fast(fmax(c1, 0.0), nk1)), 1e-6f);
        const double nsum = nk0 + nk1;
        s.w0 = div_fast(nk0, nsum);
        s.w1 = div_fast(nk1, nsum);
        s.pc0 = __fdiv_rn(1.0f, __fsqrt_rn(s.var0));
        s.pc1 = __fdiv_rn(1.0f, __fsqrt_rn(s.var1));
        s.first = false;
        const double lb = div_fast(s_lpn, (double)n);
        const double change = lb - lower;
        lower = lb;
        if (fabs(change) < 1e-3) {
            converged = true;
            break;
        }
    }
    // final E-step: predict (argmax, ties -> 0) and score_samples
    const double lw0 = log(s.w0), lw1 = log(s.w1);
    const float ld0 = __double2float_rn(log((double)s.pc0));
    const float ld1 = __double2float_rn(log((double)s.pc1));
    double best_score = -INFINITY;
    int best_idx = 0x7fffffff;
    bool any_fg = false;
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        if (lane + 32 * k < n) {
            double a0, a1;
            weighted_log_prob(x[k], s, lw0, lw1, ld0, ld1, &a0, &a1);
            if (!(a1 > a0)) {                 // component 0 == foreground (loss.py:206)
                const double sc = logsumexp2(a0, a1);
                any_fg = true;
                if (sc > best_score) {        // k ascending => first index among equals kept
                    best_score = sc;
                    best_idx = lane + 32 * k;
                }
            }
        }
    }
    // warp arg-max with smallest index among equal scores
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(PAA_FULL, best_score, o);
        const int oi = __shfl_xor_sync(PAA_FULL, best_idx, o);
        if (os > best_score || (os == best_score && oi < best_idx)) {
            best_score = os;
            best_idx = oi;
        }
    }
    any_fg = __any_sync(PAA_FULL, any_fg);
    if (out8 && lane == 0) {
        out8[0] = s.w0;
        out8[1] = s.w1;
        out8[2] = s.mu0;
        out8[3] = s.mu1;
        out8[4] = (double)s.var0;
        out8[5] = (double)s.var1;
        out8[6] = (double)n_iter;
        out8[7] = converged ? 1.0 : 0.0;
    }
    return any_fg ? best_idx + 1 : n;
}

// Offers `key` (valid on lanes where `is`) to the warp's running top-K list: lane j < K holds the j-th
// smallest key so far.  Keys that cannot enter the list are dropped before the serial insertion (the
// K-th key only ever decreases, so comparing against a stale bound is safe).
__device__ __forceinline__ void topk_offer(unsigned long long& mine, unsigned long long key, bool is, int K,
                                           int lane) {
    const unsigned long long kth = __shfl_sync(PAA_FULL, mine, K - 1);
    is = is && (key < kth);
    unsigned hm = __ballot_sync(PAA_FULL, is);
    while (hm) {
        const int src = __ffs(hm) - 1;
        hm &= hm - 1;
        const unsigned long long nk = __shfl_sync(PAA_FULL, key, src);
        const unsigned less = __ballot_sync(PAA_FULL, lane < K && mine < nk);
        const int pos = __popc(less);
        const unsigned long long up = __shfl_up_sync(PAA_FULL, mine, 1);
        if (pos < K) {
            if (lane == pos) mine = nk;
            else if (lane > pos && lane < K) mine = up;
        }
    }
}

template <int SPL>
__global__ void __launch_bounds__(kSelWarps * PAA_WARP)
select_gmm_kernel(const Geometry geo, const GtOffsets go, int num_gt_total,
                  const float* __restrict__ gt_boxes, const int64_t* __restrict__ gt_labels,
                  const LossScalars sc, const uint4* __restrict__ tile_gtmask,
                  const int* __restrict__ matched, const float* __restrict__ score,
                  const int* __restrict__ seg_count, const unsigned long long* __restrict__ seg_pool,
                  int* __restrict__ paa_label,
                  int* __restrict__ part_npos, double* __restrict__ part_siou,
                  unsigned* __restrict__ ticket, double* __restrict__ local_norm,
                  double* __restrict__ normalisers, const LossDebug dbg) {
    __shared__ unsigned long long s_key[kSelWarps][PAA_MAX_CANDIDATES];
    __shared__ unsigned long long s_sorted[kSelWarps][PAA_MAX_CANDIDATES];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gi = blockIdx.x * kSelWarps + warp;
    const int K = sc.topk;
    const int cap = geo.num_levels * K;
    int n_pos = 0;
    double siou = 0.0;
    if (gi < num_gt_total) {
        // image of this GT
        int n = 0;
        for (int k = 1; k < geo.num_images; ++k)
            if (gi >= go.v[k]) n = k;
        const int g_local = gi - go.v[n];
        const float4 gt = ldg4(gt_boxes + (size_t)gi * 4);
        const int cls_label = (int)gt_labels[gi];
        const unsigned gbit = 1u << (g_local & 31);
        const int gword = (g_local & 127) >> 5;
        const int* mrow = matched + (size_t)n * geo.A;
        const float* srow = score + (size_t)n * geo.A;
        unsigned long long* keys = s_key[warp];
        int n_cand = 0;

#ifdef PAA_PROFILE_GMM
        const long long prof_t0 = clock64();
#endif
        if (cls_label > 0) {        // loss.py:166 requires a positive IoU label
            for (int l = 0; l < geo.num_levels; ++l) {
                const LevelView& lv = geo.lv[l];
                const int t_end = (l + 1 < geo.num_levels) ? geo.lv[l + 1].tile_off : geo.tiles_per_image;
                unsigned long long mine = kEmptyKey;   // lane j < K holds the j-th smallest key so far
                const int seg = gi * geo.num_levels + l;
                const int seg_n = __ldg(seg_count + seg);
                if (seg_n <= sc.seg_cap) {
                    // normal case: the anchors matched to this (GT, level) were pooled by match_score_kernel
                    const unsigned long long* pool = seg_pool + (size_t)seg * kSegCap;
                    for (int j0 = 0; j0 < seg_n; j0 += 4 * PAA_WARP) {
                        unsigned long long k4[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int j = j0 + r * PAA_WARP + lane;
                            k4[r] = (j < seg_n) ? __ldg(pool + j) : kEmptyKey;
                        }
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            if (j0 + r * PAA_WARP < seg_n) topk_offer(mine, k4[r], k4[r] != kEmptyKey, K, lane);
                    }
                } else
                // overflowed pool: search the tiles of this level that contain anchors of this GT
                for (int t0 = lv.tile_off; t0 < t_end; t0 += PAA_WARP) {
                    const int t = t0 + lane;
                    bool hit = false;
                    if (t < t_end) {
                        const uint4 gm = __ldg(tile_gtmask + (size_t)n * geo.tiles_per_image + t);
                        const unsigned word = gword == 0 ? gm.x : (gword == 1 ? gm.y : (gword == 2 ? gm.z : gm.w));
                        hit = (word & gbit) != 0u;
                    }
                    unsigned tm = __ballot_sync(PAA_FULL, hit);
                    while (tm) {
                        const int tt = t0 + __ffs(tm) - 1;
                        tm &= tm - 1;
                        const int base = (tt - lv.tile_off) * PAA_TILE;
                        int mv[PAA_TILE / PAA_WARP];
                        float sv[PAA_TILE / PAA_WARP];
#pragma unroll
                        for (int r = 0; r < PAA_TILE / PAA_WARP; ++r) {     // 8 independent loads in flight
                            const int i = base + r * PAA_WARP + lane;
                            const bool in = i < lv.n_anchor;
                            mv[r] = in ? __ldg(mrow + lv.a_off + i) : -2;
                            sv[r] = in ? __ldg(srow + lv.a_off + i) : 0.0f;
                        }
#pragma unroll
                        for (int r = 0; r < PAA_TILE / PAA_WARP; ++r) {
                            const int i = base + r * PAA_WARP + lane;
                            bool is = (mv[r] == g_local);
                            unsigned long long key = kEmptyKey;
                            if (is) key = ((unsigned long long)ordered_bits(sv[r]) << 32) | (unsigned)(lv.a_off + i);
                            topk_offer(mine, key, is, K, lane);
                        }
                    }
                }
                // append this level's candidates (ascending key) to the warp's list
                const unsigned have = __ballot_sync(PAA_FULL, lane < K && mine != kEmptyKey);
                const int cnt = __popc(have);
                if (lane < cnt) keys[n_cand + lane] = mine;
                n_cand += cnt;
            }
        }
        __syncwarp();

        // sort all candidates by (loss, index): rank by counting, n_cand <= cap <= 128
        unsigned long long* sorted = s_sorted[warp];
        for (int j = lane; j < n_cand; j += PAA_WARP) {
            const unsigned long long kj = keys[j];
            int rank = 0;
            for (int q = 0; q < n_cand; ++q) rank += (keys[q] < kj) ? 1 : 0;
            sorted[rank] = kj;
        }
        __syncwarp();

#ifdef PAA_PROFILE_GMM
        const long long prof_t1 = clock64();
#endif
        if (n_cand == 1) {
            n_pos = 1;                                     // loss.py:218-219
        } else if (n_cand > 1) {
            float x[SPL];
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                const int j = lane + 32 * k;
                x[k] = (j < n_cand) ? from_ordered_bits((unsigned)(sorted[j] >> 32)) : 0.f;
            }
            n_pos = gmm_positive_prefix<SPL>(x, n_cand, lane, dbg.gmm ? dbg.gmm + (size_t)gi * 8 : nullptr);
        }
        if (dbg.gmm && n_cand <= 1 && lane < 8) dbg.gmm[(size_t)gi * 8 + lane] = 0.0;
#ifdef PAA_PROFILE_GMM        // measurement build only: cycles of scan+sort / EM in the w0 / w1 debug slots
        if (dbg.gmm && lane == 0) {
            dbg.gmm[(size_t)gi * 8 + 0] = (double)(prof_t1 - prof_t0);
            dbg.gmm[(size_t)gi * 8 + 1] = (double)(clock64() - prof_t1);
        }
#endif

        // labels of the positive prefix + this GT's share of the IoU normaliser (loss.py:228-230,331-333)
        for (int j = lane; j < n_cand; j += PAA_WARP) {
            const int aidx = (int)(sorted[j] & 0xffffffffu);
            if (dbg.cand_idx) dbg.cand_idx[(size_t)gi * cap + j] = aidx;
            if (j < n_pos) {
                paa_label[(size_t)n * geo.A + aidx] = cls_label;
                if (dbg.paa_labels) dbg.paa_labels[(size_t)n * geo.A + aidx] = cls_label;
                if (sc.use_iou_pred) {
                    const int l = anchor_level(geo, aidx);
                    const LevelView& lv = geo.lv[l];
                    const int i = aidx - lv.a_off;
                    const float4 a = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
                    const AnchorFrame f = anchor_frame(a);
                    const float* rp = lv.reg + head_offset(n, i, 0, 4, geo.apl, lv.hw);
                    const float4 d = make_float4(__ldg(rp), __ldg(rp + lv.hw), __ldg(rp + 2 * (size_t)lv.hw),
                                                 __ldg(rp + 3 * (size_t)lv.hw));
                    const float4 pred = decode_box(d, f);
                    const float4 tgt = decode_box(encode_box(gt, f), f);
                    siou += (double)iou_plus1(tgt, area_plus1(tgt), pred, area_plus1(pred));
                }
            }
        }
        siou = warp_sum(siou);
        if (lane == 0) {
            part_npos[gi] = n_pos;
            part_siou[gi] = siou;
            if (dbg.cand_cnt) dbg.cand_cnt[gi] = n_cand;
            if (dbg.num_pos) dbg.num_pos[gi] = n_pos;
        }
    }

    // the last warp to finish folds the per-GT partials in a fixed order (deterministic sums)
    const unsigned total_warps = gridDim.x * kSelWarps;
    unsigned my_ticket = 0;
    if (lane == 0) {
        __threadfence();
        my_ticket = atomicAdd(&ticket[0], 1u);
    }
    my_ticket = __shfl_sync(PAA_FULL, my_ticket, 0);
    if (my_ticket == total_warps - 1) {
        __threadfence();
        double cnt = 0.0, sum = 0.0;
        for (int g = lane; g < num_gt_total; g += PAA_WARP) {
            cnt += (double)__ldcg(part_npos + g);
            sum += __ldcg(part_siou + g);
        }
        cnt = warp_sum(cnt);
        sum = warp_sum(sum);
        if (lane == 0) {
            local_norm[0] = cnt;
            local_norm[1] = sum;
            normalisers[0] = cnt;
            normalisers[1] = sum;
        }
    }
}

int launch_select_gmm(const Geometry& geo, const GtOffsets& go, int num_gt_total, const float* gt_boxes,
                      const int64_t* gt_labels, const LossScalars& sc, const LossWorkspace& ws,
                      const float* score_src, double* normalisers, const LossDebug& dbg,
                      cudaStream_t stream) {
    const int cap = geo.num_levels * sc.topk;
    const int grid = (num_gt_total + kSelWarps - 1) / kSelWarps;
    const int threads = kSelWarps * PAA_WARP;
    KernelTimer timer(PAA_KERNEL_SELECT_GMM, stream);
#define PAA_SEL_LAUNCH(SPL)                                                                           \
    select_gmm_kernel<SPL><<<grid, threads, 0, stream>>>(geo, go, num_gt_total, gt_boxes, gt_labels, sc, \
        ws.tile_gtmask, ws.matched, score_src, ws.seg_count, ws.seg_pool, ws.paa_label, ws.part_npos,  \
        ws.part_siou,                                                                                  \
        ws.ticket, ws.local_norm, normalisers, dbg)
    if (cap <= 32) PAA_SEL_LAUNCH(1);
    else if (cap <= 64) PAA_SEL_LAUNCH(2);
    else PAA_SEL_LAUNCH(4);
#undef PAA_SEL_LAUNCH
    PAA_LAUNCH_CHECK("select_gmm_kernel");
    return 0;
}

}  // namespace paa
