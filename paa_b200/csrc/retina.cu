// RetinaNet anchor labelling on sm_100a: rpn/loss.py:41-88 of the reference as used by
// rpn/retinanet/loss.py:45-56, in place of PAA's score / top-k / GMM stages.
//
//   assign_pass1_kernel (IoU half only)  every anchor's best GT and every GT's best IoU, exactly as for PAA
//                                        (boxlist_ops.py:81-116 + the two maxima of matcher.py:66,92); it is
//                                        indifferent to the number of anchors per location.
//   retina_labels_kernel                 one thread per anchor: Matcher decision with separate high / low
//                                        thresholds and the low-quality restore (matcher.py:66-113), label =
//                                        class of the matched GT / 0 below low / -1 between the thresholds
//                                        (rpn/loss.py:64-80), per-tile positive counts.
//   atss_norm_kernel                     fixed-order fold of the counts into normalisers[0] (no peer exchange:
//                                        the reference normalises by the rank's own num_pos + N).
// The regression targets (box_coder.py:22-50) are not materialised: the loss pass encodes the matched GT of the
// few positive anchors on the fly.
#include "kernels.h"

namespace paa {

__global__ void __launch_bounds__(PAA_TILE)
retina_labels_kernel(const Geometry geo, const GtOffsets go, const float* __restrict__ gt_boxes,
                     const int64_t* __restrict__ gt_labels, const unsigned* __restrict__ gtmax,
                     const unsigned long long* __restrict__ best, const LossScalars sc,
                     int* __restrict__ matched, int* __restrict__ label_out, double* __restrict__ tile_part,
                     const LossDebug dbg) {
    __shared__ int s_lq[PAA_TILE];
    __shared__ int s_nlq;
    __shared__ double s_part[PAA_TILE / PAA_WARP];

    const int n = blockIdx.x / geo.tiles_per_image;
    const int tile = blockIdx.x - n * geo.tiles_per_image;
    int first;
    const int l = tile_level(geo, tile, &first);
    const LevelView& lv = geo.lv[l];
    const int i = first + threadIdx.x;
    const bool valid = i < lv.n_anchor;
    const int gbase = go.v[n];
    const int G = go.v[n + 1] - gbase;
    const float high = sc.iou_threshold, low = sc.bg_threshold;

    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    float bval = 0.0f;
    int bgt = 0;
    const size_t flat = (size_t)n * geo.A + lv.a_off + (valid ? i : 0);
    if (valid) {
        a = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
        unpack_best(best[flat], &bval, &bgt);
    }
    const float area_a = area_plus1(a);
    int m = bgt;                                              // NaN: in neither set, keeps the argmax
    if (bval < low) m = -1;                                   // matcher.py:71   BELOW_LOW_THRESHOLD
    else if (bval >= low && bval < high) m = -2;              // matcher.py:72-74 BETWEEN_THRESHOLDS

    // A GT whose best IoU stays below `high` restores its best anchors (ties included) to their own argmax
    // (matcher.py:83-113); the best anchors of every other GT are matched already.
    for (int c0 = 0; c0 < G; c0 += PAA_TILE) {
        if (threadIdx.x == 0) s_nlq = 0;
        __syncthreads();
        const int g = c0 + threadIdx.x;
        if (g < G && __uint_as_float(gtmax[gbase + g]) < high) s_lq[atomicAdd(&s_nlq, 1)] = g;
        __syncthreads();
        const int nlq = s_nlq;
        if (valid && m < 0) {
            for (int k = 0; k < nlq; ++k) {
                const int gg = s_lq[k];
                const float4 b = ldg4(gt_boxes + (size_t)(gbase + gg) * 4);
                const float q = iou_plus1(b, area_plus1(b), a, area_a);
                if (q == __uint_as_float(gtmax[gbase + gg])) m = bgt;
            }
        }
        __syncthreads();
    }

    double npos = 0.0;
    if (valid) {
        int label = 0;                                        // rpn/loss.py:68-69
        if (m >= 0) label = (int)gt_labels[gbase + m];        // generate_retinanet_labels, retinanet/loss.py:84-86
        else if (m == -2) label = -1;                         // rpn/loss.py:76-78
        matched[flat] = m < 0 ? 0 : m;                        // target[matched_idxs.clamp(min=0)], rpn/loss.py:52
        label_out[flat] = label;
        if (dbg.matched_idx) dbg.matched_idx[flat] = m;
        if (dbg.iou_labels) dbg.iou_labels[flat] = label;
        npos = label > 0 ? 1.0 : 0.0;
    }
    npos = warp_sum(npos);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_part[warp] = npos;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < PAA_TILE / PAA_WARP; ++w) t += s_part[w];
        tile_part[(size_t)blockIdx.x * 2] = t;
        tile_part[(size_t)blockIdx.x * 2 + 1] = 0.0;
    }
}

int launch_retinanet_assign(const Geometry& geo, const GtOffsets& go, const float* gt_boxes, const int64_t* gt_labels,
                            const LossScalars& sc, const LossWorkspace& ws, double* normalisers, const LossDebug& dbg,
                            cudaStream_t stream) {
    int rc = launch_assign_pass1(geo, go, gt_boxes, sc, ws, stream, /*with_class_sums=*/false);
    if (rc) return rc;
    const int tiles = geo.num_images * geo.tiles_per_image;
    double* tile_part = ws.block_part;       // positive_terms_kernel reuses the slots after the fold
    retina_labels_kernel<<<tiles, PAA_TILE, 0, stream>>>(geo, go, gt_boxes, gt_labels, ws.gtmax,
                                                         reinterpret_cast<const unsigned long long*>(ws.best), sc,
                                                         ws.matched, ws.paa_label, tile_part, dbg);
    PAA_LAUNCH_CHECK("retina_labels_kernel");
    PeerExchange none;
    memset(&none, 0, sizeof(none));
    return launch_fold_norm(tile_part, tiles, ws.local_norm, normalisers, none, stream);
}

}  // namespace paa
