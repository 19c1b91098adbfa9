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

// A block takes `tiles_per_block` consecutive 128-anchor tiles; the list of the image's low-quality GTs (best IoU
// below `high`: the only ones whose restore can change anything) is rebuilt only when the image changes.
constexpr int kMaxLabelBlocks = 8192;

__global__ void __launch_bounds__(PAA_TILE)
retina_labels_kernel(const Geometry geo, const GtOffsets* __restrict__ gop, const float* __restrict__ gt_boxes,
                     const int64_t* __restrict__ gt_labels, const unsigned* __restrict__ gtmax,
                     const unsigned long long* __restrict__ best, const LossScalars sc,
                     int* __restrict__ matched, int* __restrict__ label_out, unsigned* __restrict__ ignore_bits,
                     double* __restrict__ tile_part, const LossDebug dbg, int tiles_total, int tiles_per_block,
                     bool atss_iou) {
    pdl_launch_dependents();
    const GtOffsets& go = *gop;
    __shared__ float4 s_box[PAA_TILE];
    __shared__ float s_max[PAA_TILE];
    __shared__ int s_nlq;
    __shared__ double s_part[PAA_TILE / PAA_WARP][2];
    const float high = sc.iou_threshold, low = sc.bg_threshold;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double npos = 0.0, sctr = 0.0;
    int staged_image = -1, nlq = 0;
    bool overflow = false;        // more than PAA_TILE low-quality GTs: walk the whole GT list instead

    const int t_end = min((int)(blockIdx.x + 1) * tiles_per_block, tiles_total);
    for (int t = blockIdx.x * tiles_per_block; t < t_end; ++t) {
        const int n = t / geo.tiles_per_image;
        const int tile = t - n * geo.tiles_per_image;
        const int gbase = go.v[n];
        const int G = go.v[n + 1] - gbase;
        if (n != staged_image) {                               // block-uniform
            __syncthreads();
            if (threadIdx.x == 0) s_nlq = 0;
            __syncthreads();
            for (int g = threadIdx.x; g < G; g += PAA_TILE) {
                const float mx = __uint_as_float(gtmax[gbase + g]);
                if (mx < high) {
                    const int slot = atomicAdd(&s_nlq, 1);
                    if (slot < PAA_TILE) {
                        s_box[slot] = ldg4(gt_boxes + (size_t)(gbase + g) * 4);
                        s_max[slot] = mx;
                    }
                }
            }
            __syncthreads();
            nlq = s_nlq;
            overflow = nlq > PAA_TILE;
            staged_image = n;
        }
        int first;
        const int l = tile_level(geo, tile, &first);
        const LevelView& lv = geo.lv[l];
        const int i = first + threadIdx.x;
        const bool valid = i < lv.n_anchor;
        const size_t flat = (size_t)n * geo.A + lv.a_off + (valid ? i : 0);
        float bval = 0.0f;
        int bgt = 0;
        if (valid) unpack_best(best[flat], &bval, &bgt);
        int m = bgt;                                              // NaN: in neither set, keeps the argmax
        if (bval < low) m = -1;                                   // matcher.py:71   BELOW_LOW_THRESHOLD
        else if (bval >= low && bval < high) m = -2;              // matcher.py:72-74 BETWEEN_THRESHOLDS
        // A GT whose best IoU stays below `high` restores its best anchors (ties included) to their own argmax
        // (matcher.py:83-113); the best anchors of every other GT are matched already.
        const bool needs = valid && m < 0;
        if (nlq > 0 && __any_sync(PAA_FULL, needs)) {
            float4 a = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
            if (needs) a = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
            const float area_a = area_plus1(a);
            if (!overflow) {
                // lane k tests low-quality GT k against the box around the warp's anchors; only the GTs that reach
                // it (and the degenerate ones whose best IoU is 0, which every disjoint anchor "attains") are walked
                const float wx1 = warp_min(a.x), wy1 = warp_min(a.y), wx2 = warp_max(a.z), wy2 = warp_max(a.w);
                for (int k0 = 0; k0 < nlq; k0 += PAA_WARP) {
                    const int k = k0 + lane;
                    bool hit = false;
                    if (k < nlq) {
                        const float4 b = s_box[k];
                        const float w = __fadd_rn(__fsub_rn(fminf(b.z, wx2), fmaxf(b.x, wx1)), 1.0f);
                        const float h = __fadd_rn(__fsub_rn(fminf(b.w, wy2), fmaxf(b.y, wy1)), 1.0f);
                        hit = ((w > 0.0f) && (h > 0.0f)) || !(s_max[k] > 0.0f);
                    }
                    unsigned hits = __ballot_sync(PAA_FULL, hit);
                    while (hits) {
                        const int j = k0 + __ffs(hits) - 1;
                        hits &= hits - 1;
                        const float4 b = s_box[j];
                        if (needs && iou_plus1(b, area_plus1(b), a, area_a) == s_max[j]) m = bgt;
                    }
                }
            } else if (needs) {
                for (int g = 0; g < G; ++g) {
                    const float mx = __uint_as_float(gtmax[gbase + g]);
                    if (!(mx < high)) continue;
                    const float4 b = ldg4(gt_boxes + (size_t)(gbase + g) * 4);
                    if (iou_plus1(b, area_plus1(b), a, area_a) == mx) m = bgt;
                }
            }
        }
        if (!valid) continue;
        int label = 0;                                        // rpn/loss.py:68-69
        if (m >= 0) label = (int)gt_labels[gbase + m];        // generate_retinanet_labels, retinanet/loss.py:84-86
        else if (m == -2) label = -1;                         // rpn/loss.py:76-78
        if (atss_iou && label > 0) {
            // ATSS POSITIVE_TYPE 'IoU' (atss/loss.py:214-226): a positive whose centre is not inside its GT by more
            // than 0.01 is ignored; the others contribute their centerness target to the second normaliser
            const float4 a = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
            const float4 g = ldg4(gt_boxes + (size_t)(gbase + m) * 4);
            const AnchorFrame f = anchor_frame(a);
            const float in = fminf(fminf(__fsub_rn(f.cx, g.x), __fsub_rn(f.cy, g.y)),
                                   fminf(__fsub_rn(g.z, f.cx), __fsub_rn(g.w, f.cy)));
            if (in > 0.01f) sctr += (double)centerness_target(decode_box(encode_box(g, f), f), f);
            else label = -1;
        }
        matched[flat] = m < 0 ? 0 : m;                        // target[matched_idxs.clamp(min=0)], rpn/loss.py:52
        label_out[flat] = label;
        if (label < 0) {
            // one bit per ignored anchor for the loss pass, in the order of the level's NCHW planes:
            // [image][anchor slot][location] (loss.cu: BulkPlan::ign)
            const int loc = geo.apl == 1 ? i : i / geo.apl, slot = i - loc * geo.apl;
            const unsigned bit = (unsigned)((n * geo.apl + slot) * lv.hw + loc);
            atomicOr(ignore_bits + (size_t)geo.num_images * lv.tile_off * 4 + (bit >> 5), 1u << (bit & 31u));
        }
        if (dbg.matched_idx) dbg.matched_idx[flat] = m;
        if (dbg.iou_labels) dbg.iou_labels[flat] = label;
        npos += label > 0 ? 1.0 : 0.0;
    }
    npos = warp_sum(npos);
    sctr = warp_sum(sctr);
    __syncthreads();
    if (lane == 0) {
        s_part[warp][0] = npos;
        s_part[warp][1] = sctr;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < PAA_TILE / PAA_WARP; ++w) t += s_part[w][threadIdx.x];
        tile_part[(size_t)blockIdx.x * 2 + threadIdx.x] = t;
    }
}

int launch_retinanet_assign(const Geometry& geo, const float* gt_boxes, const int64_t* gt_labels,
                            const LossScalars& sc, const LossWorkspace& ws, double* normalisers, const LossDebug& dbg,
                            cudaStream_t stream, bool atss_iou, const PeerExchange* px) {
    int rc = launch_iou_match(geo, gt_boxes, sc, ws, stream);
    if (rc) return rc;
    const int tiles = geo.num_images * geo.tiles_per_image;
    const int tiles_per_block = (tiles + kMaxLabelBlocks - 1) / kMaxLabelBlocks;
    const int grid = (tiles + tiles_per_block - 1) / tiles_per_block;
    double* tile_part = ws.block_part;       // positive_terms_kernel reuses the slots after the fold
    // the ignore bitmap lives where PAA keeps its per-tile GT masks: 128 bits per tile = one per anchor
    unsigned* ignore_bits = reinterpret_cast<unsigned*>(ws.tile_gtmask);
    PAA_CUDA_CHECK(cudaMemsetAsync(ignore_bits, 0, sizeof(uint4) * (size_t)tiles, stream));
    const GtOffsets* go = ws.go;
    retina_labels_kernel<<<grid, PAA_TILE, 0, stream>>>(geo, go, gt_boxes, gt_labels, ws.gtmax,
                                                        reinterpret_cast<const unsigned long long*>(ws.best), sc,
                                                        ws.matched, ws.paa_label, ignore_bits, tile_part, dbg, tiles,
                                                        tiles_per_block, atss_iou);
    PAA_LAUNCH_CHECK("retina_labels_kernel");
    PeerExchange none;
    memset(&none, 0, sizeof(none));
    return launch_fold_norm(tile_part, grid, ws.local_norm, normalisers, px ? *px : none, stream);
}

}  // namespace paa
