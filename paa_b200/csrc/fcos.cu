// FCOS target assignment on sm_100a: rpn/fcos/loss.py:105-201 of the reference.
//
//   fcos_assign_kernel   one thread per (image, location), 128-location tiles; the image's GT boxes and areas are
//                        staged in shared memory 128 at a time and walked in index order, so "smallest area,
//                        first on ties" (torch.min over the GT axis, :190) is a strict `<`.  Leaves the label,
//                        the chosen GT and the tile's {positives, sum of centerness targets}.
//   atss_norm_kernel     fixed-order fold of the partials into the two normalisers, published to the other ranks
//                        exactly like PAA's (fcos/loss.py:251-266 reduces both over ranks).
// The regression targets are not materialised: the loss pass recomputes the four distances of the few positive
// locations from (location, GT).
#include "kernels.h"

namespace paa {

constexpr float kFcosInf = 100000000.0f;       // fcos/loss.py:19

__global__ void __launch_bounds__(PAA_TILE)
fcos_assign_kernel(const Geometry geo, const GtOffsets* __restrict__ gop, const float* __restrict__ gt_boxes,
                   const int64_t* __restrict__ gt_labels, const LossScalars sc, int* __restrict__ matched,
                   int* __restrict__ label_out, double* __restrict__ tile_part, const LossDebug dbg, bool ssc) {
    pdl_launch_dependents();
    const GtOffsets& go = *gop;
    __shared__ float4 s_box[PAA_TILE];
    __shared__ float s_area[PAA_TILE];
    __shared__ double s_part[PAA_TILE / PAA_WARP][2];
    const int n = blockIdx.x / geo.tiles_per_image;
    const int tile = blockIdx.x - n * geo.tiles_per_image;
    int first;
    const int l = tile_level(geo, tile, &first);
    const LevelView& lv = geo.lv[l];
    const int i = first + threadIdx.x;
    const bool valid = i < lv.n_anchor;
    const int gbase = go.v[n];
    const int G = go.v[n + 1] - gbase;
    // object_sizes_of_interest, fcos/loss.py:106-112
    const float lo = l == 0 ? -1.0f : (float)(32 << l), hi = l >= 4 ? kFcosInf : (float)(64 << l);
    const float radius = ssc ? 0.0f : sc.fcos_radius[l];
    const float margin = ssc ? 0.01f : 0.0f;                 // atss/loss.py:116 vs fcos/loss.py:179
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float x = 0.f, y = 0.f;
    if (valid) {
        const float4 p = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
        x = p.x;
        y = p.y;
        if (ssc) {                                           // anchor centres, atss/loss.py:97-98
            x = __fdiv_rn(__fadd_rn(p.z, p.x), 2.0f);
            y = __fdiv_rn(__fadd_rn(p.w, p.y), 2.0f);
        }
    }
    const float wx1 = warp_min(valid ? x : INFINITY), wx2 = warp_max(valid ? x : -INFINITY);
    const float wy1 = warp_min(valid ? y : INFINITY), wy2 = warp_max(valid ? y : -INFINITY);
    // get_sample_region's "no gt" test (fcos/loss.py:68-69): the first GT's centre x summed over all locations is
    // 0 exactly when that centre is 0 -- then no location is inside any box
    bool none_inside = false;
    if (radius > 0.0f) {
        const float4 b0 = ldg4(gt_boxes + (size_t)gbase * 4);
        none_inside = __fdiv_rn(__fadd_rn(b0.x, b0.z), 2.0f) == 0.0f;
    }
    float best_area = kFcosInf;
    int best = -1;
    for (int c0 = 0; c0 < G; c0 += PAA_TILE) {
        __syncthreads();
        if (c0 + threadIdx.x < G) {
            const float4 b = ldg4(gt_boxes + (size_t)(gbase + c0 + threadIdx.x) * 4);
            s_box[threadIdx.x] = b;
            s_area[threadIdx.x] = area_plus1(b);                       // BoxList.area(), bounding_box.py:226-231
        }
        __syncthreads();
        const int cnt = min(PAA_TILE, G - c0);
        // lane j tests GT g0 + j against the box around the warp's locations (a location inside a GT, or inside
        // its centre region, lies inside the GT); only the GTs that reach it are walked, in index order
        for (int g0 = 0; g0 < cnt; g0 += PAA_WARP) {
            bool hit = false;
            if (g0 + lane < cnt) {
                const float4 b = s_box[g0 + lane];
                hit = b.x < wx2 && b.z > wx1 && b.y < wy2 && b.w > wy1;
            }
            unsigned hits = __ballot_sync(PAA_FULL, hit);
            while (hits) {
                const int k = g0 + __ffs(hits) - 1;
                hits &= hits - 1;
                const float4 b = s_box[k];
                const float dl = __fsub_rn(x, b.x), dt = __fsub_rn(y, b.y), dr = __fsub_rn(b.z, x), db = __fsub_rn(b.w, y);
                bool inside;
                if (radius > 0.0f) {
                    // centre region of the GT clipped to the GT (fcos/loss.py:71-103)
                    const float cx = __fdiv_rn(__fadd_rn(b.x, b.z), 2.0f), cy = __fdiv_rn(__fadd_rn(b.y, b.w), 2.0f);
                    const float xmin = __fsub_rn(cx, radius), ymin = __fsub_rn(cy, radius);
                    const float xmax = __fadd_rn(cx, radius), ymax = __fadd_rn(cy, radius);
                    const float rx1 = xmin > b.x ? xmin : b.x, ry1 = ymin > b.y ? ymin : b.y;
                    const float rx2 = xmax > b.z ? b.z : xmax, ry2 = ymax > b.w ? b.w : ymax;
                    const float m = fminf(fminf(__fsub_rn(x, rx1), __fsub_rn(y, ry1)),
                                          fminf(__fsub_rn(rx2, x), __fsub_rn(ry2, y)));
                    inside = m > 0.0f && !none_inside;
                } else {
                    inside = fminf(fminf(dl, dt), fminf(dr, db)) > margin;                     // :178-179
                }
                const float mx = fmaxf(fmaxf(dl, dt), fmaxf(dr, db));
                const bool cared = mx >= lo && mx <= hi;                                         // :181-185
                const float area = s_area[k];
                if (valid && inside && cared && area < best_area) {                              // :187-193
                    best_area = area;
                    best = c0 + k;
                }
            }
        }
    }
    double npos = 0.0, sctr = 0.0;
    if (valid) {
        const size_t flat = (size_t)n * geo.A + lv.a_off + i;
        int label = 0;
        if (best >= 0) label = (int)gt_labels[gbase + best];
        matched[flat] = best < 0 ? 0 : best;
        label_out[flat] = label;
        if (dbg.matched_idx) dbg.matched_idx[flat] = best;
        if (dbg.iou_labels) dbg.iou_labels[flat] = label;
        if (label > 0) {
            const float4 g = ldg4(gt_boxes + (size_t)(gbase + best) * 4);
            npos = 1.0;
            if (ssc) {
                const AnchorFrame f = anchor_frame(ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4));
                sctr = (double)centerness_target(decode_box(encode_box(g, f), f), f);      // atss/loss.py:233-245
            } else {
                sctr = (double)fcos_centerness(fcos_ltrb(x, y, g, sc.fcos_norm != 0, sc.fcos_stride[l]));
            }
        }
    }
    npos = warp_sum(npos);
    sctr = warp_sum(sctr);
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

int launch_fcos_assign(const Geometry& geo, const float* gt_boxes, const int64_t* gt_labels,
                       const LossScalars& sc, const LossWorkspace& ws, double* normalisers, const PeerExchange& px,
                       const LossDebug& dbg, cudaStream_t stream, bool ssc) {
    const int tiles = geo.num_images * geo.tiles_per_image;
    double* tile_part = ws.block_part;       // positive_terms_kernel reuses the slots after the fold
    const GtOffsets* go = ws.go;
    fcos_assign_kernel<<<tiles, PAA_TILE, 0, stream>>>(geo, go, gt_boxes, gt_labels, sc, ws.matched, ws.paa_label,
                                                       tile_part, dbg, ssc);
    PAA_LAUNCH_CHECK("fcos_assign_kernel");
    return launch_fold_norm(tile_part, tiles, ws.local_norm, normalisers, px, stream);
}

}  // namespace paa
