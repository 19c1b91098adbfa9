// ATSS anchor assignment on sm_100a: the training-side neighbour of the PAA path.  It replaces
// ATSSLossComputation.prepare_targets with POSITIVE_TYPE 'ATSS' (paa_core/modeling/rpn/atss/loss.py:139-197)
// and leaves labels / matched GTs / normalisers in the same workspace arrays that paa_loss consumes, so the
// focal + GIoU + BCE pass (loss.cu) is shared with PAA (weights = centerness targets, loss.py:262-272).
//
//   atss_candidates_kernel  one block per GT, warp l = level l: the TOPK anchors of the level whose centres are
//                           nearest to the GT centre (float32 distance as torch computes it, ties to the smaller
//                           index); warp 0 then takes the L*TOPK candidates' IoUs, forms the GT's threshold
//                           mean + unbiased std, and offers every candidate that passes it and whose centre lies
//                           inside the GT to the anchor's (IoU, GT) key with atomicMax -- an anchor wanted by
//                           several GTs goes to the largest IoU, first GT on ties (loss.py:183-192).
//   atss_labels_kernel      one thread per anchor: label and matched GT from the key, centerness target of the
//                           positives, per-tile partial normalisers.
//   atss_norm_kernel        fixed-order fold of the partials into {num_pos, sum of centerness targets}, published
//                           to the other ranks exactly like the PAA normalisers.
#include "kernels.h"

namespace paa {

constexpr unsigned long long kNoKey64 = ~0ull;

// lane j < K holds the j-th smallest key so far (same scheme as the PAA candidate selection)
__device__ __forceinline__ void nearest_offer(unsigned long long& mine, unsigned long long key, bool is, int K, int lane) {
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

__global__ void __launch_bounds__(PAA_MAX_LEVELS * PAA_WARP)
atss_candidates_kernel(const Geometry geo, const GtOffsets* __restrict__ gop, const int* __restrict__ gt_image,
                       const float* __restrict__ gt_boxes, const int K,
                       unsigned long long* __restrict__ best, const LossDebug dbg) {
    pdl_launch_dependents();
    __shared__ unsigned s_cand[PAA_MAX_LEVELS][PAA_WARP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gi = blockIdx.x;
    if (gi >= __ldg(&gop->v[geo.num_images])) return;        // the grid covers the call's GT capacity
    const int n = __ldg(gt_image + gi);
    const int g_local = gi - __ldg(&gop->v[n]);
    const float4 gt = ldg4(gt_boxes + (size_t)gi * 4);
    const float gcx = __fdiv_rn(__fadd_rn(gt.z, gt.x), 2.0f), gcy = __fdiv_rn(__fadd_rn(gt.w, gt.y), 2.0f);
    {
        // nearest K anchor centres of level `warp` (loss.py:143-160): key = (distance bits, anchor index)
        const LevelView& lv = geo.lv[warp];
        const float* anc = lv.anchors + (size_t)n * geo.anchor_image_stride;
        unsigned long long mine = kNoKey64;
        auto offer = [&](int i) {
            unsigned long long key = kNoKey64;
            if (i >= 0) {
                const float4 a = ldg4(anc + (size_t)i * 4);
                const float dx = __fsub_rn(__fdiv_rn(__fadd_rn(a.z, a.x), 2.0f), gcx);
                const float dy = __fsub_rn(__fdiv_rn(__fadd_rn(a.w, a.y), 2.0f), gcy);
                const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
                key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(lv.a_off + i);
            }
            nearest_offer(mine, key, key != kNoKey64, K, lane);
        };
        const int W = lv.grid_w, H = W > 0 ? lv.hw / W : 0;
        if (W > 0 && geo.apl == 1 && K <= 9) {
            // Anchors of a level sit on a regular W x H lattice (anchor_generator.py:73-95): the 9 nearest centres
            // of any point lie inside the 3 x 3 block around its nearest lattice point (distance <= 2.13 cells),
            // every lattice point outside the 6 x 6 window around the point's cell is >= 3 cells away.  Only that
            // window is searched; distances are still computed from the anchor boxes themselves.
            const float4 a0 = ldg4(anc);
            const float cx0 = (a0.z + a0.x) * 0.5f, cy0 = (a0.w + a0.y) * 0.5f;
            float sx = 1.0f, sy = 1.0f;
            if (W > 1) {
                const float4 a1 = ldg4(anc + 4);
                sx = (a1.z + a1.x) * 0.5f - cx0;
            }
            if (H > 1) {
                const float4 ar = ldg4(anc + (size_t)W * 4);
                sy = (ar.w + ar.y) * 0.5f - cy0;
            }
            const int fx = (int)floorf((gcx - cx0) / sx), fy = (int)floorf((gcy - cy0) / sy);
            const int wx = min(6, W), wy = min(6, H);
            const int x0 = max(0, min(fx - 2, W - wx)), y0 = max(0, min(fy - 2, H - wy));
            for (int t0 = 0; t0 < wx * wy; t0 += PAA_WARP) {
                const int t = t0 + lane;
                int i = -1;
                if (t < wx * wy) i = (y0 + t / wx) * W + x0 + (t - (t / wx) * wx);
                offer(i);
            }
        } else {
            for (int i0 = 0; i0 < lv.n_anchor; i0 += PAA_WARP) offer(i0 + lane < lv.n_anchor ? i0 + lane : -1);
        }
        if (lane < K) s_cand[warp][lane] = (unsigned)(mine & 0xffffffffu);
    }
    __syncthreads();
    if (warp != 0) return;

    // warp 0: the L*K candidates, level-major, nearest first (the order of torch.cat(candidate_idxs), :161)
    const int total = geo.num_levels * K;
    const float area_g = area_plus1(gt);
    __shared__ float s_q[PAA_MAX_LEVELS * PAA_WARP];
    __shared__ float s_in[PAA_MAX_LEVELS * PAA_WARP];
    for (int c = lane; c < total; c += PAA_WARP) {
        const int aidx = (int)s_cand[c / K][c % K];
        const int l = anchor_level(geo, aidx);
        const float4 a = ldg4(geo.lv[l].anchors + (size_t)n * geo.anchor_image_stride + (size_t)(aidx - geo.lv[l].a_off) * 4);
        s_q[c] = iou_plus1(gt, area_g, a, area_plus1(a));                       // boxlist_iou(anchors, targets), :141
        const float acx = __fdiv_rn(__fadd_rn(a.z, a.x), 2.0f), acy = __fdiv_rn(__fadd_rn(a.w, a.y), 2.0f);
        s_in[c] = fminf(fminf(__fsub_rn(acx, gt.x), __fsub_rn(acy, gt.y)),
                        fminf(__fsub_rn(gt.z, acx), __fsub_rn(gt.w, acy)));      // :175-180
    }
    __syncwarp();
    double sum = 0.0;
    for (int c = 0; c < total; ++c) sum += (double)s_q[c];
    const double mean = sum / (double)total;
    double m2 = 0.0;
    for (int c = 0; c < total; ++c) {
        const double d = (double)s_q[c] - mean;
        m2 += d * d;
    }
    // iou_mean + iou_std (unbiased), both float32 tensors in the reference (:166-168)
    const float thr = __fadd_rn((float)mean, (float)sqrt(m2 / (double)(total - 1)));
    if (dbg.gmm && lane == 0) dbg.gmm[(size_t)gi * 8] = (double)thr;
    for (int c = lane; c < total; c += PAA_WARP) {
        const int aidx = (int)s_cand[c / K][c % K];
        if (dbg.cand_idx) dbg.cand_idx[(size_t)gi * total + c] = aidx;
        if (s_q[c] >= thr && s_in[c] > 0.01f)          // :169,180-181
            atomicMax(best + (size_t)n * geo.A + aidx, pack_best(s_q[c], g_local));
    }
    if (dbg.cand_cnt && lane == 0) dbg.cand_cnt[gi] = total;
}

__global__ void __launch_bounds__(PAA_TILE)
atss_labels_kernel(const Geometry geo, const GtOffsets* __restrict__ gop, const float* __restrict__ gt_boxes,
                   const int64_t* __restrict__ gt_labels, const unsigned long long* __restrict__ best,
                   int* __restrict__ matched, int* __restrict__ paa_label, double* __restrict__ tile_part,
                   const LossDebug dbg) {
    pdl_wait();
    pdl_launch_dependents();
    const GtOffsets& go = *gop;
    __shared__ double s_part[PAA_TILE / PAA_WARP][2];
    const int n = blockIdx.x / geo.tiles_per_image;
    const int tile = blockIdx.x - n * geo.tiles_per_image;
    int first;
    const int l = tile_level(geo, tile, &first);
    const LevelView& lv = geo.lv[l];
    const int i = first + threadIdx.x;
    double npos = 0.0, sctr = 0.0;
    if (i < lv.n_anchor) {
        const size_t flat = (size_t)n * geo.A + lv.a_off + i;
        const unsigned long long key = best[flat];
        int label = 0, m = 0;                    // anchors nobody claimed: background, argmax of all -INF is GT 0
        if (key != 0ull) {
            m = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
            label = (int)gt_labels[go.v[n] + m];
            if (label > 0) {
                const float4 a = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
                const AnchorFrame f = anchor_frame(a);
                const float4 gt = ldg4(gt_boxes + (size_t)(go.v[n] + m) * 4);
                const float4 tgt = decode_box(encode_box(gt, f), f);
                npos = 1.0;
                sctr = (double)centerness_target(tgt, f);
            }
        }
        matched[flat] = m;
        paa_label[flat] = label;
        if (dbg.matched_idx) dbg.matched_idx[flat] = key != 0ull ? m : -1;
        if (dbg.iou_labels) dbg.iou_labels[flat] = label;
    }
    npos = warp_sum(npos);
    sctr = warp_sum(sctr);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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

__global__ void __launch_bounds__(512)
atss_norm_kernel(const double* __restrict__ tile_part, int tiles, double* __restrict__ local_norm,
                 double* __restrict__ normalisers, const PeerExchange px) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ double s[16][2];
    double a[2] = {0.0, 0.0};
    for (int b = threadIdx.x; b < tiles; b += 512) {
        a[0] += tile_part[(size_t)b * 2];
        a[1] += tile_part[(size_t)b * 2 + 1];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a[0] = warp_sum(a[0]);
    a[1] = warp_sum(a[1]);
    if (lane == 0) {
        s[warp][0] = a[0];
        s[warp][1] = a[1];
    }
    __syncthreads();
    if (warp != 0) return;
    double cnt = 0.0, sum = 0.0;
    for (int w = 0; w < 16; ++w) {
        cnt += s[w][0];
        sum += s[w][1];
    }
    if (lane == 0) {
        local_norm[0] = cnt;
        local_norm[1] = sum;
        normalisers[0] = cnt;
        normalisers[1] = sum;
    }
    if (px.world > 1) {          // same protocol as select_gmm_kernel's publication
        double* own = px.buf[px.rank];
        unsigned long long epoch = 0;
        if (lane == 0) {
            epoch = (unsigned long long)own[kPeerEpochOffset] + 1ull;
            own[kPeerEpochOffset] = (double)epoch;
        }
        epoch = __shfl_sync(PAA_FULL, epoch, 0);
        if (lane < px.world) {
            volatile double* dst = px.buf[lane] + ((size_t)(epoch & 1ull) * kPeerMaxRanks + px.rank) * 4;
            dst[0] = cnt;
            dst[1] = sum;
            __threadfence_system();
            dst[2] = (double)epoch;
        }
    }
}

int launch_atss_assign(const Geometry& geo, const float* gt_boxes,
                       const int64_t* gt_labels, const LossScalars& sc, const LossWorkspace& ws, double* normalisers,
                       const PeerExchange& px, const LossDebug& dbg, cudaStream_t stream) {
    unsigned long long* best = reinterpret_cast<unsigned long long*>(ws.best);
    const GtOffsets* go = ws.go;
    atss_candidates_kernel<<<sc.gt_capacity, geo.num_levels * PAA_WARP, 0, stream>>>(geo, go, ws.gt_image, gt_boxes,
                                                                                   sc.topk, best, dbg);
    PAA_LAUNCH_CHECK("atss_candidates_kernel");
    const int tiles = geo.num_images * geo.tiles_per_image;
    // the per-tile partials live where positive_terms_kernel later puts its own (it runs after the fold)
    double* tile_part = ws.block_part;
    PAA_PDL_LAUNCH(atss_labels_kernel, tiles, PAA_TILE, stream, geo, go, gt_boxes, gt_labels, best, ws.matched,
                   ws.paa_label, tile_part, dbg);
    return launch_fold_norm(tile_part, tiles, ws.local_norm, normalisers, px, stream);
}

int launch_fold_norm(const double* tile_part, int tiles, double* local_norm, double* normalisers,
                     const PeerExchange& px, cudaStream_t stream) {
    PAA_PDL_LAUNCH(atss_norm_kernel, 1, 512, stream, tile_part, tiles, local_norm, normalisers, px);
    return 0;
}

}  // namespace paa
