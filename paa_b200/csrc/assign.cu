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
// K1: IoU matching -- every anchor's best GT (first maximum) and every GT's maximal IoU
// (boxlist_iou + the two reductions of Matcher, boxlist_ops.py:81-116 / matcher.py:66,92), without the [G, A]
// matrix.  Touches no head tensor.
//
// Every WARP works on its own: 32 anchors (an 8 x 4 patch of a fine level's grid where the grid width is known --
// a compact footprint intersects ~35 % fewer GT boxes than 32 neighbours in a row -- else 32 consecutive anchors),
// the image's GT boxes 32 at a time, one per lane, straight from global memory (L1-resident after the first
// warp).  A lane whose GT intersects the warp's bounding box raises a ballot bit (non-intersecting pairs have IoU
// exactly +0 and can neither raise a maximum nor win the first-maximum rule against the initial (0, GT 0)); the
// round's boxes are parked in the warp's own slice of shared memory and the hits evaluated from there (branch-free,
// bit-exact IoU; broadcast reads).  Per-GT maxima go to global memory as one RED.MAX per (warp, hit GT).  No
// block barrier: round 1's block-staged version spent its time waiting (30 us alone on the C2 batch).
// For crowded images the coarse levels, whose anchors overlap every GT, are additionally cut into `parts` ranges
// of the GT list whose per-anchor results meet in an atomicMax on packed (IoU, GT) keys.
// ---------------------------------------------------------------------------------------------
constexpr int kPassThreads = 256;

struct Pass1Plan {
    unsigned iou_blocks;
    // tiles below `light_tiles` (fine levels) go two to a block over all GTs of the image; the tiles of the coarse
    // levels are additionally cut into `parts` ranges of the GT list
    int light_tiles, light_pairs, heavy_pairs, parts;
    int patches;                                 // 1: fine levels are matched in 8 x 4 patches per warp
    unsigned patch_off[PAA_MAX_LEVELS + 1];      // first 32 x 8 region of each fine level (per image)
    int patch_rx[PAA_MAX_LEVELS];                // regions per grid row
    int patch_h[PAA_MAX_LEVELS];                 // grid height of the level
    unsigned patch_magic[PAA_MAX_LEVELS];        // floor(2^16 / patch_rx) + 1: r / patch_rx == (r * magic) >> 16 while r * patch_rx < 2^16
    // L2 prefetch of the classification logits for match_score_kernel (PAA only; 0 pieces = off): image-major pieces
    // of kPfPiece bytes, `pf_level_off[l]` = first piece of level l inside an image
    unsigned pf_pieces, pf_per_image;
    unsigned pf_level_off[PAA_MAX_LEVELS + 1];
};
constexpr unsigned kPfPiece = 16384;

// order-preserving map float -> unsigned (and back), so that floats compare / reduce as integers
__device__ __forceinline__ unsigned ordered_bits(float v) {
    unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// The GT boxes of one round as the warp staged them (one per lane) and their areas.
struct WarpGts {
    float4 box[PAA_WARP];
    float area[PAA_WARP];
};

template <int kWide>
__device__ __forceinline__ void iou_hits(unsigned m, int c0, int lane, bool valid, const float4& a, float area_a,
                                         const WarpGts& gts, float& best_v, int& best_g,
                                         unsigned* __restrict__ gtmax_image) {
    while (m) {
        int js[kWide];
        float qv[kWide];
#pragma unroll
        for (int u = 0; u < kWide; ++u) {
            js[u] = __ffs(m) - 1;          // -1 once the hits are used up
            m &= m - 1u;
        }
#pragma unroll
        for (int u = 0; u < kWide; ++u) {
            const int j = js[u] >= 0 ? js[u] : 0;
            const float v = iou_plus1_flat(gts.box[j], gts.area[j], a, area_a);      // broadcast reads
            qv[u] = (js[u] >= 0 && valid) ? v : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < kWide; ++u) {          // ascending GT order: the first-maximum rule of torch.max
            if (qv[u] > best_v) {
                best_v = qv[u];
                best_g = c0 + js[u];
            }
        }
        unsigned wm[kWide];
#pragma unroll
        for (int u = 0; u < kWide; ++u) wm[u] = __reduce_max_sync(PAA_FULL, __float_as_uint(qv[u]));
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < kWide; ++u)
                if (js[u] >= 0 && wm[u] != 0u) atomicMax(gtmax_image + c0 + js[u], wm[u]);
        }
    }
}

__global__ void __launch_bounds__(kPassThreads)
iou_match_kernel(const Geometry geo, const GtOffsets* __restrict__ gop, const Pass1Plan plan,
                 const float* __restrict__ gt_boxes, unsigned* __restrict__ gtmax,
                 unsigned long long* __restrict__ best) {
    PAA_TRACE_SCOPE(1);
    pdl_launch_dependents();
    // This kernel is bound by instruction issue and moves half a megabyte: the memory system idles under it.  The next
    // kernel (match_score_kernel) waits on the latency of its class-row loads, so the logits it will read first --
    // image by image, in its own block order -- are pulled towards the L2 from here (bulk prefetch, one instruction
    // per 16 KB piece, issued by one thread per block before the dependency wait: the logits do not depend on
    // anything this step computes).
    if (plan.pf_pieces && threadIdx.x == 0) {
        const unsigned nblk = gridDim.x * gridDim.y;
        for (unsigned p = blockIdx.y * gridDim.x + blockIdx.x; p < plan.pf_pieces; p += nblk) {
            const unsigned img = p / plan.pf_per_image, r = p - img * plan.pf_per_image;
            int l = 0;
#pragma unroll 1
            for (int k = 1; k < geo.num_levels; ++k)
                if (r >= plan.pf_level_off[k]) l = k;
            const unsigned long long level_bytes = (unsigned long long)geo.lv[l].n_anchor * geo.C * 4ull;
            const unsigned long long first = (unsigned long long)(r - plan.pf_level_off[l]) * kPfPiece;
            const unsigned long long left = level_bytes - first;
            const unsigned bytes = (unsigned)(left < kPfPiece ? left : kPfPiece) & ~15u;
            const char* src = reinterpret_cast<const char*>(geo.lv[l].cls) + (unsigned long long)img * level_bytes + first;
            if (bytes && (reinterpret_cast<uintptr_t>(src) & 15u) == 0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
        }
    }
    pdl_wait();                       // the GT ranges and the cleared maxima come from prep_step_kernel
    PAA_TRACE_WAITED();
    __shared__ WarpGts s_gts[kPassThreads / PAA_WARP];
    const GtOffsets& go = *gop;
    // grid: x = work item inside an image (the coarse levels first: their anchors intersect every GT), y = image in
    // the order of decreasing GT count (the cost of a block grows with the GT count of its image) -- blocks are
    // dispatched x-fastest, so the heaviest work of the heaviest image starts first.  No division by a run-time
    // value on this path: the kernel is bound by instruction issue and every warp pays for the prologue.
    const unsigned q = blockIdx.x;
    const int n = go.by_load[blockIdx.y];
    const unsigned heavy_items = (unsigned)plan.heavy_pairs * (unsigned)plan.parts;       // per image
    const bool heavy = q < heavy_items;
    int tile0 = 0, tile_end = 0, part = 0, parts = 1;
    bool split = false;
    if (heavy) {
        parts = plan.parts;
        const unsigned pair = parts > 1 ? q / (unsigned)parts : q;
        part = (int)(q - pair * (unsigned)parts);
        tile0 = plan.light_tiles + 2 * (plan.heavy_pairs - 1 - (int)pair);
        tile_end = geo.tiles_per_image;
        split = parts > 1;
    } else {
        tile0 = 2 * (plan.light_pairs - 1 - (int)(q - heavy_items));
        tile_end = plan.light_tiles;
    }
    int l = 0, i = 0;
    bool valid = false;
    if (!heavy && plan.patches) {
        // region index inside the image, coarsest fine level first
        const unsigned per_image = plan.patch_off[PAA_MAX_LEVELS];
        const unsigned reg = per_image - 1u - (q - heavy_items);
#pragma unroll
        for (int k = 1; k < PAA_MAX_LEVELS; ++k)
            if (k < geo.num_levels && plan.patch_off[k] <= reg && plan.patch_off[k] < per_image) l = k;
        unsigned off_l = plan.patch_off[0], magic = plan.patch_magic[0];
        int prx = plan.patch_rx[0], H = plan.patch_h[0];
#pragma unroll
        for (int k = 1; k < PAA_MAX_LEVELS; ++k)       // selects, not indexed reads: the plan stays in the constant bank
            if (k == l) {
                off_l = plan.patch_off[k];
                prx = plan.patch_rx[k];
                magic = plan.patch_magic[k];
                H = plan.patch_h[k];
            }
        const unsigned rr = reg - off_l;
        const int W = geo.lv[l].grid_w;
        const int ry = (int)((rr * magic) >> 16), rx = (int)rr - ry * prx;       // rr / prx
        const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
        const int col = rx * 32 + 8 * (w & 3) + (ln & 7), row = ry * 8 + 4 * (w >> 2) + (ln >> 3);
        valid = col < W && row < H;
        i = row * W + col;
    } else {
        const int tile = tile0 + (threadIdx.x >> 7);
        const bool tile_ok = tile < tile_end;
        int first = 0;
        l = tile_ok ? tile_level(geo, tile, &first) : 0;
        i = first + (threadIdx.x & (PAA_TILE - 1));
        valid = tile_ok && i < geo.lv[l].n_anchor;
    }
    const LevelView& lv = geo.lv[l];
    const int lane = threadIdx.x & 31;
    if (!__any_sync(PAA_FULL, valid)) return;          // warp-uniform: nothing of this warp lies inside the level

    float4 a = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
    if (valid) a = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
    const float area_a = area_plus1(a);
    // the warp's bounding box: four integer reductions on order-preserving keys
    const float wx1 = from_ordered_bits(__reduce_min_sync(PAA_FULL, ordered_bits(a.x)));
    const float wy1 = from_ordered_bits(__reduce_min_sync(PAA_FULL, ordered_bits(a.y)));
    const float wx2 = from_ordered_bits(__reduce_max_sync(PAA_FULL, ordered_bits(a.z)));
    const float wy2 = from_ordered_bits(__reduce_max_sync(PAA_FULL, ordered_bits(a.w)));

    const int gbase = go.v[n];
    const int G_all = go.v[n + 1] - gbase;
    const int g_lo = (int)(((long long)G_all * part) / parts), G = (int)(((long long)G_all * (part + 1)) / parts);
    float best_v = 0.0f;
    int best_g = 0;
    const bool crowded = G_all > 128;
    unsigned* gtmax_image = gtmax + gbase;

    for (int c0 = g_lo; c0 < G; c0 += PAA_WARP) {
        const int g = c0 + lane;
        float4 b = make_float4(0.f, 0.f, -1.f, -1.f);
        bool hit = false;
        if (g < G) {
            b = ldg4(gt_boxes + (size_t)(gbase + g) * 4);
            const float w = __fadd_rn(__fsub_rn(fminf(b.z, wx2), fmaxf(b.x, wx1)), 1.0f);
            const float h = __fadd_rn(__fsub_rn(fminf(b.w, wy2), fmaxf(b.y, wy1)), 1.0f);
            hit = (w > 0.0f) && (h > 0.0f);
        }
        const unsigned m = __ballot_sync(PAA_FULL, hit);
        if (m == 0u) continue;                                   // warp-uniform
        WarpGts& gts = s_gts[threadIdx.x >> 5];
        __syncwarp();                                            // the previous round's reads are done
        gts.box[lane] = b;
        gts.area[lane] = area_plus1(b);
        __syncwarp();
        // the kernel is bound by instruction issue: a fine level's warp (a handful of hits per round) takes its
        // hits one at a time (no idle slots); the coarse levels, whose anchors overlap every GT, and crowded images
        // take four at a time as independent chains -- their walk over all the GTs is the kernel's longest path
        if (crowded || heavy) iou_hits<4>(m, c0, lane, valid, a, area_a, gts, best_v, best_g, gtmax_image);
        else iou_hits<1>(m, c0, lane, valid, a, area_a, gts, best_v, best_g, gtmax_image);
    }
    if (valid) {
        unsigned long long* dst = best + (size_t)n * geo.A + lv.a_off + i;
        if (!split) *dst = pack_best(best_v, best_g);
        else if (best_v > 0.0f) atomicMax(dst, pack_best(best_v, best_g));     // pre-zeroed by prep_step_kernel
    }
}

// Levels from this one on are "coarse": few, large anchors that overlap (nearly) every GT of the image.
int first_heavy_level(const Geometry& geo) {
    int l = geo.num_levels;
    while (l > 0 && geo.lv[l - 1].n_anchor <= 2048) --l;
    return l;
}

// Number of GT-list parts for the coarse tiles, whose anchors overlap every GT of the image: their walk over the GT
// list is the kernel's longest dependent path, so it is cut into ranges of about 32 GTs (at most 8 parts) whose
// per-anchor results meet in an atomicMax -- one part, i.e. no split and no atomics, up to 32 GTs per image.
int gt_parts_for(int max_gt_per_image) {
    int p = (max_gt_per_image + 31) / 32;
    return p < 1 ? 1 : (p > 8 ? 8 : p);
}

// ---------------------------------------------------------------------------------------------
// K0: first launch of every assign call.  All blocks clear the workspace's zeroed prefix (per-GT maxima, flags,
// ticket, pool counters) and, where asked, the best-GT keys that are merged by atomicMax; block 0 leaves the
// per-image GT ranges, the image order by load and every GT's image in device memory, where all later kernels
// of the step read them.  Nothing downstream depends on HOST values of the GT counts any more, so a CUDA graph
// captured around the step can be replayed on another batch (PaaLossArgs::gt_offsets_dev).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prep_step_kernel(const GtOffsets host_go, const int* __restrict__ dev_offsets, int num_images, int capacity,
                 GtOffsets* __restrict__ dst, int* __restrict__ gt_image, uint4* __restrict__ zero_base,
                 size_t zero_vec, uint2* __restrict__ best, int A, int best_a0, int best_n) {
    PAA_TRACE_SCOPE(0);
    pdl_launch_dependents();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < zero_vec; i += nthr) zero_base[i] = make_uint4(0u, 0u, 0u, 0u);
    if (best_n > 0) {
        const size_t total = (size_t)num_images * best_n;
        for (size_t i = tid; i < total; i += nthr) {
            const size_t n = i / best_n, a = i - n * best_n;
            best[n * A + best_a0 + a] = make_uint2(0u, 0u);
        }
    }
    if (blockIdx.x != 0) return;
    __shared__ int s_v[PAA_MAX_IMAGES + 1];
    for (int i = threadIdx.x; i <= num_images; i += blockDim.x) {
        int v = dev_offsets ? __ldg(dev_offsets + i) : host_go.v[i];
        s_v[i] = v < 0 ? 0 : (v > capacity ? capacity : v);          // never index past what the workspace holds
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= num_images; i += blockDim.x) dst->v[i] = s_v[i];
    // images with many GT boxes cost proportionally more in the IoU kernel: most GTs first, ties by index
    for (int i = threadIdx.x; i < num_images; i += blockDim.x) {
        const int gi = s_v[i + 1] - s_v[i];
        int rank = 0;
        for (int j = 0; j < num_images; ++j) {
            const int gj = s_v[j + 1] - s_v[j];
            rank += (gj > gi || (gj == gi && j < i)) ? 1 : 0;
        }
        dst->by_load[rank] = (unsigned char)i;
    }
    for (int n = 0; n < num_images; ++n)
        for (int g = s_v[n] + threadIdx.x; g < s_v[n + 1]; g += blockDim.x) gt_image[g] = n;
}

#ifdef PAA_TRACE
PAA_TRACE_SETTER(trace_set_assign)
#endif

int launch_prep_step(const Geometry& geo, const GtOffsets& host_go, const int* dev_offsets, const LossScalars& sc,
                     const LossWorkspace& ws, void* zero_base, bool clear_heavy_best, bool clear_all_best,
                     cudaStream_t stream) {
    const size_t zero_vec = ws.zero_bytes / sizeof(uint4);
    int a0 = 0, n_best = 0;
    if (clear_all_best) {
        n_best = geo.A;
    } else if (clear_heavy_best) {
        const int hl = first_heavy_level(geo);
        if (hl < geo.num_levels && sc.gt_parts > 1) {
            a0 = geo.lv[hl].a_off;
            n_best = geo.A - a0;
        }
    }
    const size_t work = zero_vec + (size_t)geo.num_images * n_best;
    int grid = (int)((work + 1023) / 1024);
    grid = grid < 1 ? 1 : (grid > 148 * 4 ? 148 * 4 : grid);
    KernelTimer timer(PAA_KERNEL_PREP_STEP, stream);
    prep_step_kernel<<<grid, 256, 0, stream>>>(host_go, dev_offsets, geo.num_images, sc.gt_capacity, ws.go, ws.gt_image,
                                               reinterpret_cast<uint4*>(zero_base), zero_vec, ws.best, geo.A, a0,
                                               n_best);
    PAA_LAUNCH_CHECK("prep_step_kernel");
    return 0;
}

int launch_iou_match(const Geometry& geo, const float* gt_boxes, const LossScalars& sc, const LossWorkspace& ws,
                     cudaStream_t stream) {
    Pass1Plan plan;
    const int heavy_level = first_heavy_level(geo);
    plan.light_tiles = heavy_level < geo.num_levels ? geo.lv[heavy_level].tile_off : geo.tiles_per_image;
    plan.light_pairs = (plan.light_tiles + 1) / 2;
    plan.heavy_pairs = (geo.tiles_per_image - plan.light_tiles + 1) / 2;
    plan.parts = sc.gt_parts;
    // fine levels in 2-D patches when their grid widths are known
    plan.patches = geo.apl == 1 ? 1 : 0;
    unsigned regions = 0;
    for (int l = 0; l < PAA_MAX_LEVELS; ++l) {
        plan.patch_off[l] = regions;
        plan.patch_rx[l] = 1;
        plan.patch_h[l] = 1;
        plan.patch_magic[l] = 65537u;
        if (l >= heavy_level) continue;
        const LevelView& lv = geo.lv[l];
        if (lv.grid_w <= 0) {
            plan.patches = 0;
            continue;
        }
        plan.patch_rx[l] = (lv.grid_w + 31) / 32;
        plan.patch_h[l] = lv.hw / lv.grid_w;
        plan.patch_magic[l] = 65536u / (unsigned)plan.patch_rx[l] + 1u;
        const unsigned level_regions = (unsigned)plan.patch_rx[l] * (unsigned)((plan.patch_h[l] + 7) / 8);
        if ((unsigned long long)level_regions * (unsigned)plan.patch_rx[l] >= 65536ull) plan.patches = 0;   // beyond the magic division's range (r * d < 2^16)
        regions += level_regions;
    }
    plan.patch_off[PAA_MAX_LEVELS] = regions;
    if (regions == 0) plan.patches = 0;
    const unsigned light_items = plan.patches ? regions : (unsigned)plan.light_pairs;
    plan.iou_blocks = light_items + (unsigned)(plan.heavy_pairs * plan.parts);          // per image
    // PAA: the logits of the first images of the call towards the L2 while this kernel runs (see the kernel)
    plan.pf_pieces = 0;
    plan.pf_per_image = 1;
    for (int l = 0; l <= PAA_MAX_LEVELS; ++l) plan.pf_level_off[l] = 0;
    if (sc.flavour == PAA_LOSS_PAA) {
        unsigned per_image = 0;
        for (int l = 0; l < PAA_MAX_LEVELS; ++l) {
            plan.pf_level_off[l] = per_image;
            if (l < geo.num_levels)
                per_image += (unsigned)(((unsigned long long)geo.lv[l].n_anchor * geo.C * 4ull + kPfPiece - 1) / kPfPiece);
        }
        plan.pf_level_off[PAA_MAX_LEVELS] = per_image;
        plan.pf_per_image = per_image ? per_image : 1;
        // Measured (C2, tools/step_trace.py, profiles/r2_match_prefetch.txt): with 2 images per rank (14 MB of logits)
        // match_score_kernel 13.8 -> 11.3 us and the step 73.7 -> 72.4 us; with 16 images any budget from 16 to 128 MB
        // gives back in this kernel's issue slots (13.1 -> 14-16.6 us) what it saves in the next (36.2 -> 34.3 us).
        int budget_mb = 16;
        if (const char* e = getenv("PAA_MATCH_PREFETCH_MB")) budget_mb = atoi(e);      // measurement switch
        const unsigned long long want = (unsigned long long)budget_mb * (1ull << 20) / kPfPiece;
        const unsigned long long all = (unsigned long long)per_image * (unsigned)geo.num_images;
        plan.pf_pieces = (unsigned)(want < all ? want : all);
    }
    const GtOffsets* gop = ws.go;
    unsigned long long* best = reinterpret_cast<unsigned long long*>(ws.best);
    KernelTimer timer(PAA_KERNEL_PASS1, stream);
    PAA_PDL_LAUNCH(iou_match_kernel, dim3(plan.iou_blocks, (unsigned)geo.num_images), kPassThreads, stream, geo, gop,
                   plan, gt_boxes, ws.gtmax, best);
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
// The class sum -- the only pass over the logits that the assignment needs -- is taken HERE, and only where it
// is used: for the tile's IoU-positive anchors (45 % of the anchors of the C2 batch, clustered around the ground
// truth: a tile without any skips the pass, a float4 column without any its loads; measured DRAM traffic 92 MB
// instead of the 115 MB of all logits -- the fetch granularity, not the 32-byte sector, decides what a skipped
// column saves).  Every class enters as a negative, sum_c p_c^gamma * softplus(x_c); the labelled class's term is
// then swapped for the positive one.  A block is one tile of 128 consecutive anchors of one level: as 32 float4
// columns x 4 class groups for the sums (a warp reads up to 512 contiguous bytes of one class row, four rows in
// flight per thread), the four partial sums of a column meet in shared memory in a fixed order, then one thread
// per anchor again.  Measured alternatives (C2 batch, stand-alone event time, this kernel 41 us): the pass as a
// kernel of its own in front of this one 35 + 16 us (it waits on memory latency, not on bytes: as long as for ALL
// columns); eight rows in flight per thread 45 us (registers cost occupancy); the tile's rows staged in shared memory
// by bulk asynchronous copies behind an mbarrier, issued ahead of the Matcher, 56 us (40 KB per block leave five
// blocks per SM for the latency-bound Matcher); bulk L2 prefetches of the rows ahead of the Matcher 45 us.
// ---------------------------------------------------------------------------------------------
constexpr int kSumGroups = PAA_TILE / PAA_WARP;       // class groups = warps of the block

// One negative-class focal term without its (1-alpha) factor: p^gamma * softplus(x).
__device__ __forceinline__ float neg_term_only(float x, float gamma, bool g2) {
    const SigmoidLean sl = sigmoid_lean(x);
    return focal_pow(sl.p, gamma, g2) * sl.sp;
}

__device__ __forceinline__ void add_terms(float4& acc, float4 x, float gamma, bool g2) {
    acc.x += neg_term_only(x.x, gamma, g2);
    acc.y += neg_term_only(x.y, gamma, g2);
    acc.z += neg_term_only(x.z, gamma, g2);
    acc.w += neg_term_only(x.w, gamma, g2);
}

// kNhwc: the channels-last instantiation.  An anchor's C logits are contiguous there, a warp's 32 anchors one
// contiguous run of 32*C floats: the class sums are a flat, fully coalesced stream (512 bytes per warp instruction,
// four in flight per lane) whose per-float4 partial sums meet per anchor in the warp's slice of shared memory -- no
// block barrier, no strided class rows.
constexpr int kFlatQ = 20;                            // float4 per anchor the flat path handles (C <= 80, C % 4 == 0)

template <bool kNhwc>
__global__ void __launch_bounds__(PAA_TILE)
match_score_kernel(const Geometry geo, const GtOffsets* __restrict__ gop, const float* __restrict__ gt_boxes,
                   const int64_t* __restrict__ gt_labels, const unsigned* __restrict__ gtmax,
                   const unsigned long long* __restrict__ best, const LossScalars sc,
                   int* __restrict__ matched, float* __restrict__ score, int* __restrict__ paa_label, uint4* __restrict__ tile_gtmask,
                   int* __restrict__ seg_count, unsigned long long* __restrict__ seg_pool,
                   const float* __restrict__ teacher_score, const LossDebug dbg) {
    PAA_TRACE_SCOPE(2);
    pdl_wait();
    PAA_TRACE_WAITED();
    pdl_launch_dependents();
    const GtOffsets& go = *gop;
    __shared__ int s_lq[PAA_TILE];
    __shared__ int s_nlq;
    __shared__ unsigned s_mask[4];
    __shared__ unsigned s_need[PAA_TILE / PAA_WARP];
    __shared__ __align__(16) float s_part[kNhwc ? 1 : kSumGroups][PAA_TILE];
    __shared__ float s_flat[kNhwc ? PAA_TILE * (kFlatQ + 1) : 1];
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
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    float bval = 0.0f;
    int bgt = 0;
    const size_t flat = (size_t)n * geo.A + lv.a_off + (valid ? i : 0);
    if (valid) {
        a = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
        unpack_best(best[flat], &bval, &bgt);
    }
    const float area_a = area_plus1(a);
    int m = (bval >= thr) ? bgt : -1;
    if (!(bval >= thr) && !(bval < thr)) m = bgt;          // NaN: neither below nor restored -> keeps argmax

    const int C = geo.C;
    // The kernel is a chain of memory round trips (ten per block before this reordering), so everything whose
    // address is known early is requested early.  The matched GT can only be the best GT (`m` ends up as bgt or -1):
    // its label, its box and this anchor's regression outputs are asked for now -- the label as a load that completes
    // under the low-quality pass, the others as L2 prefetches (no registers held across the class sums).
    int label_of_best = 0;
    if (valid) {
        label_of_best = (int)__ldg(gt_labels + gbase + bgt);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(gt_boxes + (size_t)(gbase + bgt) * 4));
        const float* rp0 = lv.reg + head_offset(geo, lv, n, i, 0, 4);
        const size_t cs0 = head_cstride(geo, lv);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp0));
        if (cs0 > 8) {                                     // NCHW: the four channels live in four planes
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rp0 + cs0));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rp0 + 2 * cs0));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rp0 + 3 * cs0));
        }
    }

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
                if (q == __uint_as_float(gtmax[gbase + gg])) m = bgt;
            }
        }
        __syncthreads();
    }
    if (!valid) m = -1;
    const int label = m >= 0 ? label_of_best : 0;           // m is bgt or -1
    const bool positive = m >= 0 && label > 0;
    // the candidate-pool slot is reserved here (its atomic returns under the class sums), the key is stored at the end
    int pool_slot = -1;
    const int seg = positive ? (gbase + m) * geo.num_levels + l : -1;
    {
        // consecutive anchors mostly share their GT: one atomic per (warp, segment) instead of one per anchor
        const unsigned peers = __match_any_sync(PAA_FULL, seg);
        const int leader = __ffs(peers) - 1;
        int first = 0;
        if (positive && lane == leader) first = atomicAdd(&seg_count[seg], __popc(peers));
        first = __shfl_sync(PAA_FULL, first, leader);
        if (positive) pool_slot = first + __popc(peers & ((1u << lane) - 1u));
        if (positive) asm volatile("prefetch.global.L2 [%0];" ::"l"(lv.cls + head_offset(geo, lv, n, i, label - 1, C)));
    }
    // which GTs (index mod 128) have matched anchors in this tile: lets the per-GT selection kernel
    // skip every tile that cannot contain one of its anchors, with no false negatives
    if (m >= 0) atomicOr(&s_mask[(m & 127) >> 5], 1u << (m & 31));
    const unsigned pos_mask = __ballot_sync(PAA_FULL, positive);
    if (lane == 0) s_need[warp] = pos_mask;
    __syncthreads();
    if (threadIdx.x == 0)
        tile_gtmask[(size_t)n * geo.tiles_per_image + tile] = make_uint4(s_mask[0], s_mask[1], s_mask[2], s_mask[3]);
    if (valid) {
        matched[flat] = m;
        paa_label[flat] = 0;
        if (dbg.matched_idx) dbg.matched_idx[flat] = m;
        if (dbg.iou_labels) dbg.iou_labels[flat] = label;
    }
    const bool any_positive = (s_need[0] | s_need[1] | s_need[2] | s_need[3]) != 0u;       // block-uniform
    if (!any_positive) {
        if (valid && dbg.combined_loss) dbg.combined_loss[flat] = 1.0e8f;
        return;
    }

    // ---- class sums of the tile's IoU-positive anchors --------------------------------------------
    const bool g2 = (sc.gamma == 2.0f);
    const bool vec = !kNhwc && geo.apl == 1 && (lv.hw & 3) == 0 && (reinterpret_cast<uintptr_t>(lv.cls) & 15u) == 0;
    const bool flat_rows = kNhwc && (C & 3) == 0 && C <= 4 * kFlatQ && (reinterpret_cast<uintptr_t>(lv.cls) & 15u) == 0;
    float negsum = 0.0f;
    if (flat_rows) {
        const int q = C >> 2;                                             // float4 per anchor
        const unsigned magic = (65536u + (unsigned)q - 1u) / (unsigned)q; // f / q for f < 32 * q <= 640
        float* wbuf = s_flat + warp * PAA_WARP * (kFlatQ + 1);
        if (pos_mask != 0u) {                                             // warp-uniform
            const float4* p = reinterpret_cast<const float4*>(
                lv.cls + ((size_t)n * lv.n_anchor + first + PAA_WARP * warp) * C);
            for (int k0 = 0; k0 < q; k0 += 4) {
                float4 x[4];
                int slot[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned f = (unsigned)lane + 32u * (unsigned)(k0 + j);
                    const unsigned an = (f * magic) >> 16;                 // anchor of the warp this float4 belongs to
                    slot[j] = -1;
                    if (k0 + j < q && ((pos_mask >> an) & 1u)) {
                        x[j] = __ldg(p + f);
                        slot[j] = (int)(an * (kFlatQ + 1) + (f - an * (unsigned)q));
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (slot[j] >= 0) {
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        add_terms(acc, x[j], sc.gamma, g2);
                        wbuf[slot[j]] = (acc.x + acc.y) + (acc.z + acc.w);
                    }
            }
            __syncwarp();
            if (positive) {
                const float* row = wbuf + lane * (kFlatQ + 1);
                for (int t = 0; t < q; ++t) negsum += row[t];
            }
        }
    } else if (vec) {
        const int col = lane, grp = warp;                                 // anchors 4*col .. 4*col+3 of the tile
        const unsigned need = (s_need[col >> 3] >> ((col & 7) * 4)) & 0xfu;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (need != 0u) {
            const int cg = (C + kSumGroups - 1) / kSumGroups;
            const int c_begin = grp * cg, c_end = min(C, c_begin + cg);
            const unsigned st = (unsigned)lv.hw >> 2;                     // float4 per class row
            const float4* p = reinterpret_cast<const float4*>(lv.cls + (size_t)n * C * lv.hw + first + 4 * col);
            int c0 = c_begin;
            for (; c0 + 4 <= c_end; c0 += 4) {
                float4 x[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = __ldg(p + (size_t)(c0 + j) * st);
#pragma unroll
                for (int j = 0; j < 4; ++j) add_terms(acc, x[j], sc.gamma, g2);
            }
            for (; c0 < c_end; ++c0) add_terms(acc, __ldg(p + (size_t)c0 * st), sc.gamma, g2);
        }
        *reinterpret_cast<float4*>(&s_part[grp][4 * col]) = acc;
        __syncthreads();
        // group 0's sum first, then groups 1, 2, 3: the order round 1 fixed (bit-identical scores)
        negsum = s_part[0][threadIdx.x];
#pragma unroll
        for (int g = 1; g < kSumGroups; ++g) negsum += s_part[g][threadIdx.x];
    } else if (positive) {
        const float* p = lv.cls + head_offset(geo, lv, n, i, 0, C);
        const size_t st = head_cstride(geo, lv);
        int c0 = 0;
        for (; c0 + 8 <= C; c0 += 8) {
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = __ldg(p + (size_t)(c0 + j) * st);
#pragma unroll
            for (int j = 0; j < 8; ++j) negsum += neg_term_only(x[j], sc.gamma, g2);
        }
        for (; c0 < C; ++c0) negsum += neg_term_only(__ldg(p + (size_t)c0 * st), sc.gamma, g2);
    }

    float s = 1.0e8f;   // loss.py:15,301-306 filler, only visible through the debug output
    if (positive) {
        // sum_c focal(x_c | label): every class as a negative, then the labelled class's negative term swapped
        // for the positive one (accurate path for that single term)
        const float xl = __ldg(lv.cls + head_offset(geo, lv, n, i, label - 1, C));
        float tp, gp;
        focal_positive(xl, sigmoid_parts(xl), sc.gamma, g2, sc.alpha, &tp, &gp);
        const float fsum = fmaf(1.0f - sc.alpha, negsum - neg_term_only(xl, sc.gamma, g2), tp);
        const float* rp = lv.reg + head_offset(geo, lv, n, i, 0, 4);
        const float4 d = load_channels4(rp, head_cstride(geo, lv));
        const AnchorFrame f = anchor_frame(a);
        const float4 pred = decode_box(d, f);
        const float4 gt = ldg4(gt_boxes + (size_t)(gbase + m) * 4);
        const float4 tgt = decode_box(encode_box(gt, f), f);
        s = __fadd_rn(fsum, giou_loss_boxes(pred, tgt));
        score[flat] = s;
        // hand the anchor to its (GT, level) candidate pool: the per-GT selection then reads one short
        // contiguous list instead of searching the image for its anchors
        const float key_score = teacher_score ? teacher_score[flat] : s;
        if (pool_slot < sc.seg_cap)
            seg_pool[(size_t)seg * kSegCap + pool_slot] =
                ((unsigned long long)ordered_bits(key_score) << 32) | (unsigned)(lv.a_off + i);
    }
    if (valid && dbg.combined_loss) dbg.combined_loss[flat] = s;
}

int launch_match_score(const Geometry& geo, const float* gt_boxes,
                       const int64_t* gt_labels, const LossScalars& sc, const LossWorkspace& ws,
                       const float* teacher_score, const LossDebug& dbg, cudaStream_t stream) {
    int grid = geo.num_images * geo.tiles_per_image;
    KernelTimer timer(PAA_KERNEL_MATCH_SCORE, stream);
    const GtOffsets* gop = ws.go;
    if (geo.nhwc)
        PAA_PDL_LAUNCH(match_score_kernel<true>, grid, PAA_TILE, stream, geo, gop, gt_boxes, gt_labels, ws.gtmax,
                       reinterpret_cast<const unsigned long long*>(ws.best), sc, ws.matched, ws.score, ws.paa_label,
                       ws.tile_gtmask, ws.seg_count, ws.seg_pool, teacher_score, dbg);
    else
        PAA_PDL_LAUNCH(match_score_kernel<false>, grid, PAA_TILE, stream, geo, gop, gt_boxes, gt_labels, ws.gtmax,
                       reinterpret_cast<const unsigned long long*>(ws.best), sc, ws.matched, ws.score, ws.paa_label,
                       ws.tile_gtmask, ws.seg_count, ws.seg_pool, teacher_score, dbg);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K3: one block per GT -- warp l collects the top-k candidates of level l, then warp 0 sorts them,
// fits the two-component GMM, labels the positive prefix and leaves the GT's partial normalisers.
// loss.py:151-236 + sklearn GaussianMixture.fit/predict/score_samples.
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long kEmptyKey = ~0ull;

// sklearn/utils/_array_api.py:1338-1366 for two entries.
__device__ __forceinline__ double logsumexp2(double a0, double a1) {
    if (a0 == a1) return (log1p(0.0) + 0.6931471805599453094) + a0;   // m = 2, masked sum = 0
    const double hi = fmax(a0, a1), lo = fmin(a0, a1);
    return (log1p(exp(lo - hi)) + 0.0) + hi;                          // m = 1, log(m) = 0
}

// Mixture parameters as every lane holds them between iterations: mpc = mu * pc, lw = log(w),
// ld = float32(log(pc)).
struct GmmState {
    double mu0, mu1, mpc0, mpc1, lw0, lw1;
    float pc0, pc1, ld0, ld1;
};

// a_k = log N(x | mu_k, pc_k) + log w_k with the dtype flow of _estimate_log_gaussian_prob
// (sklearn/mixture/_gaussian_mixture.py:490-553): float32 product x*pc and float32 log-det once the
// precisions have been re-estimated, float32 buffer for the squared distance, float64 elsewhere.
// The initial float64 precisions (1.0) need no special case: x * 1.0f and lp + 0.0f are exact.
__device__ __forceinline__ void weighted_log_prob(float x, const GmmState& s, double* a0, double* a1) {
    const float LOG2PI = 1.8378770664093453f;
    const double y0 = (double)__fmul_rn(x, s.pc0) - s.mpc0;
    const double y1 = (double)__fmul_rn(x, s.pc1) - s.mpc1;
    const float q0 = __double2float_rn(y0 * y0);
    const float q1 = __double2float_rn(y1 * y1);
    const float lp0 = __fmul_rn(-0.5f, __fadd_rn(LOG2PI, q0));
    const float lp1 = __fmul_rn(-0.5f, __fadd_rn(LOG2PI, q1));
    *a0 = (double)__fadd_rn(lp0, s.ld0) + s.lw0;
    *a1 = (double)__fadd_rn(lp1, s.ld1) + s.lw1;
}

__device__ __forceinline__ double warp_prod(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(PAA_FULL, v, o);
    return v;
}

// Sums seven per-lane quantities over the warp with 9 + 7 64-bit shuffles instead of 35: each of the
// first three butterfly rounds halves the number of quantities a lane still carries (it keeps the half
// selected by its lane bit and hands the other half to its partner), two more rounds finish the one
// quantity left, and the totals are broadcast from the lanes that own them (quantity q: lanes 4q..4q+3).
// v[7] is a padding slot.
__device__ __forceinline__ void warp_sum7(double (&v)[8], int lane) {
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
    double w[4], x[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double send = b4 ? v[j] : v[4 + j], keep = b4 ? v[4 + j] : v[j];
        w[j] = keep + __shfl_xor_sync(PAA_FULL, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const double send = b3 ? w[j] : w[2 + j], keep = b3 ? w[2 + j] : w[j];
        x[j] = keep + __shfl_xor_sync(PAA_FULL, send, 8);
    }
    const double send = b2 ? x[0] : x[1], keep = b2 ? x[1] : x[0];
    double y = keep + __shfl_xor_sync(PAA_FULL, send, 4);
    y += __shfl_xor_sync(PAA_FULL, y, 2);
    y += __shfl_xor_sync(PAA_FULL, y, 1);
#pragma unroll
    for (int q = 0; q < 7; ++q) v[q] = __shfl_sync(PAA_FULL, y, 4 * q);
}

// warp_sum7 with the warp PRODUCT of `P` riding in the padding slot (v[7] = P on entry; the slot is multiplied where
// the others are added): the log-likelihood's prod(1 + s) costs no butterfly of its own.  Returns the product.
__device__ __forceinline__ double warp_sum7_prod(double (&v)[8], int lane) {
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
    double w[4], x[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double send = b4 ? v[j] : v[4 + j], keep = b4 ? v[4 + j] : v[j];
        const double recv = __shfl_xor_sync(PAA_FULL, send, 16);
        w[j] = (j == 3 && b4) ? keep * recv : keep + recv;               // slot 7 lives in the b4 half's w[3]
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const double send = b3 ? w[j] : w[2 + j], keep = b3 ? w[2 + j] : w[j];
        const double recv = __shfl_xor_sync(PAA_FULL, send, 8);
        x[j] = (j == 1 && b4 && b3) ? keep * recv : keep + recv;
    }
    const double send = b2 ? x[0] : x[1], keep = b2 ? x[1] : x[0];
    const bool prod_lane = b4 && b3 && b2;                               // lanes 28..31 own slot 7
    double recv = __shfl_xor_sync(PAA_FULL, send, 4);
    double y = prod_lane ? keep * recv : keep + recv;
    recv = __shfl_xor_sync(PAA_FULL, y, 2);
    y = prod_lane ? y * recv : y + recv;
    recv = __shfl_xor_sync(PAA_FULL, y, 1);
    y = prod_lane ? y * recv : y + recv;
#pragma unroll
    for (int q = 0; q < 7; ++q) v[q] = __shfl_sync(PAA_FULL, y, 4 * q);
    return __shfl_sync(PAA_FULL, y, 28);
}

// The seven sums through shared memory (`red`: kGmmRedDoubles doubles owned by the fitting warp; the other warps of a
// select_gmm block have exited by then): every lane parks its seven values, lane (q = lane & 7, part = lane >> 3) adds
// entries 8 part .. 8 part + 7 of quantity q as a three-level tree, two shuffle rounds join the four parts, and the
// totals travel back through seven words that every lane reads.  Two shared-memory round trips + two shuffle rounds on
// the dependent path instead of five shuffle rounds + a broadcast round of seven more; the product of slot 7 is a
// five-round butterfly of its own, an independent chain hidden behind them.  Measured (tools/gmm_sections.py, C2):
// 1534 -> 1395 cycles per iteration (E-step 378 -> 342, reduction 341 -> 329, M-step 277 -> 179: the totals no longer
// arrive one shuffle at a time), select_gmm_kernel 46.8 -> 42.0 us in the C2 step; every fit of C2 / C3 / C5 still stops
// at scikit-learn's iteration.  Row stride 33 doubles: the half-warp's rows x parts fall on different bank pairs.
constexpr int kGmmRedStride = 33;
constexpr int kGmmRedDoubles = 8 * kGmmRedStride + 8;
__device__ __forceinline__ double warp_sum7_prod_smem(double (&v)[8], int lane, double* red) {
#pragma unroll
    for (int q = 0; q < 7; ++q) red[q * kGmmRedStride + lane] = v[q];
    __syncwarp();
    // the product of slot 7 as a butterfly of its own: an independent chain that the sums' two memory round trips hide
    double P = v[7];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) P *= __shfl_xor_sync(PAA_FULL, P, o);
    const int q = (lane & 7) < 7 ? (lane & 7) : 0, part = lane >> 3;         // (lanes 7, 15, 23, 31 repeat quantity 0)
    const double* row = red + q * kGmmRedStride + part * 8;
    double e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] = row[j];
#pragma unroll
    for (int w = 4; w > 0; w >>= 1)
#pragma unroll
        for (int j = 0; j < w; ++j) e[j] = e[2 * j] + e[2 * j + 1];
    double y = e[0];
    y += __shfl_xor_sync(PAA_FULL, y, 8);
    y += __shfl_xor_sync(PAA_FULL, y, 16);
    double* tot = red + 8 * kGmmRedStride;
    if (lane < 7) tot[lane] = y;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 7; ++k) v[k] = tot[k];
    return P;
}

// Fits the mixture on the warp's n sorted samples (lane holds x[lane + 32*k]) and returns the length
// of the positive prefix (loss.py:206-217).  out8 (nullable, lane 0 writes) receives the parameters.
//
// The fit is a serial chain of float64 operations with nothing to overlap it with (measured on B200:
// DFMA 8 cycles, 64-bit shuffle 24, IEEE float32 sqrt / div ~60 each), so an iteration is organised for
// the shortest chain and the fewest issued instructions:
//  * E-step, one sample per lane: s = exp(lo - hi), responsibilities 1/(1+s) and s/(1+s) (one exp and one
//    reciprocal instead of sklearn's three exps; they differ from exp(a_k - lse) by a few 1e-16, far
//    below what the float32 roundings of the next E-step can see).
//  * The mean log-likelihood that drives the stopping rule is sum(hi) + log(prod(1 + s)): one logarithm
//    per iteration instead of one log1p per sample (the product of <= 128 factors in [1, 2] is exact to
//    ~1e-14 relative, the rule compares against 1e-3).
//  * One reduction round: moments are taken about the PREVIOUS means (p0, p1), from which the new mean and
//    the variance about the new mean follow exactly: with d = mu - p = (A - p eps) / nk,
//        mu = p + d,   sum r (x - mu)^2 = B - 2 d A + d^2 S.
//    (sklearn evaluates sum r (x - mu)^2 directly; the two agree to ~1e-14 relative, far below the
//    float32 rounding of the variance that follows.)
//  * The iteration's five logarithms are one call: log var_k in lanes 0-1, log w_k in lanes 2-3,
//    log prod in lane 4.  float32(log pc_k) is recovered from log var_k, which is known 120 cycles before
//    pc_k = 1 / sqrt(var_k) has gone through its two IEEE float32 roundings:
//        log pc = -log(var)/2 + log1p(e)/2,   e = pc^2 var - 1  (|e| <= 2^-22, one exact fma).
//  * The loop is rotated: the next E-step's log-probabilities are issued before the branch on the
//    stopping rule, so the rule's own chain (log prod -> mean -> difference) is off the critical path; on
//    exit they are exactly the final E-step's.
template <int SPL>
__device__ int gmm_positive_prefix(const float (&x)[SPL], int n, int lane, double* out8, double* red) {
    GmmState s;
    s.lw0 = s.lw1 = -0.6931471805599453094;                              // log 0.5
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
    s.mpc0 = s.mu0;
    s.mpc1 = s.mu1;
    s.pc0 = s.pc1 = 1.0f;
    s.ld0 = s.ld1 = 0.0f;
    double w0 = 0.5, w1 = 0.5;      // last M-step's weights / variances, for the debug output only
    float var0 = 1.0f, var1 = 1.0f;
    double lower = -INFINITY;
    int n_iter = 0;
    bool converged = false;
    const double EPS10 = 10.0 * 2.220446049250313e-16;
    const double inv_n = 1.0 / (double)n;
    double a0[SPL], a1[SPL];
#ifdef PAA_PROFILE_GMM
    long long prof_sec[5] = {0, 0, 0, 0, 0}, prof_last = 0;
#endif
    for (;;) {
#pragma unroll
        for (int k = 0; k < SPL; ++k) weighted_log_prob(x[k], s, &a0[k], &a1[k]);
        if (converged || n_iter == 100) break;
        ++n_iter;
#ifdef PAA_PROFILE_GMM
        const long long tA = clock64();
        if (n_iter > 1) prof_sec[4] += tA - prof_last;
#endif
        const double p0 = s.mu0, p1 = s.mu1;
        double v[8];          // S0 S1 A0 A1 B0 B1 H pad
        double P = 1.0;
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            if (k > 0 && n <= 32 * k) break;                  // warp-uniform
            const bool first_hi = a0[k] >= a1[k];
            const double hi = first_hi ? a0[k] : a1[k], lo = first_hi ? a1[k] : a0[k];
            const double sx = exp_nonpos_clamped(lo - hi);
            const double u = 1.0 + sx;
            const double rh = rcp_fast(u), rl = sx * rh;
            const bool valid = lane + 32 * k < n;
            const double r0 = valid ? (first_hi ? rh : rl) : 0.0;
            const double r1 = valid ? (first_hi ? rl : rh) : 0.0;
            const double e0 = (double)x[k] - p0, e1 = (double)x[k] - p1;
            const double t0 = r0 * e0, t1 = r1 * e1;
            if (k == 0) {
                v[0] = r0;
                v[1] = r1;
                v[2] = t0;
                v[3] = t1;
                v[4] = t0 * e0;
                v[5] = t1 * e1;
                v[6] = valid ? hi : 0.0;
                P = valid ? u : 1.0;
            } else {
                v[0] += r0;
                v[1] += r1;
                v[2] += t0;
                v[3] += t1;
                v[4] = fma(t0, e0, v[4]);
                v[5] = fma(t1, e1, v[5]);
                v[6] += valid ? hi : 0.0;
                P *= valid ? u : 1.0;
            }
        }
        v[7] = P;
#ifdef PAA_PROFILE_GMM
        const long long tB = clock64();
        prof_sec[0] += tB - tA;
#endif
        P = red ? warp_sum7_prod_smem(v, lane, red) : warp_sum7_prod(v, lane);      // (shuffle-only form: no scratch space)
#ifdef PAA_PROFILE_GMM
        const long long tC = clock64();
        prof_sec[1] += tC - tB;
#endif
        // M-step, both components in every lane (two independent chains fill each other's latency)
        const double nk0 = v[0] + EPS10, nk1 = v[1] + EPS10;
        const double rn0 = rcp_fast(nk0), rn1 = rcp_fast(nk1), rs = rcp_fast(nk0 + nk1);
        const double d0 = fma(-p0, EPS10, v[2]) * rn0, d1 = fma(-p1, EPS10, v[3]) * rn1;
        const double c0 = fma(d0, fma(d0, v[0], -2.0 * v[2]), v[4]);
        const double c1 = fma(d1, fma(d1, v[1], -2.0 * v[3]), v[5]);
        var0 = __fadd_rn(__double2float_rn(fmax(c0, 0.0) * rn0), 1e-6f);
        var1 = __fadd_rn(__double2float_rn(fmax(c1, 0.0) * rn1), 1e-6f);
        w0 = nk0 * rs;
        w1 = nk1 * rs;
#ifdef PAA_PROFILE_GMM
        const long long tD = clock64();
        prof_sec[2] += tD - tC;
#endif
        // precisions first: the next E-step needs them before anything else, and as branch-free code (no range
        // checks, no slow-path calls: the variances are normal numbers >= 1e-6) the two components' square roots and
        // divisions interleave with each other and with the logarithm below instead of running one after the other
        s.pc0 = div_rn_normal(1.0f, sqrt_rn_normal(var0));
        s.pc1 = div_rn_normal(1.0f, sqrt_rn_normal(var1));
        const double dv0 = (double)var0, dv1 = (double)var1;
        const double a_var = (lane & 1) ? dv1 : dv0, a_w = (lane & 1) ? w1 : w0;     // selects, not branches
        const double arg = lane < 2 ? a_var : (lane < 4 ? a_w : P);
        const double lg = log_pos(arg);
        s.mu0 = p0 + d0;
        s.mu1 = p1 + d1;
        const double pd0 = (double)s.pc0, pd1 = (double)s.pc1;
        s.mpc0 = s.mu0 * pd0;
        s.mpc1 = s.mu1 * pd1;
        const double e0 = fma(pd0 * pd0, (double)var0, -1.0), e1 = fma(pd1 * pd1, (double)var1, -1.0);
        const double h0 = (0.5 * e0) * fma(-0.5, e0, 1.0), h1 = (0.5 * e1) * fma(-0.5, e1, 1.0);
        s.ld0 = __double2float_rn(fma(-0.5, __shfl_sync(PAA_FULL, lg, 0), h0));
        s.ld1 = __double2float_rn(fma(-0.5, __shfl_sync(PAA_FULL, lg, 1), h1));
        s.lw0 = __shfl_sync(PAA_FULL, lg, 2);
        s.lw1 = __shfl_sync(PAA_FULL, lg, 3);
        const double lb = (v[6] + __shfl_sync(PAA_FULL, lg, 4)) * inv_n;
        converged = fabs(lb - lower) < 1e-3;
        lower = lb;
#ifdef PAA_PROFILE_GMM
        prof_last = clock64();
        prof_sec[3] += prof_last - tD;
#endif
    }
    // a0 / a1 are the final E-step: predict (argmax, ties -> 0) and score_samples
    double best_score = -INFINITY;
    int best_idx = 0x7fffffff;
    bool any_fg = false;
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        if (lane + 32 * k < n) {
            if (!(a1[k] > a0[k])) {           // component 0 == foreground (loss.py:206)
                const double sc = logsumexp2(a0[k], a1[k]);
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
#ifdef PAA_PROFILE_GMM
    if (out8 && lane == 0) {      // sections: E-step, reductions, M-step to var, log/sqrt to loop end, next E-step's log-probs
        out8[2] = (double)prof_sec[0];
        out8[3] = (double)prof_sec[1];
        out8[4] = (double)prof_sec[2];
        out8[5] = (double)prof_sec[3];
        out8[7] = (double)prof_sec[4];
        out8[6] = (double)n_iter;
        return any_fg ? best_idx + 1 : n;
    }
#endif
    if (out8 && lane == 0) {
        out8[0] = w0;
        out8[1] = w1;
        out8[2] = s.mu0;
        out8[3] = s.mu1;
        out8[4] = (double)var0;
        out8[5] = (double)var1;
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

// Starts the list from one batch of up to 32 keys (one per lane, kEmptyKey where there is none): every
// lane ranks its key against the other 31 with independent shuffles (no serial insertion), the keys
// are scattered by rank through `scratch` (32 entries of this warp).  Valid keys are distinct.
__device__ __forceinline__ unsigned long long topk_first_batch(unsigned long long key, int K, int lane,
                                                               unsigned long long* scratch) {
    int rank = 0;
#pragma unroll
    for (int j = 0; j < PAA_WARP; ++j) rank += (__shfl_sync(PAA_FULL, key, j) < key) ? 1 : 0;
    scratch[lane] = kEmptyKey;
    __syncwarp();
    if (key != kEmptyKey) scratch[rank] = key;
    __syncwarp();
    const unsigned long long mine = (lane < K) ? scratch[lane] : kEmptyKey;
    __syncwarp();
    return mine;
}

// MODE 0: the whole per-GT program in one block (warp l selects level l, warp 0 fits): the latency-optimal form, what
//         runs for calls sized for up to kSplitAboveGts GTs.
// MODE 1 / MODE 2: the same program cut in two launches for calls sized for more GTs than that (crowded or large
//         batches: throughput, not latency) -- MODE 1 the selection alone (block per GT, sorted candidates to memory),
//         MODE 2 the fit alone with ONE WARP per block and GT: up to 24 resident fits per SM instead of the 6 that MODE
//         0's five-warp blocks allow.  Bit-identical results (tests/test_gpu_loss.py).  Measured (B200, graph replay):
//         C3 (16 000 GTs) 0.7005 against 0.7290 ms per step, C5 (3327 GTs) 0.4036 against 0.4085; C2 (861 GTs) gets
//         slower (149.1 against 146.8 us: one more launch on the latency path) -- hence the threshold.  With four GTs
//         per block the fit launch was no faster than MODE 0 (a block lives as long as its slowest fit).
//         What the measurements say about the fit:
//         * tools/fp64_rate_probe.cu: a dependent DFMA takes 8 cycles, a sub-partition's FP64 pipe accepts a warp-wide
//           DFMA every 2 cycles (64 FMA/clk/SM).  A lone fit -- the tail of every step -- is a latency chain with idle
//           issue slots (its transcendental pieces are Estrin-evaluated already, fastmath64.cuh); from four resident
//           fits per sub-partition on, the pipe is the limit: 6 fits x ~185 FP64 instructions x 2 cycles = 2220 cycles
//           per iteration, which is what MODE 2 measures on C3 (2290; 1480 for a fit alone; tools/gmm_profile.py,
//           profiles/r2_gmm_c3_fused_vs_split.txt).
//         * Even so MODE 2's launch takes 142 us on C3 where 3552 resident fits of 11.3 us each would need 51 us: on
//           average a third of the warp slots hold a fit in its pipe-bound phase; the rest of a block's life (the
//           dependent loads in front of the fit, the fence + ticket atomic behind it) was not broken down further.
//         * Letting the block index pick the fitting warp of MODE 0, in case warp 0 of every block lands on one
//           sub-partition, changed nothing (profiles/r2_gmm_split_trace.txt).
constexpr int kSelectFused = 0, kSelectOnly = 1, kFitOnly = 2;
constexpr int kSplitAboveGts = 4096;       // calls sized for more GTs than this take the two-launch form (see above)
constexpr int kFitWarps = 1;

template <int SPL, int MODE>
// (64 registers: with the usual five levels six blocks fit an SM, so ~900 GTs are one wave)
// (four samples per lane -- TOPK 20 -- need more than 64 registers: no spills at two 256-thread blocks per SM)
__global__ void __launch_bounds__(MODE == kFitOnly ? kFitWarps * PAA_WARP : PAA_MAX_LEVELS * PAA_WARP,
                                  MODE == kFitOnly ? (SPL >= 4 ? 12 : 24) : (SPL >= 4 ? 2 : 4))
select_gmm_kernel(const Geometry geo, const GtOffsets* __restrict__ gop, const int* __restrict__ gt_image,
                  const float* __restrict__ gt_boxes, const int64_t* __restrict__ gt_labels,
                  const LossScalars sc, const uint4* __restrict__ tile_gtmask,
                  const int* __restrict__ matched, const float* __restrict__ score,
                  const int* __restrict__ seg_count, const unsigned long long* __restrict__ seg_pool,
                  int* __restrict__ paa_label, int* __restrict__ pos_list,
                  int* __restrict__ part_npos, double* __restrict__ part_siou,
                  unsigned* __restrict__ ticket, double* __restrict__ local_norm,
                  double* __restrict__ normalisers, const PeerExchange px, const LossDebug dbg,
                  unsigned long long* __restrict__ cand_sorted, int* __restrict__ cand_n) {
    // lets the loss pass (a programmatic dependent launch) become resident and prefetch while the slowest fits run
    PAA_TRACE_SCOPE(MODE == kFitOnly ? 6 : 3);
    pdl_wait();
    PAA_TRACE_WAITED();
    pdl_launch_dependents();
    __shared__ unsigned long long s_level[PAA_MAX_LEVELS][PAA_WARP];   // per-level top-K, ascending
    __shared__ int s_cnt[PAA_MAX_LEVELS];
    __shared__ unsigned long long s_key[PAA_MAX_CANDIDATES];
    __shared__ unsigned long long s_sorted[PAA_MAX_CANDIDATES];
    __shared__ double s_red[MODE == kFitOnly ? kFitWarps : 1][kGmmRedDoubles];      // the fitting warp's reduction space

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gi = MODE == kFitOnly ? blockIdx.x * kFitWarps + warp : blockIdx.x;
    const int K = sc.topk;
    const int cap = geo.num_levels * K;
    // the grid covers the call's GT capacity; the step's own count lives in device memory
    const int num_gt_total = __ldg(&gop->v[geo.num_images]);
    if (gi >= num_gt_total) return;
    // one more fit lives on this SM (a hint for bulk_focal_early_kernel, which keeps off the SMs where a fit runs:
    // a fit is one dependent instruction chain and every other ready warp on its sub-partition stretches it)
    unsigned smid = 0;
    if (MODE == kSelectFused) {
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        if (threadIdx.x == 0) atomicAdd(ticket + kCtlSmLive + (smid & (kCtlMaxSms - 1)), 1u);
    }
    const int n = __ldg(gt_image + gi);              // image of this GT
    const int g_local = gi - __ldg(&gop->v[n]);
    const int cls_label = (int)gt_labels[gi];

#ifdef PAA_PROFILE_GMM
    const long long prof_t0 = clock64();
#endif
    int n_cand = 0;
    if (MODE != kFitOnly) {
    // ---- phase 1: warp l = level l (loss.py:160-172) ------------------------------------------------
    {
        const int l = warp;
        unsigned long long mine = kEmptyKey;   // lane j < K holds the j-th smallest key so far
        if (cls_label > 0) {                   // loss.py:166 requires a positive IoU label
            const LevelView& lv = geo.lv[l];
            const int seg = gi * geo.num_levels + l;
            // the pool's first 128 slots are fetched together with its fill count (slots past the count hold
            // stale keys and are masked once the count is known): one memory round trip instead of two
            const unsigned long long* pool = seg_pool + (size_t)seg * kSegCap;
            const int seg_n = __ldg(seg_count + seg);
            unsigned long long k4[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) k4[r] = __ldg(pool + r * PAA_WARP + lane);
            if (seg_n <= sc.seg_cap) {
                // normal case: the anchors matched to this (GT, level) were pooled by match_score_kernel
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (r * PAA_WARP + lane >= seg_n) k4[r] = kEmptyKey;
                if (seg_n > 0) mine = topk_first_batch(k4[0], K, lane, s_level[l]);
#pragma unroll
                for (int r = 1; r < 4; ++r)
                    if (r * PAA_WARP < seg_n) topk_offer(mine, k4[r], k4[r] != kEmptyKey, K, lane);
                for (int j0 = 4 * PAA_WARP; j0 < seg_n; j0 += 4 * PAA_WARP) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int j = j0 + r * PAA_WARP + lane;
                        k4[r] = (j < seg_n) ? __ldg(pool + j) : kEmptyKey;
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (j0 + r * PAA_WARP < seg_n) topk_offer(mine, k4[r], k4[r] != kEmptyKey, K, lane);
                }
            } else {
                // overflowed pool: search the tiles of this level that contain anchors of this GT
                const unsigned gbit = 1u << (g_local & 31);
                const int gword = (g_local & 127) >> 5;
                const int* mrow = matched + (size_t)n * geo.A;
                const float* srow = score + (size_t)n * geo.A;
                const int t_end = (l + 1 < geo.num_levels) ? geo.lv[l + 1].tile_off : geo.tiles_per_image;
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
                            const bool is = (mv[r] == g_local);
                            unsigned long long key = kEmptyKey;
                            if (is) key = ((unsigned long long)ordered_bits(sv[r]) << 32) | (unsigned)(lv.a_off + i);
                            topk_offer(mine, key, is, K, lane);
                        }
                    }
                }
            }
        }
        const unsigned have = __ballot_sync(PAA_FULL, lane < K && mine != kEmptyKey);
        if (lane < K) s_level[l][lane] = mine;
        if (lane == 0) s_cnt[l] = __popc(have);
    }
    __syncthreads();
    // (Letting the block index pick the warp that fits -- in case "always warp 0" puts the fits of all resident blocks
    // on one sub-partition's FP64 pipe -- was measured and is slower, profiles/r2_gmm_split_trace.txt.)
    if (warp != 0) return;

    // ---- phase 2 (warp 0): concatenate the levels, sort, fit, label --------------------------------
    for (int l = 0; l < geo.num_levels; ++l) {
        const int cnt = s_cnt[l];
        if (lane < cnt) s_key[n_cand + lane] = s_level[l][lane];
        n_cand += cnt;
    }
    __syncwarp();
    // sort all candidates by (loss, index): rank by counting, n_cand <= cap <= 128
    for (int j = lane; j < n_cand; j += PAA_WARP) {
        const unsigned long long kj = s_key[j];
        int rank = 0;
        for (int q = 0; q < n_cand; ++q) rank += (s_key[q] < kj) ? 1 : 0;
        s_sorted[rank] = kj;
        if (MODE == kSelectOnly) cand_sorted[(size_t)gi * PAA_MAX_CANDIDATES + rank] = kj;
    }
    __syncwarp();
    if (MODE == kSelectOnly) {
        if (lane == 0) cand_n[gi] = n_cand;
        return;
    }
    } else {
        n_cand = __ldg(cand_n + gi);
    }

#ifdef PAA_PROFILE_GMM
    const long long prof_t1 = clock64();
#endif
    // this lane's candidates (sorted order), with the anchors and regression outputs the IoU normaliser
    // will need: the loads are issued before the fit and consumed after it
    int aidx[SPL];
    float x[SPL];
    float4 pa[SPL], pd[SPL];
    const float4 gt = ldg4(gt_boxes + (size_t)gi * 4);
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        const int j = lane + 32 * k;
        aidx[k] = -1;
        x[k] = 0.f;
        pa[k] = make_float4(0.f, 0.f, 1.f, 1.f);
        pd[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < n_cand) {
            const unsigned long long key = MODE == kFitOnly ? __ldg(cand_sorted + (size_t)gi * PAA_MAX_CANDIDATES + j)
                                                            : s_sorted[j];
            aidx[k] = (int)(key & 0xffffffffu);
            x[k] = from_ordered_bits((unsigned)(key >> 32));
            if (sc.use_iou_pred) {
                const int l = anchor_level(geo, aidx[k]);
                const LevelView& lv = geo.lv[l];
                const int i = aidx[k] - lv.a_off;
                pa[k] = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)i * 4);
                const float* rp = lv.reg + head_offset(geo, lv, n, i, 0, 4);
                pd[k] = load_channels4(rp, head_cstride(geo, lv));
            }
        }
    }
    int n_pos = 0;
    if (n_cand == 1) {
        n_pos = 1;                                     // loss.py:218-219
    } else if (n_cand > 1) {
        n_pos = gmm_positive_prefix<SPL>(x, n_cand, lane, dbg.gmm ? dbg.gmm + (size_t)gi * 8 : nullptr,
                                         s_red[MODE == kFitOnly ? warp : 0]);
    }
    if (MODE == kSelectFused && lane == 0) atomicSub(ticket + kCtlSmLive + (smid & (kCtlMaxSms - 1)), 1u);
    if (dbg.gmm && n_cand <= 1 && lane < 8) dbg.gmm[(size_t)gi * 8 + lane] = 0.0;
#ifdef PAA_PROFILE_GMM        // measurement build only: cycles of scan+sort / EM in the w0 / w1 debug slots
    if (dbg.gmm && lane == 0) {
        dbg.gmm[(size_t)gi * 8 + 0] = (double)(prof_t1 - prof_t0);
        dbg.gmm[(size_t)gi * 8 + 1] = (double)(clock64() - prof_t1);
    }
#endif

    // labels of the positive prefix + this GT's share of the IoU normaliser (loss.py:228-230,331-333)
    double siou = 0.0;
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        const int j = lane + 32 * k;
        if (j < n_cand) {
            if (dbg.cand_idx) dbg.cand_idx[(size_t)gi * cap + j] = aidx[k];
            if (j < n_pos) {
                pos_list[(size_t)gi * cap + j] = aidx[k];
                paa_label[(size_t)n * geo.A + aidx[k]] = cls_label;
                if (dbg.paa_labels) dbg.paa_labels[(size_t)n * geo.A + aidx[k]] = cls_label;
                if (sc.use_iou_pred) {
                    const AnchorFrame f = anchor_frame(pa[k]);
                    const float4 pred = decode_box(pd[k], f);
                    const float4 tgt = decode_box(encode_box(gt, f), f);
                    siou += (double)iou_plus1(tgt, area_plus1(tgt), pred, area_plus1(pred));
                }
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

    // the last block to finish folds the per-GT partials in a fixed order (deterministic sums)
    unsigned my_ticket = 0;
    if (lane == 0) {
        __threadfence();
        my_ticket = atomicAdd(&ticket[0], 1u);
    }
    my_ticket = __shfl_sync(PAA_FULL, my_ticket, 0);
#ifdef PAA_PROFILE_GMM        // measurement build only: the fitting warp's whole life (up to its ticket) in the w0 slot
    if (dbg.gmm && lane == 0 && MODE == kFitOnly) dbg.gmm[(size_t)gi * 8 + 0] = (double)(clock64() - prof_t0);
#endif
    if (my_ticket == (unsigned)num_gt_total - 1u) {
        __threadfence();
        // eight independent loads per lane in flight (this runs after the slowest fit: it is pure kernel tail);
        // the order of the additions is fixed, so the totals are reproducible
        double cnt = 0.0, sum = 0.0;
        for (int g0 = lane; g0 < num_gt_total; g0 += 8 * PAA_WARP) {
            int c[8];
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int g = g0 + u * PAA_WARP;
                c[u] = g < num_gt_total ? __ldcg(part_npos + g) : 0;
                v[u] = g < num_gt_total ? __ldcg(part_siou + g) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                cnt += (double)c[u];
                sum += v[u];
            }
        }
        cnt = warp_sum(cnt);
        sum = warp_sum(sum);
        if (lane == 0) {
            local_norm[0] = cnt;
            local_norm[1] = sum;
            normalisers[0] = cnt;
            normalisers[1] = sum;
            __threadfence();
            ticket[kCtlReady] = 1u;         // hint: early bulk chunks stop here (the dependency wait is what orders)
        }
        // publish this rank's pair to every rank's exchange buffer over NVLink (lane r -> rank r): data,
        // system-scope fence, then the epoch that makes it visible to norm_wait_kernel over there
        if (px.world > 1) {
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
}

// First kernel of paa_loss when the peer exchange is on: lane r waits for rank r's contribution of this
// step in THIS rank's buffer (local memory, written remotely), then the totals are added in rank order --
// the same order on every rank, so all ranks normalise by bit-identical values.  Like the all-reduce it
// replaces (loss.py:22-28) the wait has no deadline by default: a rank that saves a checkpoint or stalls in
// its data loader is simply waited for.  With a deadline (PaaLossArgs::peer_timeout_s > 0) a contribution
// that does not arrive in time is an ERROR, never data: the kernel records {1, rank, epoch} in the
// host-mapped status words and traps, which fails every later CUDA call of the process.
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(PAA_WARP)
norm_wait_kernel(const PeerExchange px, double* __restrict__ normalisers) {
    // A programmatic dependent launch itself, and the loss pass behind it may start at once: its blocks become
    // resident, prefetch (and work on unscaled chunks, bulk_focal_early_kernel) while this rank's last EM fits run
    // and the peers' contributions travel; they wait for THIS kernel's end before they touch the normalisers.
    pdl_launch_dependents();
    pdl_wait();                       // the epoch and this rank's own contribution come from select_gmm_kernel
    const int lane = threadIdx.x;
    const double* own = px.buf[px.rank];
    const unsigned long long epoch = (unsigned long long)own[kPeerEpochOffset];   // set by select_gmm_kernel
    double cnt = 0.0, sum = 0.0;
    if (lane < px.world) {
        const volatile double* src = own + ((size_t)(epoch & 1ull) * kPeerMaxRanks + lane) * 4;
        const unsigned long long t0 = global_timer_ns();
        unsigned polls = 0;
        while (src[2] != (double)epoch) {
            if (++polls < 64u) continue;                  // the usual case: the peers are microseconds apart
            __nanosleep(200);
            if (px.timeout_ns != 0ull && global_timer_ns() - t0 > px.timeout_ns) {
                if (px.status) {
                    volatile int* st = px.status;
                    st[1] = lane;
                    st[2] = (int)epoch;
                    __threadfence_system();
                    st[0] = 1;
                    __threadfence_system();
                }
                __trap();
            }
        }
        __threadfence_system();
        cnt = src[0];
        sum = src[1];
    }
    double tc = 0.0, ts = 0.0;
    for (int r = 0; r < px.world; ++r) {
        tc += __shfl_sync(PAA_FULL, cnt, r);
        ts += __shfl_sync(PAA_FULL, sum, r);
    }
    if (lane == 0) {
        normalisers[0] = tc;
        normalisers[1] = ts;
    }
}

int launch_norm_wait(const PeerExchange& px, double* normalisers, cudaStream_t stream) {
    KernelTimer timer(PAA_KERNEL_NORM_WAIT, stream);
    PAA_PDL_LAUNCH(norm_wait_kernel, 1, PAA_WARP, stream, px, normalisers);
    return 0;
}

bool select_gmm_two_launch(const LossScalars& sc) {
    int split_above = kSplitAboveGts;
    if (const char* e = getenv("PAA_GMM_SPLIT_ABOVE")) split_above = atoi(e);       // measurement switch
    return sc.gt_capacity > split_above;
}

int launch_select_gmm(const Geometry& geo, const float* gt_boxes,
                      const int64_t* gt_labels, const LossScalars& sc, const LossWorkspace& ws,
                      const float* score_src, double* normalisers, const PeerExchange& px, const LossDebug& dbg,
                      cudaStream_t stream) {
    const int cap = geo.num_levels * sc.topk;
    const int grid = sc.gt_capacity;
    const int threads = geo.num_levels * PAA_WARP;
    const GtOffsets* gop = ws.go;
    KernelTimer timer(PAA_KERNEL_SELECT_GMM, stream);
#define PAA_SEL_LAUNCH_MODE(SPL, MODE, G, T)                                                           \
    PAA_PDL_LAUNCH((select_gmm_kernel<SPL, MODE>), G, T, stream, geo, gop, ws.gt_image, gt_boxes, gt_labels, sc, \
        ws.tile_gtmask, ws.matched, score_src, ws.seg_count, ws.seg_pool, ws.paa_label, ws.pos_list, ws.part_npos, \
        ws.part_siou,                                                                                  \
        ws.ticket, ws.local_norm, normalisers, px, dbg, ws.cand_sorted, ws.cand_n)
    // selection and fit as two launches for calls sized for many GTs (PAA_GMM_SPLIT_ABOVE overrides the threshold)
    const bool split = select_gmm_two_launch(sc);
#define PAA_SEL_LAUNCH(SPL)                                                                            \
    do {                                                                                               \
        if (split) {                                                                                   \
            PAA_SEL_LAUNCH_MODE(SPL, kSelectOnly, grid, threads);                                      \
            PAA_SEL_LAUNCH_MODE(SPL, kFitOnly, (grid + kFitWarps - 1) / kFitWarps, kFitWarps * PAA_WARP); \
        } else {                                                                                       \
            PAA_SEL_LAUNCH_MODE(SPL, kSelectFused, grid, threads);                                     \
        }                                                                                              \
    } while (0)
    if (cap <= 32) PAA_SEL_LAUNCH(1);
    else if (cap <= 64) PAA_SEL_LAUNCH(2);
    else PAA_SEL_LAUNCH(4);
#undef PAA_SEL_LAUNCH_MODE
#undef PAA_SEL_LAUNCH
    return 0;
}

}  // namespace paa
