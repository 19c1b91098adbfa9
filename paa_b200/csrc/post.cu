// PAA post-processing on sm_100a: candidate selection, label-aware NMS, top-D cut, score voting.
//
// Replaces PAAPostProcessor.forward (paa_core/modeling/rpn/paa/inference.py:84-159), entirely on the
// device and batched over the images of the call:
//   forward_for_single_feature_map  inference.py:36-82   (threshold, per-level top-k, decode, clip)
//   boxlist_ml_nms / _C.ml_nms       boxlist_ops.py:35-59, csrc/cuda/ml_nms.cu:13-136
//   kthvalue cut                     inference.py:114-122
//   score voting                     inference.py:123-157
//
// Pipeline (one launch each, no host round trip):
//   candidates : one pass over the logits (NCHW, coalesced); every (anchor, class) with
//                sigmoid(x) > thr is appended to its (image, level) list as (score, index) and
//                counted into a 2048-bin score histogram of that list
//   threshold  : per (image, level) the histogram bin that holds the k-th best score
//   filter     : streams the lists once: entries above the bin are selected, entries inside it are
//                remembered as the boundary set
//   select     : per (image, level) finishes the top-k from the boundary set (exact radix select),
//                orders the k survivors by candidate index, decodes + clips their boxes
//   rank       : per image sorts the <= L*k boxes by (label, score desc) by counting
//   segments   : start of every label run in the sorted order
//   nms_mask   : 64x64 suppression bit tiles, only where row and column runs share a label
//   nms_scan   : one warp per label run walks its boxes in score order (greedy suppression)
//   finish     : top-D score cut, compaction in pre-NMS index order, output rows
//   vote       : one warp per output row averages the same-class pre-NMS boxes
#include "post.h"

#include <climits>

namespace paa {

constexpr int kHistBins = 2048;
constexpr int kMaxTopN = 4096;
constexpr int kFilterBlocks = 8;       // blocks per (image, level) list in the filter pass
constexpr int kFilterPer = 4;          // entries per thread and trip
constexpr int kFilterChunk = 256 * kFilterPer;

struct PostWorkspace {
    // cleared at the start of every call
    int* cand_count;      // [N*L]
    int* sel_count;       // [N*L]
    int* bnd_count;       // [N*L]
    int* hist;            // [N*L*kHistBins]
    int* nms_big;         // [N]        image has a label run too long for the fused NMS kernel
    size_t zero_bytes;
    // fully written before read
    uint2* cand;          // [N * A*C]  (score bits, anchor*C + class) per (image, level) list
    unsigned* bnd;        // [N * A*C]  positions (within the list) of boundary-bin entries
    int* thr_bin;         // [N*L]
    int* n_above;         // [N*L]
    int* k_sel;           // [N*L]
    uint2* sel;           // [N*capN]   selected raw entries per (image, level), stride topn
    int* pre_cnt;         // [N*L]      boxes kept per level after decode / clip / size filter
    float4* pre_box;      // [N*capN]   level-major, stride topn per level
    float* pre_score;     // [N*capN]
    int* pre_label;       // [N*capN]
    int* total;           // [N]        boxes per image
    float4* s_box;        // [N*capN]   sorted by (label, score desc, position)
    float* s_score;
    int* s_label;
    int* s_pos;           // position (level*topn + j) of the sorted box
    int* seg_start;       // [N*(capN+1)]
    int* n_seg;           // [N]
    unsigned long long* mask;   // [N * capN * nbw]
    unsigned char* keep_sorted; // [N*capN]
    unsigned char* row_long;    // [N*capN]   1 = the row's label run is left to the mask + scan pair
    int* out_rank;        // [N*capN]   sorted index of every output row (for voting)
    unsigned char* flag_by_pos; // [N*capN] survivor flag by pre-NMS position
    int* rank_by_pos;     // [N*capN]   sorted index by pre-NMS position
    int* g_pos;           // [N*capN]   positions grouped by label (arbitrary order inside a label)
    float* g_score;       // [N*capN]
    size_t total_bytes;
};

static PostWorkspace carve_post(void* base, int N, int A, int C, int L, int topn) {
    PostWorkspace w;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* q = p ? p + off : nullptr;
        off += (bytes + 255) / 256 * 256;
        return q;
    };
    const size_t NL = (size_t)N * L, capN = (size_t)L * topn, NAC = (size_t)N * A * C;
    const size_t nbw = (capN + 63) / 64;
    w.cand_count = (int*)take(NL * 4);
    w.sel_count = (int*)take(NL * 4);
    w.bnd_count = (int*)take(NL * 4);
    w.hist = (int*)take(NL * kHistBins * 4);
    w.nms_big = (int*)take((size_t)N * 4);
    w.zero_bytes = off;
    w.cand = (uint2*)take(NAC * 8);
    w.bnd = (unsigned*)take(NAC * 4);
    w.thr_bin = (int*)take(NL * 4);
    w.n_above = (int*)take(NL * 4);
    w.k_sel = (int*)take(NL * 4);
    w.sel = (uint2*)take((size_t)N * capN * 8);
    w.pre_cnt = (int*)take(NL * 4);
    w.pre_box = (float4*)take((size_t)N * capN * 16);
    w.pre_score = (float*)take((size_t)N * capN * 4);
    w.pre_label = (int*)take((size_t)N * capN * 4);
    w.total = (int*)take((size_t)N * 4);
    w.s_box = (float4*)take((size_t)N * capN * 16);
    w.s_score = (float*)take((size_t)N * capN * 4);
    w.s_label = (int*)take((size_t)N * capN * 4);
    w.s_pos = (int*)take((size_t)N * capN * 4);
    w.seg_start = (int*)take((size_t)N * (capN + 1) * 4);
    w.n_seg = (int*)take((size_t)N * 4);
    w.mask = (unsigned long long*)take((size_t)N * capN * nbw * 8);
    w.keep_sorted = (unsigned char*)take((size_t)N * capN);
    w.row_long = (unsigned char*)take((size_t)N * capN);
    w.out_rank = (int*)take((size_t)N * capN * 4);
    w.flag_by_pos = (unsigned char*)take((size_t)N * capN);
    w.rank_by_pos = (int*)take((size_t)N * capN * 4);
    w.g_pos = (int*)take((size_t)N * capN * 4);
    w.g_score = (float*)take((size_t)N * capN * 4);
    w.total_bytes = off;
    return w;
}

size_t post_workspace_bytes(int num_images, int anchors_per_image, int num_classes, int num_levels,
                            int pre_nms_top_n) {
    if (num_images < 1 || anchors_per_image < 1 || num_classes < 1 || num_levels < 1 || pre_nms_top_n < 1)
        return 0;
    return carve_post(nullptr, num_images, anchors_per_image, num_classes, num_levels, pre_nms_top_n).total_bytes;
}

struct ImageSizes {
    float wh[PAA_MAX_IMAGES][2];
};

struct DecodeSpec {
    int mode;                 // PAA_DECODE_*
    float wx, wy, ww, wh, clip;
};

__device__ __forceinline__ int score_bin(float s) {
    int b = (int)(s * (float)kHistBins);
    return b < 0 ? 0 : (b >= kHistBins ? kHistBins - 1 : b);
}

// ---------------------------------------------------------------------------------------------
// candidates: one streaming pass over the logits.
//
// Which anchor or class a logit belongs to only matters for the few that pass the threshold, so each
// image's [C, H*W] block of a level is read as one flat array (the access pattern of a memcpy), a float4
// per lane, four float4 in flight per thread and the next chunk's four already issued.  The gate is a
// plain compare against logit(thr) (no transcendental) that leaves a 16-bit mask per thread.  Gated
// elements (a fraction of a percent on the fine levels) are pushed, with their position, into a queue in
// shared memory that lives across the block's chunks; whenever the queue holds a few warps' worth of work
// (and at the end) it is drained by ALL threads of the block: sigmoid, exact threshold test, the
// IoU-prediction sigmoid of the anchor, index arithmetic, and the append to the (image, level) list (one
// global atomic per (warp, list)) with its 2048-bin score histogram.  The expensive path therefore runs on
// full warps instead of on the one or two lanes of a warp that happen to hold a candidate.
// Chunks (4096 elements, never crossing an image) are dealt to the blocks round-robin, so the small dense
// levels (most of P6 / P7 passes the threshold) spread over the whole grid.
// ---------------------------------------------------------------------------------------------
constexpr int kCandThreads = 256;
constexpr int kCandVecs = 4;                                     // float4 per thread per chunk
constexpr int kCandChunk = kCandThreads * kCandVecs * 4;        // elements per chunk
constexpr int kCandDrainAt = 384;                                // drain when at least this many are queued
constexpr int kCandQueue = kCandDrainAt + kCandChunk;            // queue capacity
constexpr int kCandBlocksPerSM = 4;

struct CandPlan {
    unsigned chunk_off[PAA_MAX_LEVELS + 1];     // first chunk of each level in the global chunk order
    unsigned chunks_per_image[PAA_MAX_LEVELS];
    float inv_cpi[PAA_MAX_LEVELS];              // 1 / chunks_per_image (divmod_small)
    unsigned char vec[PAA_MAX_LEVELS];          // 16-byte loads allowed
    unsigned total;
};

struct CandChunk {
    int l, n;
    unsigned e_base, per_image;
    const float* src;
};

__device__ __forceinline__ void divmod_small(unsigned n, unsigned d, float inv, unsigned* q, unsigned* r);

__device__ __forceinline__ CandChunk cand_chunk(const Geometry& geo, const CandPlan& plan, unsigned ch) {
    CandChunk k;
    k.l = 0;
#pragma unroll
    for (int q = 1; q < PAA_MAX_LEVELS; ++q)
        if (q < geo.num_levels && ch >= plan.chunk_off[q]) k.l = q;
    const unsigned rel = ch - plan.chunk_off[k.l];
    unsigned n, in_image;
    divmod_small(rel, plan.chunks_per_image[k.l], plan.inv_cpi[k.l], &n, &in_image);      // no integer division here
    k.n = (int)n;
    k.e_base = in_image * kCandChunk;
    k.per_image = (unsigned)(geo.apl * geo.C) * (unsigned)geo.lv[k.l].hw;      // elements of one image
    k.src = geo.lv[k.l].cls + (size_t)k.n * k.per_image;
    return k;
}

__device__ __forceinline__ void cand_load(const CandChunk& k, bool vec, float4 (&x)[kCandVecs]) {
    if (vec && k.e_base + (unsigned)kCandChunk <= k.per_image) {
        // the usual chunk, whole inside its image: one pointer, four loads at constant offsets (block-uniform branch)
        const float4* p = reinterpret_cast<const float4*>(k.src + k.e_base) + threadIdx.x;
#pragma unroll
        for (int j = 0; j < kCandVecs; ++j) x[j] = __ldcs(p + j * kCandThreads);
        return;
    }
#pragma unroll
    for (int j = 0; j < kCandVecs; ++j) {
        const unsigned e = k.e_base + (unsigned)(j * kCandThreads + threadIdx.x) * 4u;
        if (vec && e + 4u <= k.per_image) {
            x[j] = __ldcs(reinterpret_cast<const float4*>(k.src + e));
        } else {
            x[j].x = (e < k.per_image) ? __ldg(k.src + e) : -INFINITY;
            x[j].y = (e + 1u < k.per_image) ? __ldg(k.src + e + 1u) : -INFINITY;
            x[j].z = (e + 2u < k.per_image) ? __ldg(k.src + e + 2u) : -INFINITY;
            x[j].w = (e + 3u < k.per_image) ? __ldg(k.src + e + 3u) : -INFINITY;
        }
    }
}

// n / d and n % d for n < 2^24 * d through one float multiply and a +-1 fix-up (d > 0, inv = 1.0f / d).
__device__ __forceinline__ void divmod_small(unsigned n, unsigned d, float inv, unsigned* q, unsigned* r) {
    unsigned qq = (unsigned)((float)n * inv);
    int rr = (int)(n - qq * d);
    if (rr < 0) {
        --qq;
        rr += (int)d;
    } else if (rr >= (int)d) {
        ++qq;
        rr -= (int)d;
    }
    *q = qq;
    *r = (unsigned)rr;
}

// The block's queue of gated elements (shared memory) and what every kernel variant does with it.
struct CandQueue {
    float* x;                 // [kCandQueue] gated logit
    unsigned* e;              // [kCandQueue] its element index inside the image's [C, H*W] block
    unsigned short* where;    // [kCandQueue] (image << 3) | level
    unsigned* hist2;          // [kHistBins / 2] score histogram of one list (big drains), two 16-bit bins a word
    int* cnt;
};

// all threads: exact test + score + append for every queued element (block-uniform call).  A big drain
// (a dense chunk: thousands of candidates of ONE list) counts its scores into a shared-memory histogram
// that is added to the list's global one at the end; small drains update the global histogram directly.
__device__ __forceinline__ void cand_drain(const Geometry& geo, const CandQueue& q, float thr, uint2* __restrict__ cand,
                                           int* __restrict__ cand_count, int* __restrict__ hist) {
    const int lane = threadIdx.x & 31;
    const float inv_c = 1.0f / (float)geo.C;
    const int total = *q.cnt;
    const bool big = total >= 4 * kCandThreads;
    int hist_seg = -1;
    if (big) {
        const unsigned where0 = q.where[0];
        hist_seg = (int)(where0 >> 3) * geo.num_levels + (int)(where0 & 7u);
        for (int b = threadIdx.x; b < kHistBins / 2; b += kCandThreads) q.hist2[b] = 0u;
        __syncthreads();
    }
    for (int t0 = 0; t0 < total; t0 += kCandThreads) {
        const int t = t0 + threadIdx.x;
        bool is = false;
        float score = 0.0f;
        unsigned entry = 0u;
        int seg = -1, n = 0, l = 0;
        if (t < total) {
            const float xval = q.x[t];
            const float p = 1.0f / (1.0f + expf(-xval));                   // inference.py:43
            if (p > thr) {                                                 // inference.py:48
                const unsigned where = q.where[t];
                n = (int)(where >> 3);
                l = (int)(where & 7u);
                const LevelView& lv = geo.lv[l];
                const unsigned hw = (unsigned)lv.hw;
                unsigned chn, loc, a, c;
                size_t iou_at;
                if (geo.nhwc) {
                    // channels-last: the element index inside the image's block IS anchor * C + class
                    entry = q.e[t];
                    divmod_small(entry, (unsigned)geo.C, inv_c, &a, &c);       // a = anchor of the level
                    iou_at = (size_t)n * lv.n_anchor + a;
                } else {
                    divmod_small(q.e[t], hw, 1.0f / (float)hw, &chn, &loc);
                    divmod_small(chn, (unsigned)geo.C, inv_c, &a, &c);
                    entry = (loc * (unsigned)geo.apl + a) * (unsigned)geo.C + c;   // anchor * C + class
                    iou_at = ((size_t)n * geo.apl + a) * hw + loc;
                }
                score = p;
                if (lv.iou != nullptr) {
                    const float xi = __ldg(lv.iou + iou_at);
                    const float qq = 1.0f / (1.0f + expf(-xi));            // inference.py:55
                    score = sqrtf(__fmul_rn(p, qq));                       // inference.py:56
                }
                seg = n * geo.num_levels + l;
                is = true;
            }
        }
        // one global atomic per (warp, list); queue neighbours mostly share their list
        const unsigned act = __ballot_sync(PAA_FULL, is);
        if (is) {
            const unsigned peers = __match_any_sync(act, seg);
            const int leader = __ffs(peers) - 1;
            int pos = 0;
            if (lane == leader) pos = atomicAdd(&cand_count[seg], __popc(peers));
            pos = __shfl_sync(peers, pos, leader) + __popc(peers & ((1u << lane) - 1u));
            cand[((size_t)n * geo.A + geo.lv[l].a_off) * geo.C + pos] = make_uint2(__float_as_uint(score), entry);
            const int bin = score_bin(score);
            if (seg == hist_seg) atomicAdd(&q.hist2[bin >> 1], 1u << (16 * (bin & 1)));   // <= kCandQueue < 65536 per bin
            else atomicAdd(&hist[(size_t)seg * kHistBins + bin], 1);
        }
    }
    __syncthreads();
    if (big) {
        int* gh = hist + (size_t)hist_seg * kHistBins;
        for (int b = threadIdx.x; b < kHistBins / 2; b += kCandThreads) {
            const unsigned v = q.hist2[b];
            if (v & 0xffffu) atomicAdd(&gh[2 * b], (int)(v & 0xffffu));
            if (v >> 16) atomicAdd(&gh[2 * b + 1], (int)(v >> 16));
        }
    }
    if (threadIdx.x == 0) *q.cnt = 0;
    __syncthreads();
}

// The gate over the kCandVecs float4 a thread holds (float4 j = elements e_first + (j * kCandThreads + threadIdx.x) * 4
// .. + 3 of image n, level l) and the push of the gated ones into the block's queue.  Warp-uniform control flow.
// Returns the queue's fill as this warp last saw it: right after its own push, or a plain read if it pushed nothing.
// The warp that pushes last in a round sees the round's final fill, so `__syncthreads_or(seen >= limit)` gives every
// thread of the block the same answer to "does the queue have to be drained" -- reading the counter after the barrier
// would race with the next round's pushes of faster warps.
__device__ __forceinline__ int cand_gate(const float4 (&x)[kCandVecs], float logit_gate, int n, int l, unsigned e_first,
                                          const CandQueue& q) {
    const int lane = threadIdx.x & 31;
    // bit 4j+t: element t of float4 j passes the gate (16 compares, no transcendental)
    // (most threads hold no gated element at all: one maximum over the 16 values decides that first)
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kCandVecs; ++j) mx = fmaxf(mx, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
    unsigned mask = 0u;
    if (mx > logit_gate) {
#pragma unroll
        for (int j = 0; j < kCandVecs; ++j) {
            mask |= (x[j].x > logit_gate ? 1u : 0u) << (4 * j);
            mask |= (x[j].y > logit_gate ? 1u : 0u) << (4 * j + 1);
            mask |= (x[j].z > logit_gate ? 1u : 0u) << (4 * j + 2);
            mask |= (x[j].w > logit_gate ? 1u : 0u) << (4 * j + 3);
        }
    }
    int seen;
    if (__any_sync(PAA_FULL, mask != 0u)) {
        // queue slots for the warp's gated elements: warp scan of the per-lane counts, one shared atomic
        const int mine = __popc(mask);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(PAA_FULL, incl, o);
            if (lane >= o) incl += v;
        }
        int base = 0;
        if (lane == 31) base = atomicAdd(q.cnt, incl);
        base = __shfl_sync(PAA_FULL, base, 31);
        seen = base + __shfl_sync(PAA_FULL, incl, 31);
        int slot = base + incl - mine;
        const unsigned short where = (unsigned short)((n << 3) | l);
#pragma unroll 1
        while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1u;
            const int j = b >> 2, t = b & 3;
            const float4 xs = j == 0 ? x[0] : (j == 1 ? x[1] : (j == 2 ? x[2] : x[3]));
            q.x[slot] = t == 0 ? xs.x : (t == 1 ? xs.y : (t == 2 ? xs.z : xs.w));
            q.e[slot] = e_first + (unsigned)(j * kCandThreads + threadIdx.x) * 4u + (unsigned)t;
            q.where[slot] = where;
            ++slot;
        }
    } else {
        seen = *reinterpret_cast<volatile int*>(q.cnt);
    }
    return seen;
}

__global__ void __launch_bounds__(kCandThreads, kCandBlocksPerSM)
post_candidates_kernel(const Geometry geo, const CandPlan plan, const float thr, const float logit_gate,
                       uint2* __restrict__ cand, int* __restrict__ cand_count, int* __restrict__ hist) {
    PAA_TRACE_SCOPE(8);
    __shared__ float q_x[kCandQueue];
    __shared__ unsigned q_e[kCandQueue];
    __shared__ unsigned short q_where[kCandQueue];
    __shared__ unsigned s_hist2[kHistBins / 2];
    __shared__ int q_cnt;
    const CandQueue q = {q_x, q_e, q_where, s_hist2, &q_cnt};
    if (threadIdx.x == 0) q_cnt = 0;
    __syncthreads();

    unsigned ch = blockIdx.x;
    if (ch >= plan.total) return;
    CandChunk cur = cand_chunk(geo, plan, ch);
    float4 x[kCandVecs], nx[kCandVecs];
    cand_load(cur, plan.vec[cur.l] != 0, x);
    for (;;) {
        const unsigned ch_next = ch + gridDim.x;
        const bool more = ch_next < plan.total;
        CandChunk nxt = cur;
        if (more) {
            nxt = cand_chunk(geo, plan, ch_next);
            cand_load(nxt, plan.vec[nxt.l] != 0, nx);
        }
        const int seen = cand_gate(x, logit_gate, cur.n, cur.l, cur.e_base, q);
        if (__syncthreads_or(seen >= kCandDrainAt) || !more) cand_drain(geo, q, thr, cand, cand_count, hist);
        if (!more) break;
        ch = ch_next;
        cur = nxt;
#pragma unroll
        for (int j = 0; j < kCandVecs; ++j) x[j] = nx[j];
    }
}

// ---------------------------------------------------------------------------------------------
// candidates, bulk-copy variant: the same pass with the logits brought on chip by 1-D bulk asynchronous copies
// (cp.async.bulk.shared.global + mbarrier: the copy engine behind TMA) instead of per-thread loads.
//
// One persistent block per SM; thread 0 keeps kRingStages copies of kRingStageBytes in flight (pieces of one image's
// block of one level, dealt to the blocks round-robin like the chunks above), all threads wait on a stage's mbarrier
// and walk it in rounds of kCandChunk elements with the gate / queue / drain code of the kernel above.  The block
// barrier that closes a round also releases the stage: thread 0 refills it right away.  Why: straight after the
// bench's flush the register-staged kernel reads at 3.5 TB/s (tools/stream_probe.cu: every register-staged variant,
// whatever its blocks per SM and loads in flight, stays between 3.6 and 3.9 TB/s on this buffer), the bulk-copy ring
// with one block per SM and >= 32 KB stages at 4.45 TB/s -- the requests of a whole stage reach the memory system as
// one descriptor, with no warp waiting on a scoreboard for them.  Needs every level's image block to be a multiple of
// 16 bytes at a 16-byte aligned address (C % 4 == 0: the usual 80 classes).
// MEASURED AND NOT KEPT AS THE DEFAULT (PAA_POST_RING=1 selects it): C4, 64 images, 254 us against 134 us for the
// kernel above (8 images: 52 against 31 us).  The copy side delivers; the consumer side cannot hide its own latencies
// with one block per SM -- every drain (an IoU-prediction load and an atomic with a returned value per candidate, ~2 us
// a trip; three dense P6 / P7 rounds of up to 4096 candidates per block) stalls the only block the SM has, where the
// four-blocks-per-SM kernel overlaps one block's drain with the others' streaming.  Two blocks per SM do not fit
// (2 x (96 KB ring + 49 KB queue)), and the probe's 3 x 32 KB x 2 blocks variant is no faster than register staging.
// ---------------------------------------------------------------------------------------------
constexpr int kRingStageBytes = 32 * 1024;
constexpr int kRingStageElems = kRingStageBytes / 4;
constexpr int kRingStages = 5;

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nPAA_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra PAA_DONE;\n"
        "bra PAA_WAIT;\nPAA_DONE:\n}\n" ::"r"(smem_addr(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

struct RingPlan {
    unsigned piece_off[PAA_MAX_LEVELS + 1];     // first piece of each level in the global piece order
    unsigned pieces_per_image[PAA_MAX_LEVELS];
    float inv_ppi[PAA_MAX_LEVELS];              // 1 / pieces_per_image
    unsigned total;
};

struct RingPiece {
    int l, n;
    unsigned e_base, count;                     // first element inside the image's block, elements in the piece
    const float* src;
};

__device__ __forceinline__ RingPiece ring_piece(const Geometry& geo, const RingPlan& plan, unsigned p) {
    RingPiece k;
    k.l = 0;
#pragma unroll
    for (int q = 1; q < PAA_MAX_LEVELS; ++q)
        if (q < geo.num_levels && p >= plan.piece_off[q]) k.l = q;
    const unsigned rel = p - plan.piece_off[k.l];
    unsigned n, in_image;
    divmod_small(rel, plan.pieces_per_image[k.l], plan.inv_ppi[k.l], &n, &in_image);
    k.n = (int)n;
    k.e_base = in_image * kRingStageElems;
    const unsigned per_image = (unsigned)(geo.apl * geo.C) * (unsigned)geo.lv[k.l].hw;
    k.count = min((unsigned)kRingStageElems, per_image - k.e_base);
    k.src = geo.lv[k.l].cls + (size_t)k.n * per_image + k.e_base;
    return k;
}

__global__ void __launch_bounds__(kCandThreads, 1)
post_candidates_ring_kernel(const Geometry geo, const RingPlan plan, const float thr, const float logit_gate,
                            uint2* __restrict__ cand, int* __restrict__ cand_count, int* __restrict__ hist) {
    PAA_TRACE_SCOPE(8);
    extern __shared__ __align__(128) unsigned char s_ring[];       // kRingStages x kRingStageBytes
    __shared__ unsigned long long s_full[kRingStages];
    __shared__ float q_x[kCandQueue];
    __shared__ unsigned q_e[kCandQueue];
    __shared__ unsigned short q_where[kCandQueue];
    __shared__ unsigned s_hist2[kHistBins / 2];
    __shared__ int q_cnt;
    const CandQueue q = {q_x, q_e, q_where, s_hist2, &q_cnt};
    if (threadIdx.x == 0) {
        q_cnt = 0;
#pragma unroll
        for (int s = 0; s < kRingStages; ++s) mbar_init(&s_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (blockIdx.x >= plan.total) return;
    // producer state (thread 0): the next piece to request and the stage it goes to
    unsigned p_issue = blockIdx.x;
    int s_issue = 0;
    auto issue = [&]() {
        if (p_issue < plan.total) {
            const RingPiece pc = ring_piece(geo, plan, p_issue);
            mbar_expect_tx(&s_full[s_issue], pc.count * 4u);
            bulk_load(s_ring + (size_t)s_issue * kRingStageBytes, pc.src, pc.count * 4u, &s_full[s_issue]);
        }
        p_issue += gridDim.x;
        s_issue = s_issue + 1 == kRingStages ? 0 : s_issue + 1;
    };
    if (threadIdx.x == 0)
        for (int s = 0; s < kRingStages; ++s) issue();

    int stage = 0;
    unsigned parity = 0u;
    for (unsigned p = blockIdx.x; p < plan.total; p += gridDim.x) {
        const RingPiece pc = ring_piece(geo, plan, p);
        const bool last_piece = p + gridDim.x >= plan.total;
        mbar_wait(&s_full[stage], parity);
        const float4* sp = reinterpret_cast<const float4*>(s_ring + (size_t)stage * kRingStageBytes);
        for (unsigned r0 = 0; r0 < pc.count; r0 += kCandChunk) {
            float4 x[kCandVecs];
#pragma unroll
            for (int j = 0; j < kCandVecs; ++j) {
                const unsigned e = r0 + (unsigned)(j * kCandThreads + threadIdx.x) * 4u;     // count is a multiple of 4
                x[j] = e < pc.count ? sp[e >> 2] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            }
            const bool last_round = r0 + kCandChunk >= pc.count;
            const int seen = cand_gate(x, logit_gate, pc.n, pc.l, pc.e_base + r0, q);
            const int full = __syncthreads_or(seen >= kCandDrainAt);     // every thread has its part of the round in registers
            if (last_round && threadIdx.x == 0) {
                // the stage is free: order the block's reads of it before the copy engine's writes, refill
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue();
            }
            if (full || (last_piece && last_round)) cand_drain(geo, q, thr, cand, cand_count, hist);
        }
        stage = stage + 1 == kRingStages ? 0 : stage + 1;
        parity ^= (stage == 0) ? 1u : 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// threshold bin per (image, level)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
post_threshold_kernel(const int* __restrict__ cand_count, const int* __restrict__ hist, int topn,
                      int* __restrict__ thr_bin, int* __restrict__ n_above, int* __restrict__ k_sel) {
    PAA_TRACE_SCOPE(9);
    __shared__ int s_suffix[256];
    const int seg = blockIdx.x;
    const int count = cand_count[seg];
    const int k = min(count, topn);                                      // inference.py:49-50
    if (count <= topn) {
        if (threadIdx.x == 0) {
            thr_bin[seg] = -1;
            n_above[seg] = count;
            k_sel[seg] = k;
        }
        return;
    }
    constexpr int per = kHistBins / 256;
    const int* h = hist + (size_t)seg * kHistBins;
    int local[per];
    int mine = 0;
#pragma unroll
    for (int j = 0; j < per; ++j) {
        local[j] = h[threadIdx.x * per + j];
        mine += local[j];
    }
    s_suffix[threadIdx.x] = mine;
    __syncthreads();
    // suffix sums over threads (bins are ascending in score)
    for (int off = 1; off < 256; off <<= 1) {
        int v = (threadIdx.x + off < 256) ? s_suffix[threadIdx.x + off] : 0;
        __syncthreads();
        s_suffix[threadIdx.x] += v;
        __syncthreads();
    }
    const int incl = s_suffix[threadIdx.x];                 // entries in bins >= my first bin
    const int excl = incl - mine;                           // entries in bins above my range
    if (excl < k && incl >= k) {
        int above = excl;
        for (int j = per - 1; j >= 0; --j) {
            if (above + local[j] >= k) {
                thr_bin[seg] = threadIdx.x * per + j;
                n_above[seg] = above;
                k_sel[seg] = k;
                break;
            }
            above += local[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// filter: selected (above the bin) and boundary (inside the bin) entries
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
post_filter_kernel(const Geometry geo, const uint2* __restrict__ cand, const int* __restrict__ cand_count,
                   const int* __restrict__ thr_bin, int topn, uint2* __restrict__ sel,
                   int* __restrict__ sel_count, unsigned* __restrict__ bnd, int* __restrict__ bnd_count) {
    PAA_TRACE_SCOPE(10);
    const int seg = blockIdx.x / kFilterBlocks;
    const int part = blockIdx.x - seg * kFilterBlocks;
    const int n = seg / geo.num_levels, l = seg - n * geo.num_levels;
    const int count = cand_count[seg];
    if (part * kFilterChunk >= count) return;
    const int tb = thr_bin[seg];
    const size_t list_off = ((size_t)n * geo.A + geo.lv[l].a_off) * geo.C;
    const uint2* list = cand + list_off;
    unsigned* blist = bnd + list_off;
    uint2* out = sel + (size_t)seg * topn;
    const int lane = threadIdx.x & 31;
    // a block takes 1024 consecutive entries at a time: four coalesced loads per thread in flight, then ONE atomic per
    // warp and output list for all four (the kernel is a chain of memory round trips, not bandwidth)
    for (int e0 = part * kFilterChunk; e0 < count; e0 += kFilterBlocks * kFilterChunk) {
        uint2 v[kFilterPer];
        int e[kFilterPer];
#pragma unroll
        for (int k = 0; k < kFilterPer; ++k) {
            e[k] = e0 + k * 256 + threadIdx.x;
            v[k] = e[k] < count ? list[e[k]] : make_uint2(0u, 0u);
        }
        unsigned ms[kFilterPer], mb[kFilterPer];
        int n_sel = 0, n_bnd = 0;
#pragma unroll
        for (int k = 0; k < kFilterPer; ++k) {
            const int b = score_bin(__uint_as_float(v[k].x));
            ms[k] = __ballot_sync(PAA_FULL, e[k] < count && b > tb);
            mb[k] = __ballot_sync(PAA_FULL, e[k] < count && b == tb);
            n_sel += __popc(ms[k]);
            n_bnd += __popc(mb[k]);
        }
        int base_s = 0, base_b = 0;
        if (lane == 0 && n_sel) base_s = atomicAdd(&sel_count[seg], n_sel);
        if (lane == 0 && n_bnd) base_b = atomicAdd(&bnd_count[seg], n_bnd);
        base_s = __shfl_sync(PAA_FULL, base_s, 0);
        base_b = __shfl_sync(PAA_FULL, base_b, 0);
        const unsigned below = (1u << lane) - 1u;
#pragma unroll
        for (int k = 0; k < kFilterPer; ++k) {
            if ((ms[k] >> lane) & 1u) out[base_s + __popc(ms[k] & below)] = v[k];
            if ((mb[k] >> lane) & 1u) blist[base_b + __popc(mb[k] & below)] = (unsigned)e[k];
            base_s += __popc(ms[k]);
            base_b += __popc(mb[k]);
        }
    }
}

// One step of a most-significant-digit-first radix select over a 256-bin histogram in shared memory:
// the largest digit d >= 1 with hist[d] + hist[d+1] + ... + hist[255] >= need, or 0 when there is none, and how
// many entries are still needed among the keys with that digit.  Executed by ONE warp (lane handles eight
// bins, warp scan over lanes) instead of a serial walk over 255 bins.
__device__ __forceinline__ void radix_pick_digit(const int* s_hist, int need, int lane, int* digit, int* remaining) {
    int v[8];
    int loc = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int bin = 255 - (8 * lane + j);
        v[j] = bin >= 1 ? s_hist[bin] : 0;                 // digit 0 is the fall-through, not a candidate
        loc += v[j];
    }
    int incl = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(PAA_FULL, incl, o);
        if (lane >= o) incl += t;
    }
    const int excl = incl - loc;                           // entries in the bins above this lane's eight
    const bool crossing = (excl < need) && (need <= incl);
    const unsigned m = __ballot_sync(PAA_FULL, crossing);
    const int total = __shfl_sync(PAA_FULL, incl, 31);
    int d = 0, rem = need - total;                         // no crossing: digit 0
    if (m) {
        const int src = __ffs(m) - 1;
        if (lane == src) {
            rem = need - excl;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int bin = 255 - (8 * lane + j);
                if (d == 0 && bin >= 1) {
                    if (v[j] >= rem) d = bin;
                    else rem -= v[j];
                }
            }
        }
        d = __shfl_sync(PAA_FULL, d, src);
        rem = __shfl_sync(PAA_FULL, rem, src);
    }
    *digit = d;
    *remaining = rem;
}

// ---------------------------------------------------------------------------------------------
// select: finish top-k, canonical order, decode + clip
//
// One block of 256 threads per (image, level) -- every list of a 64-image call is resident at once (four blocks per
// SM; the 1024-thread version ran 320 lists in three waves of 148).  A thread owns EPT = topn / 256 entries in
// registers.  Order: the k survivors must come out by ascending candidate index (the order nonzero() enumerates
// them, inference.py:66).  Instead of a bitonic sort (55 block barriers for 1024 keys) the entries are dealt into
// 1024 buckets by the high bits of the index (counting sort: histogram, scan, scatter -- three barriers) and every
// entry finds its rank inside its bucket by counting (a bucket holds a few entries: candidates of one object
// cluster, but over many buckets); the thread that ranked an entry decodes its box and writes the row.
// ---------------------------------------------------------------------------------------------
constexpr int kSelectThreads = 256;
constexpr int kSelectBuckets = 1024;
constexpr int kSelectCached = 2;            // boundary keys a thread keeps in registers
constexpr int kSelectSmallB = 1024;         // boundary sets up to this size are ranked by counting

// key of a candidate for "better first": higher score, then lower candidate index
__device__ __forceinline__ unsigned long long better_key(uint2 v) {
    return ((unsigned long long)v.x << 32) | (unsigned long long)(0xffffffffu - v.y);
}

struct SelectSmem {                          // dynamic: [cap] index, [cap] score bits, [cap] keep flags
    unsigned* idx;
    unsigned* sc;
    unsigned char* keep;
};

template <int EPT>
__global__ void __launch_bounds__(kSelectThreads)
post_select_kernel(const Geometry geo, const ImageSizes sizes, const uint2* __restrict__ cand,
                   const unsigned* __restrict__ bnd, const int* __restrict__ bnd_count,
                   const int* __restrict__ n_above, const int* __restrict__ k_sel,
                   const int* __restrict__ thr_bin, int topn, float min_size, const DecodeSpec dec,
                   uint2* __restrict__ sel,
                   int* __restrict__ pre_cnt, float4* __restrict__ pre_box, float* __restrict__ pre_score,
                   int* __restrict__ pre_label) {
    PAA_TRACE_SCOPE(11);
    constexpr int kCap = EPT * kSelectThreads;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    unsigned* s_idx = reinterpret_cast<unsigned*>(s_dyn);
    unsigned* s_sc = s_idx + kCap;
    unsigned char* s_keep = reinterpret_cast<unsigned char*>(s_sc + kCap);
    __shared__ int s_bucket[kSelectBuckets + 1];       // counts, then exclusive starts
    __shared__ unsigned long long s_bkey[kSelectSmallB];
    __shared__ int s_hist[256];
    __shared__ int s_wsum[kSelectThreads / 32];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_need, s_fill;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg = blockIdx.x;
    const int n = seg / geo.num_levels, l = seg - n * geo.num_levels;
    const LevelView& lv = geo.lv[l];
    const size_t list_off = ((size_t)n * geo.A + lv.a_off) * geo.C;
    const uint2* list = cand + list_off;
    const unsigned* blist = bnd + list_off;
    uint2* mysel = sel + (size_t)seg * topn;
    const int k = k_sel[seg];
    const int above = n_above[seg];
    const int B = (thr_bin[seg] >= 0) ? bnd_count[seg] : 0;
    const int need = k - above;                 // entries still to take from the boundary bin

    for (int b = threadIdx.x; b <= kSelectBuckets; b += kSelectThreads) s_bucket[b] = 0;
    if (threadIdx.x == 0) s_fill = 0;

    if (B > 0 && need > 0 && B <= kSelectSmallB) {
        // the usual case (a histogram bin holds a handful of the list's scores): rank by counting
        for (int e = threadIdx.x; e < B; e += kSelectThreads) s_bkey[e] = better_key(list[blist[e]]);
        __syncthreads();
        for (int e = threadIdx.x; e < B; e += kSelectThreads) {
            const unsigned long long mine = s_bkey[e];
            int better = 0;
            for (int j = 0; j < B; ++j) better += s_bkey[j] > mine ? 1 : 0;          // keys are unique
            if (better < need)
                mysel[above + better] = make_uint2((unsigned)(mine >> 32), 0xffffffffu - (unsigned)mine);
        }
        __syncthreads();
    } else if (B > 0 && need > 0) {
        // exact radix select (8 bits x 8 passes, most significant first) of the `need` best keys; a thread
        // keeps its first kSelectCached boundary keys in registers so that the passes do not go back to memory
        unsigned long long ck[kSelectCached];
        uint2 cv[kSelectCached];
#pragma unroll
        for (int c = 0; c < kSelectCached; ++c) {
            const int e = threadIdx.x + c * kSelectThreads;
            cv[c] = make_uint2(0u, 0u);
            if (e < B) cv[c] = list[blist[e]];
            ck[c] = better_key(cv[c]);
        }
        if (threadIdx.x == 0) {
            s_prefix = 0ull;
            s_need = need;
        }
        __syncthreads();
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = 56 - 8 * pass;
            for (int b = threadIdx.x; b < 256; b += kSelectThreads) s_hist[b] = 0;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            const unsigned long long hi_mask = (pass == 0) ? 0ull : (~0ull << (shift + 8));
#pragma unroll
            for (int c = 0; c < kSelectCached; ++c)
                if (threadIdx.x + c * kSelectThreads < B && (ck[c] & hi_mask) == prefix)
                    atomicAdd(&s_hist[(int)((ck[c] >> shift) & 0xff)], 1);
            for (int e = threadIdx.x + kSelectCached * kSelectThreads; e < B; e += kSelectThreads) {
                const unsigned long long key = better_key(list[blist[e]]);
                if ((key & hi_mask) == prefix) atomicAdd(&s_hist[(int)((key >> shift) & 0xff)], 1);
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                int d, remaining;
                radix_pick_digit(s_hist, s_need, threadIdx.x, &d, &remaining);
                if (threadIdx.x == 0) {
                    s_need = remaining;           // still needed among keys with this digit
                    s_prefix = prefix | ((unsigned long long)d << shift);
                }
            }
            __syncthreads();
        }
        // keys are unique, so exactly `need` of them are >= the selected key
        const unsigned long long tkey = s_prefix;
#pragma unroll
        for (int c = 0; c < kSelectCached; ++c)
            if (threadIdx.x + c * kSelectThreads < B && ck[c] >= tkey) mysel[above + atomicAdd(&s_fill, 1)] = cv[c];
        for (int e = threadIdx.x + kSelectCached * kSelectThreads; e < B; e += kSelectThreads) {
            const uint2 v = list[blist[e]];
            if (better_key(v) >= tkey) mysel[above + atomicAdd(&s_fill, 1)] = v;
        }
        __syncthreads();
    } else {
        __syncthreads();
    }

    // ---- canonical order: counting sort over index buckets + rank inside the bucket ----------------
    // bucket = index >> shift with the level's index range (anchors * classes) spread over <= 1024 buckets
    const unsigned range = (unsigned)lv.n_anchor * (unsigned)geo.C;
    int shift = 0;
    while ((range >> shift) >= (unsigned)kSelectBuckets) ++shift;
    uint2 ent[EPT];
    int slot[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        const int t = threadIdx.x + i * kSelectThreads;
        slot[i] = -1;
        if (t < k) {
            ent[i] = mysel[t];
            slot[i] = atomicAdd(&s_bucket[ent[i].y >> shift], 1);
        }
    }
    __syncthreads();
    // exclusive scan of the 1024 counts: four per thread, warp scan, warp totals
    {
        constexpr int per = kSelectBuckets / kSelectThreads;
        int c[per];
        int sum = 0;
#pragma unroll
        for (int j = 0; j < per; ++j) {
            c[j] = s_bucket[threadIdx.x * per + j];
            sum += c[j];
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(PAA_FULL, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; ++w) before += s_wsum[w];
        int run = before + incl - sum;
#pragma unroll
        for (int j = 0; j < per; ++j) {
            s_bucket[threadIdx.x * per + j] = run;
            run += c[j];
        }
        if (threadIdx.x == kSelectThreads - 1) s_bucket[kSelectBuckets] = run;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < EPT; ++i)
        if (slot[i] >= 0) {
            const int p = s_bucket[ent[i].y >> shift] + slot[i];
            s_idx[p] = ent[i].y;
            s_sc[p] = ent[i].x;
        }
    __syncthreads();
    const float img_w = sizes.wh[n][0], img_h = sizes.wh[n][1];
    bool any_drop = false;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        const int p = threadIdx.x + i * kSelectThreads;           // position in bucket order
        if (p < k) {
            const unsigned my_idx = s_idx[p];
            const int b0 = s_bucket[my_idx >> shift], b1 = s_bucket[(my_idx >> shift) + 1];
            int t = b0;                                           // = rank of the candidate
            for (int j = b0; j < b1; ++j) t += s_idx[j] < my_idx ? 1 : 0;
            const int ai = (int)(my_idx / (unsigned)geo.C);
            const float4 a = ldg4(lv.anchors + (size_t)n * geo.anchor_image_stride + (size_t)ai * 4);
            const float* rp = lv.reg + head_offset(geo, lv, n, ai, 0, 4);
            const float4 d = load_channels4(rp, head_cstride(geo, lv));
            float4 box;
            if (dec.mode == PAA_DECODE_LEGACY) box = decode_box_legacy(d, a, dec.wx, dec.wy, dec.ww, dec.wh, dec.clip);
            else if (dec.mode == PAA_DECODE_LTRB)                         // fcos/inference.py:93-98
                box = make_float4(__fsub_rn(a.x, d.x), __fsub_rn(a.y, d.y), __fadd_rn(a.z, d.z), __fadd_rn(a.w, d.w));
            else box = decode_box(d, anchor_frame(a));                    // inference.py:71-74
            box.x = fminf(fmaxf(box.x, 0.0f), img_w - 1.0f);              // bounding_box.py:214-219
            box.y = fminf(fmaxf(box.y, 0.0f), img_h - 1.0f);
            box.z = fminf(fmaxf(box.z, 0.0f), img_w - 1.0f);
            box.w = fminf(fmaxf(box.w, 0.0f), img_h - 1.0f);
            const float ws = __fadd_rn(__fsub_rn(box.z, box.x), 1.0f);    // boxlist_ops.py:62-76
            const float hs = __fadd_rn(__fsub_rn(box.w, box.y), 1.0f);
            const bool drop = !((ws >= min_size) && (hs >= min_size));
            any_drop |= drop;
            s_keep[t] = drop ? 0 : 1;
            const size_t o = (size_t)seg * topn + t;
            pre_box[o] = box;
            pre_score[o] = __uint_as_float(s_sc[p]);
            pre_label[o] = (int)(my_idx % (unsigned)geo.C) + 1;           // inference.py:69
        }
    }
    const int dropped = __syncthreads_or(any_drop ? 1 : 0);
    if (threadIdx.x == 0) {
        int cnt = k;
        if (dropped) {
            // remove_small_boxes dropped something (non-finite boxes only when min_size == 0): close the
            // gaps in order.  Serial on purpose -- this path is not expected to run.
            cnt = 0;
            const size_t o = (size_t)seg * topn;
            for (int r = 0; r < k; ++r) {
                if (!s_keep[r]) continue;
                if (cnt != r) {
                    pre_box[o + cnt] = pre_box[o + r];
                    pre_score[o + cnt] = pre_score[o + r];
                    pre_label[o + cnt] = pre_label[o + r];
                }
                ++cnt;
            }
        }
        pre_cnt[seg] = cnt;
    }
}

static size_t select_smem_bytes(int ept) { return (size_t)ept * kSelectThreads * 9; }

// ---------------------------------------------------------------------------------------------
// rank: per image, order all boxes by (label asc, score desc, position asc) by counting
// ---------------------------------------------------------------------------------------------
constexpr int kRankThreads = 256;
constexpr unsigned long long kNoKey = ~0ull;

__device__ __forceinline__ unsigned long long nms_key(int label, float score, int pos) {
    const unsigned sb = ~__float_as_uint(score);          // scores are >= 0: descending score = ascending ~bits
    return ((unsigned long long)(unsigned)label << 48) | ((unsigned long long)sb << 16) | (unsigned long long)pos;
}

__global__ void __launch_bounds__(kRankThreads)
post_rank_kernel(int L, int topn, const int* __restrict__ pre_cnt, const float4* __restrict__ pre_box,
                 const float* __restrict__ pre_score, const int* __restrict__ pre_label,
                 float4* __restrict__ s_box, float* __restrict__ s_score, int* __restrict__ s_label,
                 int* __restrict__ s_pos, int* __restrict__ total) {
    PAA_TRACE_SCOPE(12);
    __shared__ unsigned long long s_keys[kRankThreads];
    __shared__ int s_cnt[PAA_MAX_LEVELS];
    const int n = blockIdx.y;
    const int capN = L * topn;
    if (threadIdx.x < L) s_cnt[threadIdx.x] = pre_cnt[n * L + threadIdx.x];
    __syncthreads();
    const int pos = blockIdx.x * kRankThreads + threadIdx.x;
    const size_t base = (size_t)n * capN;
    unsigned long long mine = kNoKey;
    if (pos < capN && (pos % topn) < s_cnt[pos / topn])
        mine = nms_key(pre_label[base + pos], pre_score[base + pos], pos);
    int rank = 0;
    for (int q0 = 0; q0 < capN; q0 += kRankThreads) {
        const int q = q0 + threadIdx.x;
        unsigned long long k = kNoKey;
        if (q < capN && (q % topn) < s_cnt[q / topn]) k = nms_key(pre_label[base + q], pre_score[base + q], q);
        __syncthreads();
        s_keys[threadIdx.x] = k;
        __syncthreads();
        if (mine != kNoKey) {
#pragma unroll 8
            for (int j = 0; j < kRankThreads; ++j) rank += (s_keys[j] < mine) ? 1 : 0;
        }
    }
    if (mine != kNoKey) {
        s_box[base + rank] = pre_box[base + pos];
        s_score[base + rank] = pre_score[base + pos];
        s_label[base + rank] = pre_label[base + pos];
        s_pos[base + rank] = pos;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int t = 0;
        for (int l = 0; l < L; ++l) t += s_cnt[l];
        total[n] = t;
    }
}

// ---------------------------------------------------------------------------------------------
// Fused-path ordering for labels known to lie in [0, C]: group the boxes of an image by label with a
// shared-memory histogram (which is also the table of label runs), then rank every box inside its own
// run only -- O(n * run length) instead of O(n^2).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
post_group_kernel(int L, int topn, int C, const int* __restrict__ pre_cnt, const float* __restrict__ pre_score,
                  const int* __restrict__ pre_label, int* __restrict__ g_pos, float* __restrict__ g_score,
                  int* __restrict__ seg_start, int* __restrict__ n_seg, int* __restrict__ total) {
    PAA_TRACE_SCOPE(13);
    extern __shared__ int s_lab[];          // [C+2] run starts, then [C+2] fill cursors
    __shared__ int s_cnt[PAA_MAX_LEVELS];
    const int n = blockIdx.x;
    const int capN = L * topn;
    const size_t base = (size_t)n * capN;
    int* s_start = s_lab;
    int* s_cur = s_lab + (C + 2);
    if (threadIdx.x < L) s_cnt[threadIdx.x] = pre_cnt[n * L + threadIdx.x];
    for (int c = threadIdx.x; c < C + 2; c += 1024) s_start[c] = 0;
    __syncthreads();
    for (int pos = threadIdx.x; pos < capN; pos += 1024)
        if ((pos % topn) < s_cnt[pos / topn]) {
            const int lab = min(max(pre_label[base + pos], 0), C);
            atomicAdd(&s_start[lab + 1], 1);                     // counts shifted by one: scan gives starts
        }
    __syncthreads();
    if (threadIdx.x < 32) {
        // inclusive scan of the C + 2 counts by one warp, 32 at a time: s_start[c] = first sorted index of label c
        int carry = 0;
        for (int c0 = 0; c0 < C + 2; c0 += 32) {
            const int c = c0 + (int)threadIdx.x;
            int v = c < C + 2 ? s_start[c] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(PAA_FULL, v, o);
                if ((int)threadIdx.x >= o) v += t;
            }
            v += carry;
            if (c < C + 2) s_start[c] = v;
            carry = __shfl_sync(PAA_FULL, v, 31);
        }
    }
    __syncthreads();
    int* ss = seg_start + (size_t)n * (capN + 1);
    for (int c = threadIdx.x; c < C + 2; c += 1024) {
        s_cur[c] = s_start[c];
        ss[c] = s_start[c];
    }
    if (threadIdx.x == 0) {
        n_seg[n] = C + 1;
        total[n] = s_start[C + 1];
    }
    __syncthreads();
    for (int pos = threadIdx.x; pos < capN; pos += 1024)
        if ((pos % topn) < s_cnt[pos / topn]) {
            const int lab = min(max(pre_label[base + pos], 0), C);
            const int slot = atomicAdd(&s_cur[lab], 1);
            g_pos[base + slot] = pos;
            g_score[base + slot] = pre_score[base + pos];
        }
}

__global__ void __launch_bounds__(256)
post_class_rank_kernel(int capN, const int* __restrict__ total, const int* __restrict__ seg_start,
                       const int* __restrict__ g_pos, const float* __restrict__ g_score,
                       const float4* __restrict__ pre_box, const int* __restrict__ pre_label,
                       float4* __restrict__ s_box, float* __restrict__ s_score, int* __restrict__ s_label,
                       int* __restrict__ s_pos) {
    PAA_TRACE_SCOPE(14);
    const int n = blockIdx.y;
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= total[n]) return;
    const size_t base = (size_t)n * capN;
    const int pos = g_pos[base + e];
    const float sc = g_score[base + e];
    const int lab = pre_label[base + pos];
    const int* ss = seg_start + (size_t)n * (capN + 1);
    const int a = ss[lab], b = ss[lab + 1];
    int rank = 0;
    for (int j = a; j < b; ++j) {                                // better = higher score, then lower position
        const float sj = g_score[base + j];
        const int pj = g_pos[base + j];
        rank += (sj > sc || (sj == sc && pj < pos)) ? 1 : 0;
    }
    const size_t o = base + a + rank;
    s_box[o] = pre_box[base + pos];
    s_score[o] = sc;
    s_label[o] = lab;
    s_pos[o] = pos;
}

// ---------------------------------------------------------------------------------------------
// segments: first sorted index of every run of equal labels (one block per image)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
post_segments_kernel(int capN, const int* __restrict__ total, const int* __restrict__ s_label,
                     int* __restrict__ seg_start, int* __restrict__ n_seg) {
    PAA_TRACE_SCOPE(15);
    __shared__ int s_warp[32];
    __shared__ int s_running;
    const int n = blockIdx.x;
    const int cnt = total[n];
    const int* lab = s_label + (size_t)n * capN;
    int* out = seg_start + (size_t)n * (capN + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    for (int r0 = 0; r0 < cnt; r0 += 1024) {
        const int r = r0 + threadIdx.x;
        const bool head = (r < cnt) && (r == 0 || lab[r] != lab[r - 1]);
        const unsigned m = __ballot_sync(PAA_FULL, head);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = s_running;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (head) out[before + __popc(m & ((1u << lane) - 1u))] = r;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 32; ++w) t += s_warp[w];
            s_running += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[s_running] = cnt;
        n_seg[n] = s_running;
    }
}

// ---------------------------------------------------------------------------------------------
// nms_runs: greedy suppression of one label run by ONE warp, without a bit mask in memory
// (csrc/cuda/ml_nms.cu:13-24,55-70,116-128 -- same pairs, same comparison, same keep set).
//
// Boxes are sorted by (label, score desc).  The warp walks its run 32 boxes at a time; lane = box.  (1) every box of
// the group is tested against the run's boxes KEPT so far (their coordinates wait in the warp's slice of shared
// memory, broadcast reads): a suppressed box never suppresses anything, so rows of suppressed boxes -- most of a
// crowded run -- are never evaluated, which the mask formulation cannot know; (2) the 32 x 32 triangle inside the
// group as one 32-bit word per lane, columns that (1) already removed skipped; (3) the serial greedy scan of the
// group over those words (shuffles); (4) the kept boxes join the list.  Against the two-kernel mask + scan pair
// (34.7 M warp instructions and a 64-bit word per (row, 64 columns) through memory at 64 images): no global
// intermediate, no lanes looping over other labels' columns, one launch less.  Runs longer than kRunMax boxes (a
// crowd of one class) flag their image for the mask + scan pair below, which is parallel over the rows of a run.
// ---------------------------------------------------------------------------------------------
constexpr int kRunWarps = 4;
constexpr int kRunMax = 128;

// iou_plus1(a, b) > thr, decided without the division wherever the operands leave no doubt: with t = thr * union,
// inter > t * (1 + 2^-20) makes the rounded quotient exceed thr and inter < t * (1 - 2^-20) keeps it below (the
// quotient and t each carry a relative rounding error of 2^-24); only pairs inside that band -- and anything
// degenerate (non-positive union or threshold, NaN) -- take the exact IEEE division.  Same decision as
// `iou_plus1(...) > thr` for every input.
struct IouGate {
    float thr, hi, lo;
    bool fast;
};
__device__ __forceinline__ IouGate iou_gate(float thr) {
    IouGate g;
    g.thr = thr;
    g.hi = thr * (1.0f + 9.5367431640625e-7f);
    g.lo = thr * (1.0f - 9.5367431640625e-7f);
    g.fast = thr > 0.0f;
    return g;
}
// The branch-free part: returns "certainly above", *unsure = neither certain answer applies.
__device__ __forceinline__ bool iou_certain(float4 a, float area_a, float4 b, float area_b, const IouGate& g,
                                            bool* unsure) {
    const float w = __fadd_rn(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 1.0f);
    const float h = __fadd_rn(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 1.0f);
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    const bool overlap = (w > 0.0f) && (h > 0.0f);
    const bool pos = g.fast && uni > 0.0f;
    // no overlap (or a NaN coordinate): the quotient is 0 (or NaN), which never exceeds a positive threshold
    const bool yes = overlap && pos && inter > g.hi * uni;
    const bool no = g.fast && (!overlap || (pos && inter < g.lo * uni));
    *unsure = !(yes || no);
    return yes;
}
__device__ __forceinline__ bool iou_exceeds(float4 a, float area_a, float4 b, float area_b, const IouGate& g) {
    bool unsure;
    const bool yes = iou_certain(a, area_a, b, area_b, g, &unsure);
    if (!unsure) return yes;
    return iou_plus1(a, area_a, b, area_b) > g.thr;          // inside the band / degenerate: the reference's own test
}

__global__ void __launch_bounds__(kRunWarps * 32)
post_nms_runs_kernel(int capN, int segs_per_image, int num_images, float thr, const int* __restrict__ seg_start,
                     const int* __restrict__ n_seg, const float4* __restrict__ s_box,
                     unsigned char* __restrict__ keep_sorted, unsigned char* __restrict__ row_long,
                     int* __restrict__ big) {
    PAA_TRACE_SCOPE(16);
    __shared__ float4 s_kb[kRunWarps][kRunMax + 4];   // kept boxes of the run so far (+ padding: see below)
    __shared__ float s_ka[kRunWarps][kRunMax + 4];    // their areas
    __shared__ float4 s_gb[kRunWarps][32];            // the current group
    __shared__ float s_ga[kRunWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * kRunWarps + warp;
    const int n = gw / segs_per_image;
    if (n >= num_images) return;
    const int s = gw - n * segs_per_image;
    if (s >= n_seg[n]) return;
    const int* ss = seg_start + (size_t)n * (capN + 1);
    const int a = ss[s], b = ss[s + 1];
    if (b <= a) return;
    const size_t base = (size_t)n * capN;
    if (b - a > kRunMax) {                             // a long run: its rows go through the mask + scan pair
        for (int r = a + lane; r < b; r += 32) row_long[base + r] = 1;
        if (lane == 0) big[n] = 1;
        return;
    }
    const IouGate gate = iou_gate(thr);
    float4* kb = s_kb[warp];
    float* ka = s_ka[warp];
    float4* gb = s_gb[warp];
    float* ga = s_ga[warp];
    // The kept list is read four entries at a time; the up to three entries past its end hold a box that overlaps
    // nothing (and is never "unsure"): no tail loop.
    const float4 nowhere = make_float4(3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f);
    if (lane < 4) {
        kb[lane] = nowhere;
        ka[lane] = 1.0f;
    }
    int nk = 0;                                        // kept so far (warp-uniform)
    for (int g0 = a; g0 < b; g0 += 32) {
        const int r = g0 + lane;
        const bool valid = r < b;
        const float4 box = valid ? s_box[base + r] : nowhere;
        const float area = valid ? area_plus1(box) : 1.0f;
        // (1) suppressed by a box kept earlier in the run?  Four independent tests per trip.
        bool dead = !valid;
        for (int i = 0; i < nk; i += 4) {
            bool u0, u1, u2, u3;
            const bool y0 = iou_certain(kb[i], ka[i], box, area, gate, &u0);
            const bool y1 = iou_certain(kb[i + 1], ka[i + 1], box, area, gate, &u1);
            const bool y2 = iou_certain(kb[i + 2], ka[i + 2], box, area, gate, &u2);
            const bool y3 = iou_certain(kb[i + 3], ka[i + 3], box, area, gate, &u3);
            dead = dead || y0 || y1 || y2 || y3;
            if (u0 || u1 || u2 || u3) {                // rare: a pair inside the rounding band of the threshold
                if (u0) dead = dead || (iou_plus1(kb[i], ka[i], box, area) > thr);
                if (u1) dead = dead || (iou_plus1(kb[i + 1], ka[i + 1], box, area) > thr);
                if (u2) dead = dead || (iou_plus1(kb[i + 2], ka[i + 2], box, area) > thr);
                if (u3) dead = dead || (iou_plus1(kb[i + 3], ka[i + 3], box, area) > thr);
            }
        }
        // The boxes that are still alive are packed to the front (order kept: best score first): what follows works
        // on n_a <= 32 of them, usually about half a group.
        const unsigned alive = ~__ballot_sync(PAA_FULL, dead);
        const int n_a = __popc(alive);
        const int mine_c = __popc(alive & ((1u << lane) - 1u));           // my packed index if I am alive
        if (!dead) {
            gb[mine_c] = box;
            ga[mine_c] = area;
        }
        __syncwarp();
        const bool act = lane < n_a;
        const float4 cbox = act ? gb[lane] : nowhere;
        const float carea = act ? ga[lane] : 1.0f;
        // (2) the pairs among them, each evaluated once: in trip t lane c takes the pair (c, (c + t) mod n_a); trips
        // 1 .. (n_a - 1) / 2 cover every pair whose cyclic distance is below n_a / 2 exactly once, and for even n_a the
        // trip n_a / 2 (lanes below n_a / 2) the rest.  Bit j of lane i's word <=> packed box j > i overlaps packed
        // box i by more than thr; the lane that evaluated a pair whose smaller index is not its own hands the result
        // over through the trip's ballot.
        unsigned m = 0u;
        const int trips = n_a >> 1;
#pragma unroll 2
        for (int t = 1; t <= trips; ++t) {
            int d = lane + t;
            d = d >= n_a ? d - n_a : d;
            const bool last_half = (2 * t == n_a);                            // even n_a: every pair of this trip twice
            const bool on = act && (!last_half || lane < t);
            bool unsure;
            bool hit = iou_certain(cbox, carea, gb[on ? d : 0], ga[on ? d : 0], gate, &unsure);
            if (on && unsure) hit = iou_plus1(cbox, carea, gb[d], ga[d]) > thr;
            hit = hit && on;
            const unsigned hits = __ballot_sync(PAA_FULL, hit);
            if (hit && d > lane) m |= 1u << d;
            int from = lane - t;                       // the lane whose pair of this trip has me as its second box
            from = from < 0 ? from + n_a : from;
            if (act && from > lane && ((hits >> from) & 1u)) m |= 1u << from;
        }
        // (3) greedy scan in packed order, best score first: the words are fetched up front (independent shuffles),
        // what is left is a chain of bit operations
        unsigned removed = 0u, kept_c = 0u;
#pragma unroll 1
        for (int j0 = 0; j0 < n_a; j0 += 8) {          // eight at a time: no trips past the last packed box
            unsigned mj[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) mj[u] = __shfl_sync(PAA_FULL, m, j0 + u);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + u;
                const bool free_j = j < n_a && ((removed >> j) & 1u) == 0u;
                kept_c |= free_j ? (1u << j) : 0u;
                removed |= free_j ? mj[u] : 0u;
            }
        }
        // (4) results; the kept boxes join the list the later groups are tested against
        const bool mine = !dead && ((kept_c >> mine_c) & 1u);
        if (valid) {
            keep_sorted[base + r] = mine ? 1 : 0;
            row_long[base + r] = 0;
        }
        if (act && ((kept_c >> lane) & 1u)) {                                 // packed lane -> list, order kept
            const int at = nk + __popc(kept_c & ((1u << lane) - 1u));
            kb[at] = cbox;
            ka[at] = carea;
        }
        nk += __popc(kept_c);
        if (lane < 4) {                                // padding behind the new end (nk <= kRunMax)
            kb[nk + lane] = nowhere;
            ka[nk + lane] = 1.0f;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// nms_mask: bit (i, j) set when sorted box j > i has the same label and IoU(+1) > thr
// (csrc/cuda/ml_nms.cu:13-24,55-70).  Tiles whose row and column label ranges are disjoint are
// skipped and never read by the scan.
// ---------------------------------------------------------------------------------------------
// One WARP per (image, 32 consecutive sorted rows, column piece p of kMaskPieces): it walks the 64-column tiles
// tile(first row) + p, + p + kMaskPieces, ... for as long as their first label does not exceed the warp's last label
// (boxes are sorted by label).  A tile's boxes are staged in the warp's own slice of shared memory (no block barrier);
// of its 64 columns only the contiguous range whose labels occur among the warp's rows is visited.  Splitting the
// walk into pieces keeps a long run (a class with a thousand boxes: 16 tiles) off the critical path of small batches.
constexpr int kMaskWarps = 4;                        // row groups (of 32 rows) per block
constexpr int kMaskPieces = 4;                       // column pieces per row group: a block is kMaskWarps * kMaskPieces warps

__global__ void __launch_bounds__(kMaskWarps * kMaskPieces * 32)
post_nms_mask_kernel(int capN, int nbw, float thr, const int* __restrict__ total,
                     const float4* __restrict__ s_box, const int* __restrict__ s_label,
                     unsigned long long* __restrict__ mask, const unsigned char* __restrict__ row_long,
                     const int* __restrict__ big) {
    PAA_TRACE_SCOPE(20);
    // only the rows of label runs too long for post_nms_runs_kernel are left (it flagged them and their image)
    if (!big[blockIdx.z]) return;
    __shared__ float4 s_cb[kMaskWarps * kMaskPieces][64];
    __shared__ float s_ca[kMaskWarps * kMaskPieces][64];
    __shared__ int s_cl[kMaskWarps * kMaskPieces][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int piece = warp % kMaskPieces, group = warp / kMaskPieces;
    const int n = blockIdx.z;
    const int cnt = total[n];
    const int r0 = (blockIdx.x * kMaskWarps + group) * 32;
    if (r0 >= cnt) return;
    const size_t base = (size_t)n * capN;
    const int r = r0 + lane;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    int al = -1;
    bool mine = false;
    if (r < cnt) {
        a = s_box[base + r];
        al = s_label[base + r];
        mine = row_long[base + r] != 0;
    }
    const unsigned lm = __ballot_sync(PAA_FULL, mine);
    if (lm == 0u) return;
    const float aa = area_plus1(a);
    const IouGate gate = iou_gate(thr);
    const int first_label = __shfl_sync(PAA_FULL, al, __ffs(lm) - 1);
    const int last_label = __shfl_sync(PAA_FULL, al, 31 - __clz(lm));
    float4* cbx = s_cb[warp];
    float* cba = s_ca[warp];
    int* cbl = s_cl[warp];
    for (int cb = (r0 >> 6) + piece; cb * 64 < cnt; cb += kMaskPieces) {
        const int c0 = cb * 64;
        if (s_label[base + c0] > last_label) break;               // no shared label from here on
        const int csize = min(64, cnt - c0);
        __syncwarp();
        int l0 = INT_MAX, l1 = INT_MAX;                           // past-the-end columns sort after every label
        if (lane < csize) {
            const float4 b = s_box[base + c0 + lane];
            cbx[lane] = b;
            cba[lane] = area_plus1(b);
            l0 = s_label[base + c0 + lane];
            cbl[lane] = l0;
        }
        if (lane + 32 < csize) {
            const float4 b = s_box[base + c0 + 32 + lane];
            cbx[lane + 32] = b;
            cba[lane + 32] = area_plus1(b);
            l1 = s_label[base + c0 + 32 + lane];
            cbl[lane + 32] = l1;
        }
        __syncwarp();
        // columns whose label lies in [first_label, last_label]: one contiguous range of the sorted tile
        const unsigned long long ge = (unsigned long long)__ballot_sync(PAA_FULL, l0 >= first_label) |
                                      ((unsigned long long)__ballot_sync(PAA_FULL, l1 >= first_label) << 32);
        const unsigned long long le = (unsigned long long)__ballot_sync(PAA_FULL, l0 <= last_label) |
                                      ((unsigned long long)__ballot_sync(PAA_FULL, l1 <= last_label) << 32);
        const unsigned long long live = ge & le;
        unsigned long long bits = 0ull;
        if (live && mine) {
            int j_lo = __ffsll((long long)live) - 1;
            const int j_hi = 64 - __clzll((long long)live);
            if (c0 + j_lo <= r) j_lo = r + 1 - c0;                // nothing at or left of the row itself
            for (int j = j_lo; j < j_hi; ++j) {
                if (cbl[j] != al) continue;
                if (iou_exceeds(a, aa, cbx[j], cba[j], gate)) bits |= 1ull << j;
            }
        }
        if (mine) mask[(base + r) * nbw + cb] = bits;
    }
}

static inline dim3 nms_mask_grid(int capN, int num_images) {
    return dim3((capN + kMaskWarps * 32 - 1) / (kMaskWarps * 32), 1, num_images);
}

// ---------------------------------------------------------------------------------------------
// nms_scan: greedy suppression inside one label run, boxes visited in descending score
// (csrc/cuda/ml_nms.cu:116-128).  One warp per run, one 64-box word at a time: the word's incoming
// removed-bits are OR-ed from the rows kept in earlier words of the run, the 64 x 64 diagonal block is
// held in registers (two rows per lane) and resolved with shuffles -- no memory access in the serial part.
// ---------------------------------------------------------------------------------------------
constexpr int kScanWarps = 4;

__device__ __forceinline__ unsigned long long warp_or64(unsigned long long v) {
    const unsigned lo = __reduce_or_sync(PAA_FULL, (unsigned)v);
    const unsigned hi = __reduce_or_sync(PAA_FULL, (unsigned)(v >> 32));
    return ((unsigned long long)hi << 32) | lo;
}

__global__ void __launch_bounds__(kScanWarps * 32)
post_nms_scan_kernel(int capN, int nbw, int segs_per_image, int num_images,
                     const int* __restrict__ seg_start,
                     const int* __restrict__ n_seg, const unsigned long long* __restrict__ mask,
                     unsigned char* __restrict__ keep_sorted, const int* __restrict__ big) {
    PAA_TRACE_SCOPE(17);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * kScanWarps + warp;
    const int n = gw / segs_per_image;
    if (n >= num_images) return;
    if (!big[n]) return;
    const int s = gw - n * segs_per_image;
    if (s >= n_seg[n]) return;
    const int* ss = seg_start + (size_t)n * (capN + 1);
    const int a = ss[s], b = ss[s + 1];
    if (b - a <= kRunMax) return;                       // short runs were finished by post_nms_runs_kernel
    const size_t base = (size_t)n * capN;
    const int w_lo = a >> 6, w_hi = (b - 1) >> 6;
    for (int w = w_lo; w <= w_hi; ++w) {
        const int lo_i = max(a, w * 64), hi_i = min(b, w * 64 + 64);
        // removed bits contributed by the rows kept in earlier words of this run
        unsigned long long R = 0ull;
        for (int i = a + lane; i < lo_i; i += 32)
            if (keep_sorted[base + i]) R |= mask[(base + i) * nbw + w];
        R = warp_or64(R);
        const int r0 = w * 64 + lane, r1 = r0 + 32;
        const unsigned long long d0 = (r0 >= lo_i && r0 < hi_i) ? mask[(base + r0) * nbw + w] : 0ull;
        const unsigned long long d1 = (r1 >= lo_i && r1 < hi_i) ? mask[(base + r1) * nbw + w] : 0ull;
        unsigned long long K = 0ull;
        for (int j = lo_i - w * 64; j < hi_i - w * 64; ++j) {
            if ((R >> j) & 1ull) continue;                      // warp-uniform
            K |= 1ull << j;
            R |= (j < 32) ? __shfl_sync(PAA_FULL, d0, j) : __shfl_sync(PAA_FULL, d1, j - 32);
        }
        if (r0 >= lo_i && r0 < hi_i) keep_sorted[base + r0] = (unsigned char)((K >> lane) & 1ull);
        if (r1 >= lo_i && r1 < hi_i) keep_sorted[base + r1] = (unsigned char)((K >> (lane + 32)) & 1ull);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// finish: top-D cut (inference.py:114-122), compaction in ascending pre-NMS position
// (csrc/cuda/ml_nms.cu:132-135 + boxlist[keep]), output rows.  One block per image.
// ---------------------------------------------------------------------------------------------
constexpr int kFinishThreads = 1024;
constexpr int kFinishCached = 8;          // sorted rows a thread keeps in registers (covers capN <= 8192)
constexpr int kFinishTaken = 1024;        // output rows the fast path can order in shared memory

// The usual call keeps D = 100 rows of a few thousand: the cut is a radix select over the survivors' scores in SORTED
// order (coalesced reads, the first rows of every thread cached in registers), and only the rows that pass it are
// ordered by pre-NMS position (rank by counting among <= 1024 of them).  Calls that keep more rows than that (no cut,
// or a tie at the D-th score wider than the buffer) take the general path below: survivor flags by position and an
// ordered compaction over all positions.
__global__ void __launch_bounds__(kFinishThreads)
post_finish_kernel(int L, int topn, int det_per_img, int skip_nms, const int* __restrict__ pre_cnt,
                   const int* __restrict__ total, const float4* __restrict__ pre_box,
                   const float* __restrict__ pre_score, const int* __restrict__ pre_label,
                   const int* __restrict__ s_pos, const float* __restrict__ s_score,
                   const unsigned char* __restrict__ keep_sorted,
                   unsigned char* __restrict__ flag_by_pos, int* __restrict__ rank_by_pos,
                   float* __restrict__ out_boxes, float* __restrict__ out_scores,
                   long long* __restrict__ out_labels, int* __restrict__ out_count, int* __restrict__ out_rank,
                   unsigned char* __restrict__ dbg_keep) {
    PAA_TRACE_SCOPE(18);
    __shared__ int s_hist[256];
    __shared__ int s_warp[32];
    __shared__ int s_running, s_need;
    __shared__ unsigned s_prefix;
    __shared__ int s_tp[kFinishTaken], s_tr[kFinishTaken];
    __shared__ int s_taken;
    const int n = blockIdx.x;
    const int capN = L * topn;
    const size_t base = (size_t)n * capN;
    const int cnt = total[n];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* flag = flag_by_pos + base;
    int* rpos = rank_by_pos + base;

    // ---- fast path ---------------------------------------------------------------------------------
    // key of sorted row r: the score's bit pattern (scores are > 0) if the row survived, else 0
    auto row_key = [&](int r) -> unsigned {
        const bool k = skip_nms ? true : keep_sorted[base + r] != 0;
        return k ? __float_as_uint(s_score[base + r]) : 0u;
    };
    unsigned ck[kFinishCached];
    int kept_fast = 0;
#pragma unroll
    for (int c = 0; c < kFinishCached; ++c) {
        const int r = threadIdx.x + c * kFinishThreads;
        ck[c] = r < cnt ? row_key(r) : 0u;
        kept_fast += ck[c] != 0u;
    }
    for (int r = threadIdx.x + kFinishCached * kFinishThreads; r < cnt; r += kFinishThreads) kept_fast += row_key(r) != 0u;
    if (threadIdx.x == 0) s_taken = 0;
    int wk = kept_fast;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wk += __shfl_xor_sync(PAA_FULL, wk, o);
    if (lane == 0) s_warp[warp] = wk;
    __syncthreads();
    int n_kept_f = 0;
    for (int w = 0; w < kFinishThreads / 32; ++w) n_kept_f += s_warp[w];
    __syncthreads();
    unsigned cut = 1u;                                        // every survivor (keys of survivors are >= 1)
    if (!skip_nms && det_per_img > 0 && n_kept_f > det_per_img) {
        if (threadIdx.x == 0) {
            s_prefix = 0u;
            s_need = det_per_img;
        }
        __syncthreads();
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            for (int b = threadIdx.x; b < 256; b += kFinishThreads) s_hist[b] = 0;
            __syncthreads();
            const unsigned prefix = s_prefix;
            const unsigned hi_mask = (pass == 0) ? 0u : (~0u << (shift + 8));
#pragma unroll
            for (int c = 0; c < kFinishCached; ++c)
                if (ck[c] != 0u && (ck[c] & hi_mask) == prefix) atomicAdd(&s_hist[(ck[c] >> shift) & 0xff], 1);
            for (int r = threadIdx.x + kFinishCached * kFinishThreads; r < cnt; r += kFinishThreads) {
                const unsigned key = row_key(r);
                if (key != 0u && (key & hi_mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 0xff], 1);
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                int d, remaining;
                radix_pick_digit(s_hist, s_need, threadIdx.x, &d, &remaining);
                if (threadIdx.x == 0) {
                    s_need = remaining;
                    s_prefix = prefix | ((unsigned)d << shift);
                }
            }
            __syncthreads();
        }
        cut = s_prefix;            // bit pattern of the D-th largest surviving score (ties keep more, inference.py:118-121)
    }
    // the rows that pass the cut, in arbitrary order; their number decides the path
    int mine = 0;
#pragma unroll
    for (int c = 0; c < kFinishCached; ++c) mine += (ck[c] != 0u && ck[c] >= cut) ? 1 : 0;
    for (int r = threadIdx.x + kFinishCached * kFinishThreads; r < cnt; r += kFinishThreads) {
        const unsigned key = row_key(r);
        mine += (key != 0u && key >= cut) ? 1 : 0;
    }
    int wm = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wm += __shfl_xor_sync(PAA_FULL, wm, o);
    if (lane == 0) s_warp[warp] = wm;
    __syncthreads();
    int n_taken = 0;
    for (int w = 0; w < kFinishThreads / 32; ++w) n_taken += s_warp[w];
    __syncthreads();
    if (n_taken <= kFinishTaken && dbg_keep == nullptr) {
#pragma unroll
        for (int c = 0; c < kFinishCached; ++c)
            if (ck[c] != 0u && ck[c] >= cut) {
                const int r = threadIdx.x + c * kFinishThreads;
                const int slot = atomicAdd(&s_taken, 1);
                s_tr[slot] = r;
                s_tp[slot] = s_pos[base + r];
            }
        for (int r = threadIdx.x + kFinishCached * kFinishThreads; r < cnt; r += kFinishThreads) {
            const unsigned key = row_key(r);
            if (key != 0u && key >= cut) {
                const int slot = atomicAdd(&s_taken, 1);
                s_tr[slot] = r;
                s_tp[slot] = s_pos[base + r];
            }
        }
        __syncthreads();
        if (threadIdx.x < n_taken) {
            const int p = s_tp[threadIdx.x];
            int row = 0;
            for (int j = 0; j < n_taken; ++j) row += s_tp[j] < p ? 1 : 0;      // positions are unique
            const float4 b = pre_box[base + p];
            float* ob = out_boxes + (base + row) * 4;
            ob[0] = b.x;
            ob[1] = b.y;
            ob[2] = b.z;
            ob[3] = b.w;
            out_scores[base + row] = pre_score[base + p];
            out_labels[base + row] = (long long)pre_label[base + p];
            out_rank[base + row] = s_tr[threadIdx.x];
        }
        if (threadIdx.x == 0) out_count[n] = n_taken;
        return;
    }

    // ---- general path ------------------------------------------------------------------------------
    // survivors by position
    for (int p = threadIdx.x; p < capN; p += kFinishThreads) flag[p] = 0;
    __syncthreads();
    int kept_local = 0;
    for (int r = threadIdx.x; r < cnt; r += kFinishThreads) {
        const int p = s_pos[base + r];
        const unsigned char k = skip_nms ? 1 : keep_sorted[base + r];
        flag[p] = k;
        rpos[p] = r;
        kept_local += k;
    }
    __syncthreads();
    // block sum of kept_local
    int wsum = kept_local;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(PAA_FULL, wsum, o);
    if (lane == 0) s_warp[warp] = wsum;
    __syncthreads();
    int n_kept = 0;
    for (int w = 0; w < kFinishThreads / 32; ++w) n_kept += s_warp[w];
    __syncthreads();
    if (dbg_keep) {
        for (int l = 0, o = 0; l < L; ++l) {        // debug output is compact level-major
            const int c = pre_cnt[n * L + l];
            for (int j = threadIdx.x; j < c; j += kFinishThreads) dbg_keep[base + o + j] = flag[l * topn + j];
            o += c;
        }
    }

    // more than D survivors: keep those whose score is >= the D-th largest (ties keep more)
    unsigned thr_bits = 0u;
    if (!skip_nms && det_per_img > 0 && n_kept > det_per_img) {
        if (threadIdx.x == 0) {
            s_prefix = 0u;
            s_need = det_per_img;
        }
        __syncthreads();
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            for (int b = threadIdx.x; b < 256; b += kFinishThreads) s_hist[b] = 0;
            __syncthreads();
            const unsigned prefix = s_prefix;
            const unsigned hi_mask = (pass == 0) ? 0u : (~0u << (shift + 8));
            for (int p = threadIdx.x; p < capN; p += kFinishThreads) {
                if (!flag[p]) continue;
                const unsigned key = __float_as_uint(pre_score[base + p]);
                if ((key & hi_mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 0xff], 1);
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                int d, remaining;
                radix_pick_digit(s_hist, s_need, threadIdx.x, &d, &remaining);
                if (threadIdx.x == 0) {
                    s_need = remaining;
                    s_prefix = prefix | ((unsigned)d << shift);
                }
            }
            __syncthreads();
        }
        thr_bits = s_prefix;       // bit pattern of the D-th largest surviving score
    }

    // ordered compaction over positions
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    for (int p0 = 0; p0 < capN; p0 += kFinishThreads) {
        const int p = p0 + threadIdx.x;
        bool take = false;
        if (p < capN && flag[p]) take = __float_as_uint(pre_score[base + p]) >= thr_bits;
        const unsigned m = __ballot_sync(PAA_FULL, take);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = s_running;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (take) {
            const int row = before + __popc(m & ((1u << lane) - 1u));
            const float4 b = pre_box[base + p];
            float* ob = out_boxes + (base + row) * 4;
            ob[0] = b.x;
            ob[1] = b.y;
            ob[2] = b.z;
            ob[3] = b.w;
            out_scores[base + row] = pre_score[base + p];
            out_labels[base + row] = (long long)pre_label[base + p];
            out_rank[base + row] = rpos[p];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < kFinishThreads / 32; ++w) t += s_warp[w];
            s_running += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out_count[n] = s_running;
}

// ---------------------------------------------------------------------------------------------
// vote: every output box becomes the weighted mean of all pre-NMS boxes of its class with
// IoU(+1) > 0.01, weight exp(-(1-IoU)^2 / 0.025) * score (inference.py:123-157).  One warp per row.
// ---------------------------------------------------------------------------------------------
constexpr int kVoteWarps = 4;
constexpr int kVoteBlocksPerImage = 32;

__global__ void __launch_bounds__(kVoteWarps * 32)
post_vote_kernel(int capN, const int* __restrict__ out_count, const int* __restrict__ out_rank,
                 const int* __restrict__ seg_start, const int* __restrict__ n_seg,
                 const float4* __restrict__ s_box, const float* __restrict__ s_score,
                 float* __restrict__ out_boxes) {
    PAA_TRACE_SCOPE(19);
    const int n = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * kVoteWarps + (threadIdx.x >> 5);
    const size_t base = (size_t)n * capN;
    const int rows = out_count[n];
    const int* ss = seg_start + (size_t)n * (capN + 1);
    const int ns = n_seg[n];
    for (int row = gw; row < rows; row += kVoteBlocksPerImage * kVoteWarps) {
        const int r = out_rank[base + row];
        // run containing sorted index r: last start <= r
        int lo = 0, hi = ns - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (ss[mid] <= r) lo = mid; else hi = mid - 1;
        }
        const int a = ss[lo], b = ss[lo + 1];
        const float4 det = s_box[base + r];
        const float det_area = area_plus1(det);
        float sx1 = 0.f, sy1 = 0.f, sx2 = 0.f, sy2 = 0.f, sp = 0.f;
        for (int j = a + lane; j < b; j += 32) {
            const float4 c = s_box[base + j];
            const float q = iou_plus1(det, det_area, c, area_plus1(c));
            if (q > 0.01f) {
                const float t = __fsub_rn(1.0f, q);
                const float p = __fmul_rn(expf(__fdiv_rn(-__fmul_rn(t, t), 0.025f)), s_score[base + j]);
                sx1 = fmaf(c.x, p, sx1);
                sy1 = fmaf(c.y, p, sy1);
                sx2 = fmaf(c.z, p, sx2);
                sy2 = fmaf(c.w, p, sy2);
                sp += p;
            }
        }
        sx1 = warp_sum(sx1);
        sy1 = warp_sum(sy1);
        sx2 = warp_sum(sx2);
        sy2 = warp_sum(sy2);
        sp = warp_sum(sp);
        if (lane == 0) {
            float* ob = out_boxes + (base + row) * 4;
            ob[0] = sx1 / sp;
            ob[1] = sy1 / sp;
            ob[2] = sx2 / sp;
            ob[3] = sy2 / sp;
        }
    }
}

// compact copies of the strided pre-NMS arrays for the debug outputs
__global__ void post_debug_pre_kernel(int L, int topn, const int* __restrict__ pre_cnt,
                                      const float4* __restrict__ pre_box, const float* __restrict__ pre_score,
                                      const int* __restrict__ pre_label, float* __restrict__ d_box,
                                      float* __restrict__ d_score, int* __restrict__ d_label,
                                      int* __restrict__ d_count) {
    const int n = blockIdx.x;
    const size_t base = (size_t)n * L * topn;
    int o = 0;
    for (int l = 0; l < L; ++l) {
        const int c = pre_cnt[n * L + l];
        for (int j = threadIdx.x; j < c; j += blockDim.x) {
            const size_t src = base + (size_t)l * topn + j, dst = base + o + j;
            if (d_box) {
                const float4 b = pre_box[src];
                d_box[dst * 4 + 0] = b.x;
                d_box[dst * 4 + 1] = b.y;
                d_box[dst * 4 + 2] = b.z;
                d_box[dst * 4 + 3] = b.w;
            }
            if (d_score) d_score[dst] = pre_score[src];
            if (d_label) d_label[dst] = pre_label[src];
        }
        if (d_count && threadIdx.x == 0) d_count[n * L + l] = c;
        o += c;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#ifdef PAA_TRACE
PAA_TRACE_SETTER(trace_set_post)
#endif

int run_postprocess(const Geometry& geo, const PaaPostArgs* a, cudaStream_t stream) {
    const int N = a->num_images, L = a->num_levels, topn = a->pre_nms_top_n, C = a->num_classes;
    if (topn < 1 || topn > kMaxTopN) {
        set_error("pre_nms_top_n=%d outside [1, %d]", topn, kMaxTopN);
        return PAA_ERR_UNSUPPORTED;
    }
    const int capN = L * topn;
    if (capN > 65535) {
        set_error("num_levels*pre_nms_top_n=%d exceeds 65535", capN);
        return PAA_ERR_UNSUPPORTED;
    }
    if ((size_t)geo.A * C >= (1ull << 31)) {
        set_error("anchors*classes per image too large");
        return PAA_ERR_UNSUPPORTED;
    }
    if (!a->workspace || !a->out_boxes || !a->out_scores || !a->out_labels || !a->out_count) {
        set_error("null workspace / output pointer");
        return PAA_ERR_BAD_ARGUMENT;
    }
    if (!(a->pre_nms_thresh > 0.0f && a->pre_nms_thresh < 1.0f)) {
        set_error("pre_nms_thresh=%g outside (0, 1)", (double)a->pre_nms_thresh);
        return PAA_ERR_BAD_ARGUMENT;
    }
    PostWorkspace w = carve_post(a->workspace, N, geo.A, C, L, topn);
    if (w.total_bytes > a->workspace_bytes) {
        set_error("workspace too small: need %zu bytes, got %zu", w.total_bytes, a->workspace_bytes);
        return PAA_ERR_WORKSPACE;
    }
    ImageSizes sizes;
    for (int i = 0; i < N; ++i) {
        sizes.wh[i][0] = a->image_wh[i][0];
        sizes.wh[i][1] = a->image_wh[i][1];
    }
    const int nbw = (capN + 63) / 64;
    DecodeSpec dec;
    dec.mode = a->box_decode;
    dec.wx = a->decode_weights[0];
    dec.wy = a->decode_weights[1];
    dec.ww = a->decode_weights[2];
    dec.wh = a->decode_weights[3];
    dec.clip = a->decode_clip;
    if (dec.mode < PAA_DECODE_ATSS_BOX || dec.mode > PAA_DECODE_LTRB ||
        (dec.mode == PAA_DECODE_LEGACY && !(dec.wx > 0.0f && dec.wy > 0.0f && dec.ww > 0.0f && dec.wh > 0.0f))) {
        set_error("box_decode=%d / decode_weights unsupported", dec.mode);
        return PAA_ERR_BAD_ARGUMENT;
    }
    PAA_CUDA_CHECK(cudaMemsetAsync(a->workspace, 0, w.zero_bytes, stream));
    // a logit can only pass sigmoid(x) > thr if x > logit(thr); gate a little below that and test exactly
    const float gate = logf(a->pre_nms_thresh / (1.0f - a->pre_nms_thresh)) - 1e-3f;
    {
        CandPlan plan;
        unsigned chunks = 0;
        for (int l = 0; l < PAA_MAX_LEVELS; ++l) {
            plan.chunk_off[l] = chunks;
            plan.chunks_per_image[l] = 1;
            plan.inv_cpi[l] = 1.0f;
            plan.vec[l] = 0;
            if (l >= L) continue;
            const LevelView& lv = geo.lv[l];
            const unsigned long long per_image = (unsigned long long)geo.apl * C * lv.hw;
            plan.chunks_per_image[l] = (unsigned)((per_image + kCandChunk - 1) / kCandChunk);
            plan.inv_cpi[l] = 1.0f / (float)plan.chunks_per_image[l];
            plan.vec[l] = ((per_image & 3ull) == 0 && (reinterpret_cast<uintptr_t>(lv.cls) & 15u) == 0) ? 1 : 0;
            chunks += plan.chunks_per_image[l] * (unsigned)N;
        }
        plan.chunk_off[PAA_MAX_LEVELS] = chunks;
        plan.total = chunks;
        // opt-in (measurement switch): see the kernel's comment for why the register-staged kernel stays the default
        bool ring = getenv("PAA_POST_RING") != nullptr;
        for (int l = 0; l < L; ++l) ring = ring && plan.vec[l] != 0;
        KernelTimer t(PAA_KERNEL_POST_CANDIDATES, stream);
        if (ring) {
            RingPlan rp;
            unsigned pieces = 0;
            for (int l = 0; l < PAA_MAX_LEVELS; ++l) {
                rp.piece_off[l] = pieces;
                rp.pieces_per_image[l] = 1;
                rp.inv_ppi[l] = 1.0f;
                if (l >= L) continue;
                const unsigned long long per_image = (unsigned long long)geo.apl * C * geo.lv[l].hw;
                rp.pieces_per_image[l] = (unsigned)((per_image + kRingStageElems - 1) / kRingStageElems);
                rp.inv_ppi[l] = 1.0f / (float)rp.pieces_per_image[l];
                pieces += rp.pieces_per_image[l] * (unsigned)N;
            }
            rp.piece_off[PAA_MAX_LEVELS] = pieces;
            rp.total = pieces;
            static bool attr_set = false;
            if (!attr_set) {
                PAA_CUDA_CHECK(cudaFuncSetAttribute(post_candidates_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    kRingStages * kRingStageBytes));
                attr_set = true;
            }
            unsigned grid = 148u;
            if (grid > pieces) grid = pieces;
            post_candidates_ring_kernel<<<grid, kCandThreads, kRingStages * kRingStageBytes, stream>>>(
                geo, rp, a->pre_nms_thresh, gate, w.cand, w.cand_count, w.hist);
        } else {
            unsigned grid = 148u * kCandBlocksPerSM;
            if (const char* e = getenv("PAA_CAND_BLOCKS")) grid = 148u * (unsigned)atoi(e);     // measurement switch
            if (grid > chunks) grid = chunks;
            if (grid < 1) grid = 1;
            post_candidates_kernel<<<grid, kCandThreads, 0, stream>>>(geo, plan, a->pre_nms_thresh, gate,
                                                                      w.cand, w.cand_count, w.hist);
        }
    }
    PAA_LAUNCH_CHECK("post_candidates_kernel");
    {
        KernelTimer t(PAA_KERNEL_POST_THRESHOLD, stream);
        post_threshold_kernel<<<N * L, 256, 0, stream>>>(w.cand_count, w.hist, topn, w.thr_bin, w.n_above, w.k_sel);
    }
    PAA_LAUNCH_CHECK("post_threshold_kernel");
    {
        KernelTimer t(PAA_KERNEL_POST_FILTER, stream);
        post_filter_kernel<<<N * L * kFilterBlocks, 256, 0, stream>>>(geo, w.cand, w.cand_count, w.thr_bin, topn,
                                                                      w.sel, w.sel_count, w.bnd, w.bnd_count);
    }
    PAA_LAUNCH_CHECK("post_filter_kernel");
    {
        KernelTimer t(PAA_KERNEL_POST_SELECT, stream);
        const int ept = topn <= 4 * kSelectThreads ? 4 : (topn <= 8 * kSelectThreads ? 8 : 16);
        const size_t smem = select_smem_bytes(ept);
#define PAA_SELECT(E)                                                                                         \
    do {                                                                                                      \
        static bool attr_set = false;                                                                         \
        if (!attr_set && smem > 48 * 1024) {                                                                  \
            PAA_CUDA_CHECK(cudaFuncSetAttribute(post_select_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                (int)select_smem_bytes(E)));                                  \
            attr_set = true;                                                                                  \
        }                                                                                                     \
        post_select_kernel<E><<<N * L, kSelectThreads, smem, stream>>>(geo, sizes, w.cand, w.bnd, w.bnd_count, w.n_above, \
                                                                 w.k_sel, w.thr_bin, topn, 0.0f, dec, w.sel, w.pre_cnt, \
                                                                 w.pre_box, w.pre_score, w.pre_label);                  \
    } while (0)
        if (ept == 4) PAA_SELECT(4);
        else if (ept == 8) PAA_SELECT(8);
        else PAA_SELECT(16);
#undef PAA_SELECT
    }
    PAA_LAUNCH_CHECK("post_select_kernel");
    if (a->dbg_pre_boxes || a->dbg_pre_scores || a->dbg_pre_labels || a->dbg_pre_count) {
        post_debug_pre_kernel<<<N, 256, 0, stream>>>(L, topn, w.pre_cnt, w.pre_box, w.pre_score, w.pre_label,
                                                     a->dbg_pre_boxes, a->dbg_pre_scores, a->dbg_pre_labels,
                                                     a->dbg_pre_count);
        PAA_LAUNCH_CHECK("post_debug_pre_kernel");
    }
    const bool grouped = capN >= C + 1 && (size_t)(2 * (C + 2)) * sizeof(int) <= 40 * 1024;
    if (grouped) {
        // labels are 1..C: group by label, rank inside the label run (O(n * run) instead of O(n^2))
        KernelTimer t(PAA_KERNEL_POST_RANK, stream);
        post_group_kernel<<<N, 1024, 2 * (C + 2) * sizeof(int), stream>>>(L, topn, C, w.pre_cnt, w.pre_score,
                                                                          w.pre_label, w.g_pos, w.g_score,
                                                                          w.seg_start, w.n_seg, w.total);
        dim3 grid((capN + 255) / 256, N);
        post_class_rank_kernel<<<grid, 256, 0, stream>>>(capN, w.total, w.seg_start, w.g_pos, w.g_score, w.pre_box,
                                                         w.pre_label, w.s_box, w.s_score, w.s_label, w.s_pos);
        PAA_LAUNCH_CHECK("post_group_kernel/post_class_rank_kernel");
    } else {
        {
            KernelTimer t(PAA_KERNEL_POST_RANK, stream);
            dim3 grid((capN + kRankThreads - 1) / kRankThreads, N);
            post_rank_kernel<<<grid, kRankThreads, 0, stream>>>(L, topn, w.pre_cnt, w.pre_box, w.pre_score,
                                                                w.pre_label, w.s_box, w.s_score, w.s_label, w.s_pos,
                                                                w.total);
        }
        PAA_LAUNCH_CHECK("post_rank_kernel");
        {
            KernelTimer t(PAA_KERNEL_POST_SEGMENTS, stream);
            post_segments_kernel<<<N, 1024, 0, stream>>>(capN, w.total, w.s_label, w.seg_start, w.n_seg);
        }
        PAA_LAUNCH_CHECK("post_segments_kernel");
    }
    if (!a->skip_nms) {
        const int segs = grouped ? C + 1 : (capN < C ? capN : C);   // label runs per image (upper bound)
        const int warps = N * segs;
        {
            // every label run of up to kRunMax boxes: one warp, no mask in memory
            KernelTimer t(PAA_KERNEL_POST_NMS_RUNS, stream);
            post_nms_runs_kernel<<<(warps + kRunWarps - 1) / kRunWarps, kRunWarps * 32, 0, stream>>>(
                capN, segs, N, a->nms_thresh, w.seg_start, w.n_seg, w.s_box, w.keep_sorted, w.row_long, w.nms_big);
        }
        PAA_LAUNCH_CHECK("post_nms_runs_kernel");
        {
            // images with a longer run (flagged by the kernel above; the others return at once)
            KernelTimer t(PAA_KERNEL_POST_NMS_MASK, stream);
            post_nms_mask_kernel<<<nms_mask_grid(capN, N), kMaskWarps * kMaskPieces * 32, 0, stream>>>(capN, nbw, a->nms_thresh, w.total, w.s_box, w.s_label,
                                                          w.mask, w.row_long, w.nms_big);
        }
        PAA_LAUNCH_CHECK("post_nms_mask_kernel");
        {
            KernelTimer t(PAA_KERNEL_POST_NMS_SCAN, stream);
            post_nms_scan_kernel<<<(warps + kScanWarps - 1) / kScanWarps, kScanWarps * 32, 0, stream>>>(
                capN, nbw, segs, N, w.seg_start, w.n_seg, w.mask, w.keep_sorted, w.nms_big);
        }
        PAA_LAUNCH_CHECK("post_nms_scan_kernel");
    }
    {
        KernelTimer t(PAA_KERNEL_POST_FINISH, stream);
        unsigned char* flag = w.flag_by_pos;
        int* rpos = w.rank_by_pos;
        post_finish_kernel<<<N, kFinishThreads, 0, stream>>>(
            L, topn, a->detections_per_img, a->skip_nms, w.pre_cnt, w.total, w.pre_box, w.pre_score, w.pre_label,
            w.s_pos, w.s_score, w.keep_sorted, flag, rpos, a->out_boxes, a->out_scores,
            reinterpret_cast<long long*>(a->out_labels), a->out_count, w.out_rank, a->dbg_nms_keep);
    }
    PAA_LAUNCH_CHECK("post_finish_kernel");
    if (a->score_voting && !a->skip_nms) {
        KernelTimer t(PAA_KERNEL_POST_VOTE, stream);
        dim3 grid(kVoteBlocksPerImage, N);
        post_vote_kernel<<<grid, kVoteWarps * 32, 0, stream>>>(capN, a->out_count, w.out_rank, w.seg_start, w.n_seg,
                                                               w.s_box, w.s_score, a->out_boxes);
        PAA_LAUNCH_CHECK("post_vote_kernel");
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// stand-alone label-aware NMS behind `_C.ml_nms` (csrc/ml_nms.h:10-27)
// ---------------------------------------------------------------------------------------------
struct MlNmsWorkspace {
    int* cnt;             // [2]  number of boxes; long-run flag of the fused NMS
    float4* box;          // [n]
    int* label;           // [n]
    float4* s_box;
    float* s_score;
    int* s_label;
    int* s_pos;
    int* total;
    int* seg_start;       // [n+1]
    int* n_seg;
    unsigned long long* mask;
    unsigned char* keep_sorted;
    unsigned char* row_long;
    size_t total_bytes;
};

static MlNmsWorkspace carve_ml_nms(void* base, int n) {
    MlNmsWorkspace w;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* q = p ? p + off : nullptr;
        off += (bytes + 255) / 256 * 256;
        return q;
    };
    const size_t nn = n > 0 ? n : 1, nbw = (nn + 63) / 64;
    w.cnt = (int*)take(4);
    w.box = (float4*)take(nn * 16);
    w.label = (int*)take(nn * 4);
    w.s_box = (float4*)take(nn * 16);
    w.s_score = (float*)take(nn * 4);
    w.s_label = (int*)take(nn * 4);
    w.s_pos = (int*)take(nn * 4);
    w.total = (int*)take(4);
    w.seg_start = (int*)take((nn + 1) * 4);
    w.n_seg = (int*)take(4);
    w.mask = (unsigned long long*)take(nn * nbw * 8);
    w.keep_sorted = (unsigned char*)take(nn);
    w.row_long = (unsigned char*)take(nn);
    w.total_bytes = off;
    return w;
}

size_t ml_nms_workspace_bytes(int n) { return carve_ml_nms(nullptr, n).total_bytes; }

__global__ void ml_nms_prepare_kernel(int n, const float* __restrict__ boxes, const float* __restrict__ labels,
                                      float4* __restrict__ box, int* __restrict__ label, int* __restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        cnt[0] = n;
        cnt[1] = 0;                  // "a label run too long for post_nms_runs_kernel" flag of the call
    }
    if (i < n) {
        box[i] = make_float4(boxes[i * 4], boxes[i * 4 + 1], boxes[i * 4 + 2], boxes[i * 4 + 3]);
        label[i] = (int)labels[i];
    }
}

__global__ void ml_nms_scatter_kernel(int n, const int* __restrict__ s_pos, const unsigned char* __restrict__ ks,
                                      unsigned char* __restrict__ keep, int* __restrict__ num_keep) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    int k = 0;
    if (r < n) {
        k = ks[r];
        keep[s_pos[r]] = (unsigned char)k;
    }
    const unsigned m = __ballot_sync(PAA_FULL, k != 0);
    if ((threadIdx.x & 31) == 0 && m && num_keep) atomicAdd(num_keep, __popc(m));
}

int run_ml_nms(const float* boxes, const float* scores, const float* labels, int n, float thresh,
               uint8_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (n < 0 || n > 65535) {
        set_error("paa_ml_nms: n=%d outside [0, 65535]", n);
        return PAA_ERR_UNSUPPORTED;
    }
    if (num_keep) PAA_CUDA_CHECK(cudaMemsetAsync(num_keep, 0, sizeof(int32_t), stream));
    if (n == 0) return 0;
    if (!boxes || !scores || !labels || !keep || !workspace) {
        set_error("paa_ml_nms: null pointer");
        return PAA_ERR_BAD_ARGUMENT;
    }
    MlNmsWorkspace w = carve_ml_nms(workspace, n);
    if (w.total_bytes > workspace_bytes) {
        set_error("paa_ml_nms: workspace too small: need %zu bytes, got %zu", w.total_bytes, workspace_bytes);
        return PAA_ERR_WORKSPACE;
    }
    const int nbw = (n + 63) / 64;
    ml_nms_prepare_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, boxes, labels, w.box, w.label, w.cnt);
    PAA_LAUNCH_CHECK("ml_nms_prepare_kernel");
    dim3 rgrid((n + kRankThreads - 1) / kRankThreads, 1);
    post_rank_kernel<<<rgrid, kRankThreads, 0, stream>>>(1, n, w.cnt, w.box, scores, w.label, w.s_box, w.s_score,
                                                         w.s_label, w.s_pos, w.total);
    PAA_LAUNCH_CHECK("post_rank_kernel");
    post_segments_kernel<<<1, 1024, 0, stream>>>(n, w.total, w.s_label, w.seg_start, w.n_seg);
    PAA_LAUNCH_CHECK("post_segments_kernel");
    post_nms_runs_kernel<<<(n + kRunWarps - 1) / kRunWarps, kRunWarps * 32, 0, stream>>>(
        n, n, 1, thresh, w.seg_start, w.n_seg, w.s_box, w.keep_sorted, w.row_long, w.cnt + 1);
    PAA_LAUNCH_CHECK("post_nms_runs_kernel");
    post_nms_mask_kernel<<<nms_mask_grid(n, 1), kMaskWarps * kMaskPieces * 32, 0, stream>>>(n, nbw, thresh, w.total, w.s_box, w.s_label, w.mask,
                                                                             w.row_long, w.cnt + 1);
    PAA_LAUNCH_CHECK("post_nms_mask_kernel");
    post_nms_scan_kernel<<<(n + kScanWarps - 1) / kScanWarps, kScanWarps * 32, 0, stream>>>(
        n, nbw, n, 1, w.seg_start, w.n_seg, w.mask, w.keep_sorted, w.cnt + 1);
    PAA_LAUNCH_CHECK("post_nms_scan_kernel");
    ml_nms_scatter_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, w.s_pos, w.keep_sorted, keep, num_keep);
    PAA_LAUNCH_CHECK("ml_nms_scatter_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Test-time-augmentation box merging (paa_core/engine/bbox_aug_vote.py:140-310): the
// detections pooled from all scales / flips of one image are merged class by class.
//   mode 0 'nms'        per-class NMS (:190-194), rows best first inside a class
//   mode 1 'vote'       bbox_vote (:198-246)
//   mode 2 'soft-vote'  soft_bbox_vote (:249-310)
// One warp walks one class, exactly as the reference's while-loop does: the best remaining box is the pivot,
// the remaining boxes with IoU(+1) >= vote_thresh to it form a group (found 32 at a time, in order), one box
// alone passes through, several become their score-weighted mean with the pivot's score.  The arithmetic is
// the reference's float32 numpy arithmetic: products box * score, coordinate sums accumulated row by row,
// the score sum with numpy's pairwise rule (sequential below 8 terms, eight interleaved accumulators up to
// 128 terms), so the merged boxes are bit-identical for groups of up to 128 boxes.
// ---------------------------------------------------------------------------------------------
constexpr int kVoteGroupMax = 128;
constexpr int kVoteWarpsPerBlock = 4;

// numpy's pairwise float32 sum of n <= 128 terms (numpy/core/src/umath/loops_utils.h: pairwise_sum)
__device__ __forceinline__ float numpy_sum_f32(const float* a, int n) {
    if (n < 8) {
        float r = 0.0f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
        return r;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
}

__global__ void __launch_bounds__(kVoteWarpsPerBlock * 32)
box_vote_kernel(const int* __restrict__ seg_start, const int* __restrict__ n_seg, const float4* __restrict__ s_box,
                const float* __restrict__ s_score, const unsigned char* __restrict__ keep_sorted, int mode,
                float vote_thresh, float soft_thresh, unsigned char* __restrict__ alive, float4* __restrict__ t_box,
                float* __restrict__ t_score, int* __restrict__ t_rank, int* __restrict__ run_cnt) {
    __shared__ float s_group[kVoteWarpsPerBlock][kVoteGroupMax];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int r = blockIdx.x * kVoteWarpsPerBlock + wib;
    if (r >= *n_seg) return;
    const int lo = seg_start[r], hi = seg_start[r + 1], m = hi - lo;
    const int mult = (mode == 2) ? 2 : 1;
    const int base = mult * lo;
    int cnt = 0;
    if (mode == 0) {
        // NMS survivors, already in descending score
        for (int c = lo; c < hi; c += 32) {
            const int idx = c + lane;
            const bool k = idx < hi && keep_sorted[idx] != 0;
            const unsigned b = __ballot_sync(PAA_FULL, k);
            if (k) {
                const int o = base + cnt + __popc(b & ((1u << lane) - 1u));
                t_box[o] = s_box[idx];
                t_score[o] = s_score[idx];
            }
            cnt += __popc(b);
        }
    } else if (m <= 1) {
        // fewer than two boxes: the class passes through (:220-221, :190-192)
        if (m == 1 && lane == 0) {
            t_box[base] = s_box[lo];
            t_score[base] = s_score[lo];
        }
        cnt = m;
    } else {
        for (int c = lo + lane; c < hi; c += 32) alive[c] = 1;
        __syncwarp();
        int cursor = lo;
        float* gsc = s_group[wib];
        for (;;) {
            // pivot: the first box still alive
            int p = -1;
            for (int c = cursor; c < hi && p < 0; c += 32) {
                const unsigned b = __ballot_sync(PAA_FULL, (c + lane < hi) && alive[c + lane] != 0);
                if (b) p = c + __ffs(b) - 1;
            }
            if (p < 0) break;
            cursor = p + 1;
            const float4 pb = s_box[p];
            const float ps = s_score[p];
            const float area_p = area_plus1(pb);
            float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f, seq = 0.f;
            int k = 0, nd = 0;
            for (int c = p; c < hi; c += 32) {
                const int idx = c + lane;
                const bool al = idx < hi && alive[idx] != 0;
                float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
                float sc = 0.f, o = -1.0f;
                if (al) {
                    bx = s_box[idx];
                    sc = s_score[idx];
                    o = iou_plus1(pb, area_p, bx, area_plus1(bx));
                }
                const bool mg = al && o >= vote_thresh;
                unsigned b = __ballot_sync(PAA_FULL, mg);
                if (mg) alive[idx] = 0;
                while (b) {                 // group members in ascending index, as numpy enumerates them
                    const int src = __ffs(b) - 1;
                    b &= b - 1;
                    const float gx = __shfl_sync(PAA_FULL, bx.x, src), gy = __shfl_sync(PAA_FULL, bx.y, src);
                    const float gz = __shfl_sync(PAA_FULL, bx.z, src), gw = __shfl_sync(PAA_FULL, bx.w, src);
                    const float gs = __shfl_sync(PAA_FULL, sc, src), go = __shfl_sync(PAA_FULL, o, src);
                    sx = __fadd_rn(sx, __fmul_rn(gx, gs));
                    sy = __fadd_rn(sy, __fmul_rn(gy, gs));
                    sz = __fadd_rn(sz, __fmul_rn(gz, gs));
                    sw = __fadd_rn(sw, __fmul_rn(gw, gs));
                    seq = __fadd_rn(seq, gs);
                    if (lane == 0 && k < kVoteGroupMax) gsc[k] = gs;
                    if (mode == 2) {
                        // decayed copy of the member; whether it is kept is only known to matter for groups of
                        // two or more, where it follows the merged row
                        const float dec = __fmul_rn(gs, __fsub_rn(1.0f, go));
                        if (dec >= soft_thresh) {
                            if (lane == 0) {
                                t_box[base + cnt + 1 + nd] = make_float4(gx, gy, gz, gw);
                                t_score[base + cnt + 1 + nd] = dec;
                            }
                            ++nd;
                        }
                    }
                    ++k;
                }
            }
            __syncwarp();
            if (k <= 1) {
                if (lane == 0) {
                    t_box[base + cnt] = pb;
                    t_score[base + cnt] = ps;
                }
                cnt += 1;
            } else {
                float ssum = seq;
                if (k <= kVoteGroupMax) ssum = numpy_sum_f32(gsc, k);     // every lane reads the same values
                if (lane == 0) {
                    t_box[base + cnt] = make_float4(__fdiv_rn(sx, ssum), __fdiv_rn(sy, ssum), __fdiv_rn(sz, ssum),
                                                    __fdiv_rn(sw, ssum));
                    t_score[base + cnt] = ps;                           // the pivot carries the group's maximum
                }
                cnt += 1 + nd;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    // position of every row inside its class: as produced, or by descending score for soft-vote (:303-304)
    for (int i = lane; i < cnt; i += 32) {
        int rank = i;
        if (mode == 2 && m > 1) {
            const float si = t_score[base + i];
            rank = 0;
            for (int q = 0; q < cnt; ++q) {
                const float sq = t_score[base + q];
                rank += (sq > si || (sq == si && q > i)) ? 1 : 0;      // reversed stable order on ties
            }
        }
        t_rank[base + i] = rank;
    }
    if (lane == 0) run_cnt[r] = cnt;
}

// One block: class offsets, the max_detections cut (kthvalue, :163-172; ties keep more), output rows.
__global__ void __launch_bounds__(1024)
box_vote_finish_kernel(const int* __restrict__ seg_start, const int* __restrict__ n_seg,
                       const int* __restrict__ s_label, int mult, const float4* __restrict__ t_box,
                       const float* __restrict__ t_score, const int* __restrict__ t_rank,
                       const int* __restrict__ run_cnt, int max_det, int* __restrict__ run_off,
                       float4* __restrict__ u_box, float* __restrict__ u_score, int* __restrict__ u_label,
                       float* __restrict__ out_boxes, float* __restrict__ out_scores,
                       long long* __restrict__ out_labels, int* __restrict__ out_count) {
    __shared__ int s_hist[256];
    __shared__ int s_warp[32];
    __shared__ int s_running, s_need, s_total;
    __shared__ unsigned s_prefix;
    const int runs = *n_seg;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {                         // a few dozen classes: serial exclusive scan
        int t = 0;
        for (int r = 0; r < runs; ++r) {
            run_off[r] = t;
            t += run_cnt[r];
        }
        s_total = t;
    }
    __syncthreads();
    const int total = s_total;
    // rows of all classes, class-major, in their final in-class order
    for (int r = warp; r < runs; r += 32) {
        const int base = mult * seg_start[r], cnt = run_cnt[r], off = run_off[r];
        const int label = s_label[seg_start[r]];
        for (int i = lane; i < cnt; i += 32) {
            const int o = off + t_rank[base + i];
            u_box[o] = t_box[base + i];
            u_score[o] = t_score[base + i];
            u_label[o] = label;
        }
    }
    __syncthreads();
    unsigned thr_bits = 0u;
    if (max_det > 0 && total > max_det) {
        if (threadIdx.x == 0) {
            s_prefix = 0u;
            s_need = max_det;
        }
        __syncthreads();
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            for (int b = threadIdx.x; b < 256; b += 1024) s_hist[b] = 0;
            __syncthreads();
            const unsigned prefix = s_prefix;
            const unsigned hi_mask = (pass == 0) ? 0u : (~0u << (shift + 8));
            for (int i = threadIdx.x; i < total; i += 1024) {
                const unsigned key = __float_as_uint(u_score[i]);
                if ((key & hi_mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 0xff], 1);
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                int d, remaining;
                radix_pick_digit(s_hist, s_need, threadIdx.x, &d, &remaining);
                if (threadIdx.x == 0) {
                    s_need = remaining;
                    s_prefix = prefix | ((unsigned)d << shift);
                }
            }
            __syncthreads();
        }
        thr_bits = s_prefix;
    }
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    for (int i0 = 0; i0 < total; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const bool take = i < total && __float_as_uint(u_score[i]) >= thr_bits;
        const unsigned mk = __ballot_sync(PAA_FULL, take);
        if (lane == 0) s_warp[warp] = __popc(mk);
        __syncthreads();
        int before = s_running;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (take) {
            const int row = before + __popc(mk & ((1u << lane) - 1u));
            const float4 b = u_box[i];
            out_boxes[row * 4 + 0] = b.x;
            out_boxes[row * 4 + 1] = b.y;
            out_boxes[row * 4 + 2] = b.z;
            out_boxes[row * 4 + 3] = b.w;
            out_scores[row] = u_score[i];
            out_labels[row] = (long long)u_label[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 32; ++w) t += s_warp[w];
            s_running += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_count = s_running;
}

struct VoteWorkspace {
    MlNmsWorkspace nms;
    unsigned char* alive;   // [n]
    float4* t_box;          // [2n]
    float* t_score;         // [2n]
    int* t_rank;            // [2n]
    int* run_cnt;           // [n]
    int* run_off;           // [n]
    float4* u_box;          // [2n]
    float* u_score;         // [2n]
    int* u_label;           // [2n]
    size_t total_bytes;
};

static VoteWorkspace carve_vote(void* base, int n) {
    VoteWorkspace w;
    w.nms = carve_ml_nms(base, n);
    char* p = static_cast<char*>(base);
    size_t off = w.nms.total_bytes;
    auto take = [&](size_t bytes) {
        char* q = p ? p + off : nullptr;
        off += (bytes + 255) / 256 * 256;
        return q;
    };
    const size_t nn = n > 0 ? n : 1;
    w.alive = (unsigned char*)take(nn);
    w.t_box = (float4*)take(2 * nn * 16);
    w.t_score = (float*)take(2 * nn * 4);
    w.t_rank = (int*)take(2 * nn * 4);
    w.run_cnt = (int*)take(nn * 4);
    w.run_off = (int*)take(nn * 4);
    w.u_box = (float4*)take(2 * nn * 16);
    w.u_score = (float*)take(2 * nn * 4);
    w.u_label = (int*)take(2 * nn * 4);
    w.total_bytes = off;
    return w;
}

size_t box_vote_workspace_bytes(int n) { return carve_vote(nullptr, n).total_bytes; }

int run_box_vote(const float* boxes, const float* scores, const float* labels, int n, int mode, float vote_thresh,
                 float nms_thresh, float soft_score_thresh, int max_detections, float* out_boxes, float* out_scores,
                 long long* out_labels, int32_t* out_count, void* workspace, size_t workspace_bytes,
                 cudaStream_t stream) {
    if (n < 0 || n > 65535 || mode < 0 || mode > 2) {
        set_error("paa_box_vote: n=%d outside [0, 65535] or mode=%d", n, mode);
        return PAA_ERR_UNSUPPORTED;
    }
    if (!out_count) {
        set_error("paa_box_vote: null out_count");
        return PAA_ERR_BAD_ARGUMENT;
    }
    PAA_CUDA_CHECK(cudaMemsetAsync(out_count, 0, sizeof(int32_t), stream));
    if (n == 0) return 0;
    if (!boxes || !scores || !labels || !out_boxes || !out_scores || !out_labels || !workspace) {
        set_error("paa_box_vote: null pointer");
        return PAA_ERR_BAD_ARGUMENT;
    }
    VoteWorkspace w = carve_vote(workspace, n);
    if (w.total_bytes > workspace_bytes) {
        set_error("paa_box_vote: workspace too small: need %zu bytes, got %zu", w.total_bytes, workspace_bytes);
        return PAA_ERR_WORKSPACE;
    }
    const int nbw = (n + 63) / 64;
    ml_nms_prepare_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, boxes, labels, w.nms.box, w.nms.label, w.nms.cnt);
    PAA_LAUNCH_CHECK("ml_nms_prepare_kernel");
    dim3 rgrid((n + kRankThreads - 1) / kRankThreads, 1);
    post_rank_kernel<<<rgrid, kRankThreads, 0, stream>>>(1, n, w.nms.cnt, w.nms.box, scores, w.nms.label, w.nms.s_box,
                                                         w.nms.s_score, w.nms.s_label, w.nms.s_pos, w.nms.total);
    PAA_LAUNCH_CHECK("post_rank_kernel");
    post_segments_kernel<<<1, 1024, 0, stream>>>(n, w.nms.total, w.nms.s_label, w.nms.seg_start, w.nms.n_seg);
    PAA_LAUNCH_CHECK("post_segments_kernel");
    if (mode == 0) {
        post_nms_runs_kernel<<<(n + kRunWarps - 1) / kRunWarps, kRunWarps * 32, 0, stream>>>(
            n, n, 1, nms_thresh, w.nms.seg_start, w.nms.n_seg, w.nms.s_box, w.nms.keep_sorted, w.nms.row_long, w.nms.cnt + 1);
        PAA_LAUNCH_CHECK("post_nms_runs_kernel");
        post_nms_mask_kernel<<<nms_mask_grid(n, 1), kMaskWarps * kMaskPieces * 32, 0, stream>>>(n, nbw, nms_thresh, w.nms.total, w.nms.s_box, w.nms.s_label,
                                                       w.nms.mask, w.nms.row_long, w.nms.cnt + 1);
        PAA_LAUNCH_CHECK("post_nms_mask_kernel");
        post_nms_scan_kernel<<<(n + kScanWarps - 1) / kScanWarps, kScanWarps * 32, 0, stream>>>(
            n, nbw, n, 1, w.nms.seg_start, w.nms.n_seg, w.nms.mask, w.nms.keep_sorted, w.nms.cnt + 1);
        PAA_LAUNCH_CHECK("post_nms_scan_kernel");
    }
    box_vote_kernel<<<(n + kVoteWarpsPerBlock - 1) / kVoteWarpsPerBlock, kVoteWarpsPerBlock * 32, 0, stream>>>(
        w.nms.seg_start, w.nms.n_seg, w.nms.s_box, w.nms.s_score, w.nms.keep_sorted, mode, vote_thresh,
        soft_score_thresh, w.alive, w.t_box, w.t_score, w.t_rank, w.run_cnt);
    PAA_LAUNCH_CHECK("box_vote_kernel");
    box_vote_finish_kernel<<<1, 1024, 0, stream>>>(w.nms.seg_start, w.nms.n_seg, w.nms.s_label, mode == 2 ? 2 : 1,
                                                   w.t_box, w.t_score, w.t_rank, w.run_cnt, max_detections, w.run_off,
                                                   w.u_box, w.u_score, w.u_label, out_boxes, out_scores, out_labels,
                                                   out_count);
    PAA_LAUNCH_CHECK("box_vote_finish_kernel");
    return 0;
}

}  // namespace paa
