// Host-side workspace carving shared by the entry points.  Not a public header.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace paa {

struct GtOffsets;

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Workspace of the assign+loss path.  The first `zero_bytes` are cleared by prep_step_kernel at the start
// of paa_assign (per-GT IoU maxima, image flags, the completion ticket); everything else is fully
// overwritten by the kernels before it is read.
struct LossWorkspace {
    // zeroed region
    unsigned* gtmax;        // [sumG]   bit pattern of the per-GT maximal IoU
    int* img_flags;         // [N]      bit0: some GT of the image overlaps no anchor (culling off)
    unsigned* ticket;       // [kTicketWords] completion counters; [kCtlReady] / [kCtlChunk] / [kCtlSmLive + sm]: the
                            //          hints that let bulk_focal_early_kernel work in select_gmm_kernel's shadow
    int* seg_count;         // [sumG*L] anchors matched to (GT, level)
    size_t zero_bytes;
    // plain region
    uint2* best;            // [N*A]    (IoU bits, GT index) of every anchor's best GT
    int* matched;           // [N*A]
    float* score;           // [N*A]    anchor score (combined loss) of IoU-positive anchors
    int* paa_label;         // [N*A]
    uint4* tile_gtmask;     // [N*T]    bit (g mod 128): GT g has a matched anchor in the tile
    int* part_npos;         // [sumG]
    double* part_siou;      // [sumG]
    double* local_norm;     // [2]      this rank's {num_pos, sum_iou} (before the all-reduce)
    double* block_part;     // [blocks*3] per-block partial loss sums of the final kernel
    unsigned long long* seg_pool;   // [sumG*L*kSegCap] (score bits, anchor) keys of the anchors matched to (GT, level)
    int* pos_list;          // [sumG*PAA_MAX_CANDIDATES] anchors of every GT's positive prefix (stride = levels * topk)
    GtOffsets* go;          // the step's per-image GT ranges as every kernel reads them (written by prep_step_kernel)
    int* gt_image;          // [sumG]   image of every GT
    unsigned long long* cand_sorted;   // [sumG*PAA_MAX_CANDIDATES] two-launch selection: every GT's sorted candidate keys
    int* cand_n;            // [sumG]   ... and their number
    size_t total_bytes;
};

// Capacity of one (GT, level) candidate pool.  The anchors IoU-matched to one GT on one level number a few
// hundred at most for PAA's anchor layout; a segment that overflows falls back to scanning the tiles.
constexpr int kSegCap = 1024;
// ticket[]: 0 = select_gmm_kernel's completion ticket, 1 = positive_list_kernel's; then the control words of the early
// bulk pass: "the normalisers exist", the bulk pass's chunk counter, and per SM the number of EM fits still running there
constexpr int kCtlReady = 4, kCtlChunk = 5, kCtlSmLive = 8, kCtlMaxSms = 256, kTicketWords = kCtlSmLive + kCtlMaxSms;
constexpr size_t kGtOffsetsBytes = 4 * 257 + 256;      // sizeof(GtOffsets), checked where the type is complete

inline LossWorkspace carve_loss_workspace(void* base, int N, int A, int sumG, int tiles_per_image,
                                          int loss_blocks, int L) {
    LossWorkspace w;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* q = p ? p + off : nullptr;
        off += align_up(bytes);
        return q;
    };
    w.gtmax = reinterpret_cast<unsigned*>(take(sizeof(unsigned) * (size_t)(sumG > 0 ? sumG : 1)));
    w.img_flags = reinterpret_cast<int*>(take(sizeof(int) * (size_t)N));
    w.ticket = reinterpret_cast<unsigned*>(take(sizeof(unsigned) * kTicketWords));
    w.seg_count = reinterpret_cast<int*>(take(sizeof(int) * (size_t)(sumG > 0 ? sumG : 1) * L));
    w.zero_bytes = off;
    size_t NA = (size_t)N * A;
    w.best = reinterpret_cast<uint2*>(take(sizeof(uint2) * NA));
    w.matched = reinterpret_cast<int*>(take(sizeof(int) * NA));
    w.score = reinterpret_cast<float*>(take(sizeof(float) * NA));
    w.paa_label = reinterpret_cast<int*>(take(sizeof(int) * NA));
    w.tile_gtmask = reinterpret_cast<uint4*>(take(sizeof(uint4) * (size_t)N * tiles_per_image));
    w.part_npos = reinterpret_cast<int*>(take(sizeof(int) * (size_t)(sumG > 0 ? sumG : 1)));
    w.part_siou = reinterpret_cast<double*>(take(sizeof(double) * (size_t)(sumG > 0 ? sumG : 1)));
    w.local_norm = reinterpret_cast<double*>(take(sizeof(double) * 2));
    w.block_part = reinterpret_cast<double*>(take(sizeof(double) * 3 * (size_t)loss_blocks));
    w.seg_pool = reinterpret_cast<unsigned long long*>(
        take(sizeof(unsigned long long) * (size_t)(sumG > 0 ? sumG : 1) * L * kSegCap));
    w.pos_list = reinterpret_cast<int*>(take(sizeof(int) * (size_t)(sumG > 0 ? sumG : 1) * 128));
    w.go = reinterpret_cast<GtOffsets*>(take(kGtOffsetsBytes));
    w.gt_image = reinterpret_cast<int*>(take(sizeof(int) * (size_t)(sumG > 0 ? sumG : 1)));
    w.cand_sorted = reinterpret_cast<unsigned long long*>(
        take(sizeof(unsigned long long) * (size_t)(sumG > 0 ? sumG : 1) * 128));
    w.cand_n = reinterpret_cast<int*>(take(sizeof(int) * (size_t)(sumG > 0 ? sumG : 1)));
    w.total_bytes = off;
    return w;
}

}  // namespace paa
