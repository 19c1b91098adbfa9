// Internal launcher declarations (one per kernel family).  Not a public header.
#pragma once
#include "common.cuh"
#include "layout.h"

namespace paa {

static_assert(sizeof(GtOffsets) <= kGtOffsetsBytes, "LossWorkspace::go is too small");

struct LossScalars {
    float gamma, alpha, iou_threshold, reg_loss_weight, iou_loss_weight;
    int topk, use_iou_pred, world_size;
    int flavour;       // PAA_LOSS_PAA / PAA_LOSS_ATSS / PAA_LOSS_RETINANET
    // PAA_LOSS_RETINANET: Matcher low threshold, BoxCoder weights, smooth-L1 beta, BBOX_REG_WEIGHT, images of the call
    float bg_threshold, code_w[4], beta, reg_norm_weight;
    int num_images;
    // PAA_LOSS_FCOS: per-level stride and centre-sampling radius in pixels (0 = off), IoU loss type, target scaling
    float fcos_stride[PAA_MAX_LEVELS], fcos_radius[PAA_MAX_LEVELS];
    int fcos_iou_type, fcos_norm;
    int atss_type;     // PAA_ATSS_POSITIVE_*
    int seg_cap;       // usable entries of a (GT, level) candidate pool, <= kSegCap (PAA_SEG_CAP shrinks it for tests)
    // Grids and the workspace are sized by these, never by the step's actual GT counts, which the kernels read from
    // device memory (LossWorkspace::go): a captured CUDA graph stays valid for any batch within the capacities.
    int gt_capacity;   // upper bound on the GTs of the call (== their number unless PaaLossArgs::gt_offsets_dev is set)
    int gt_parts;      // GT-list parts of the coarse IoU tiles (from the per-image capacity)
};

// Peer-memory exchange of the loss normalisers (include/paa_b200.h, PaaLossArgs::peer_norm).
// Buffer layout (doubles): slot (parity p, rank r) at (p * kPeerMaxRanks + r) * 4 = {num_pos, sum_iou, epoch, -};
// the owner's call counter at kPeerEpochOffset.
constexpr int kPeerMaxRanks = PAA_MAX_PEERS;
constexpr int kPeerEpochOffset = 2 * kPeerMaxRanks * 4;
struct PeerExchange {
    double* buf[PAA_MAX_PEERS];
    int rank, world;       // world == 0: disabled
    int* status;           // host-mapped [4] {timed out, rank waited for, epoch, -} or null (PaaLossArgs::peer_status)
    unsigned long long timeout_ns;   // 0: wait for the peers without a deadline, like the all-reduce it replaces
};

struct LossDebug {
    int* matched_idx;
    int* iou_labels;
    float* combined_loss;
    int* cand_idx;
    int* cand_cnt;
    int* num_pos;
    double* gmm;
    int* paa_labels;
};

// assign.cu
int first_heavy_level(const Geometry& geo);
int gt_parts_for(int max_gt_per_image);
// first launch of every assign call: clears the workspace's zeroed prefix (and, with `clear_heavy_best`, the coarse
// levels' best-GT keys that the GT-list parts merge by atomicMax) and leaves the per-image GT ranges in ws.go --
// from `dev_offsets` when the caller keeps them on the device, else from the host values in `host_go`
int launch_prep_step(const Geometry& geo, const GtOffsets& host_go, const int* dev_offsets, const LossScalars& sc,
                     const LossWorkspace& ws, void* zero_base, bool clear_heavy_best, bool clear_all_best,
                     cudaStream_t stream);
int launch_iou_match(const Geometry& geo, const float* gt_boxes, const LossScalars& sc, const LossWorkspace& ws,
                     cudaStream_t stream);
int launch_match_score(const Geometry& geo, const float* gt_boxes,
                       const int64_t* gt_labels, const LossScalars& sc, const LossWorkspace& ws,
                       const float* teacher_score, const LossDebug& dbg, cudaStream_t stream);
int launch_select_gmm(const Geometry& geo, const float* gt_boxes,
                      const int64_t* gt_labels, const LossScalars& sc, const LossWorkspace& ws,
                      const float* score_src, double* normalisers, const PeerExchange& px, const LossDebug& dbg,
                      cudaStream_t stream);
// true when launch_select_gmm takes the two-launch form for this call (many GTs: the fits run as a launch of their own)
bool select_gmm_two_launch(const LossScalars& sc);
// atss.cu
int launch_atss_assign(const Geometry& geo, const float* gt_boxes,
                       const int64_t* gt_labels, const LossScalars& sc, const LossWorkspace& ws, double* normalisers,
                       const PeerExchange& px, const LossDebug& dbg, cudaStream_t stream);
int launch_norm_wait(const PeerExchange& px, double* normalisers, cudaStream_t stream);
// folds per-tile {count, sum} partials into local_norm / normalisers and publishes them to the peers
int launch_fold_norm(const double* tile_part, int tiles, double* local_norm, double* normalisers,
                     const PeerExchange& px, cudaStream_t stream);
// fcos.cu
// `ssc`: ATSS POSITIVE_TYPE 'SSC' -- the same rule on anchor centres with a 0.01 margin, ATSS centerness sums
int launch_fcos_assign(const Geometry& geo, const float* gt_boxes, const int64_t* gt_labels,
                       const LossScalars& sc, const LossWorkspace& ws, double* normalisers, const PeerExchange& px,
                       const LossDebug& dbg, cudaStream_t stream, bool ssc = false);
// retina.cu
// `atss_iou`: ATSS POSITIVE_TYPE 'IoU' -- additionally ignores positives whose centre is outside their GT, sums
// the centerness targets and publishes both normalisers through `px`
int launch_retinanet_assign(const Geometry& geo, const float* gt_boxes, const int64_t* gt_labels,
                            const LossScalars& sc, const LossWorkspace& ws, double* normalisers, const LossDebug& dbg,
                            cudaStream_t stream, bool atss_iou = false, const PeerExchange* px = nullptr);

// rpn.cu
int launch_rpn_loss(const Geometry& geo, const GtOffsets& go, const float* gt_boxes, const int* matched,
                    const long long* sampled, int n_pos, int n_neg, const float* weights, float beta,
                    const float* grad_losses, float* losses, cudaStream_t stream);
// aux.cu
int launch_selftest_roots(const float* x, int n, float* out, cudaStream_t stream);
// loss.cu
int loss_grid_blocks(int num_images, int tiles_per_image);
int launch_final_loss(const Geometry& geo, const float* gt_boxes, const int64_t* gt_labels,
                      const LossScalars& sc, const LossWorkspace& ws, const double* normalisers,
                      const float* grad_losses, float* losses, bool write_grads, cudaStream_t stream);
int launch_rescale_grads(const Geometry& geo, const float* old_g, const float* new_g, cudaStream_t stream);
int launch_focal_forward(const float* logits, const int* targets, int n, int C, float gamma, float alpha,
                         float* losses, cudaStream_t stream);
int launch_focal_backward(const float* logits, const int* targets, const float* d_losses, int n, int C,
                          float gamma, float alpha, float* d_logits, cudaStream_t stream);

}  // namespace paa
