/*
 * paa_b200.h -- C ABI of libpaa_b200.so: the PAA assign + loss path and the PAA post-processing
 * path as hand-written sm_100a CUDA kernels.
 *
 * Drop-in boundary.  In the reference (JunhoPark0314/PAA) this path is Python that sits on the
 * pybind11 module `paa_core._C` (paa_core/csrc/vision.cpp:10-27).  This library replaces, for the
 * path only:
 *
 *   paa_assign / paa_loss / paa_assign_loss
 *       everything PAALossComputation.__call__ does between its arguments and its three returned
 *       losses (paa_core/modeling/rpn/paa/loss.py:267-359), i.e. boxlist_iou
 *       (structures/boxlist_ops.py:81-116), Matcher (modeling/matcher.py:42-113), BoxCoder
 *       encode/decode (modeling/rpn/atss/atss.py:33-50,68-96), concat_box_prediction_layers
 *       (modeling/rpn/utils.py:17-45; never materialised here), `_C.sigmoid_focalloss_forward`
 *       / `_backward` (csrc/SigmoidFocalLoss.h:10-41, csrc/cuda/SigmoidFocalLoss_cuda.cu:20-101),
 *       GIoULoss (loss.py:46-87), compute_paa (loss.py:128-236) including the
 *       sklearn.mixture.GaussianMixture fit (loss.py:197-203), compute_ious (loss.py:258-265), the
 *       two normalisers (loss.py:321-322,338) and the autograd backward of the three losses.
 *   paa_postprocess
 *       PAAPostProcessor.forward (paa_core/modeling/rpn/paa/inference.py:84-159) including
 *       `_C.ml_nms` (csrc/ml_nms.h:10-27, csrc/cuda/ml_nms.cu:26-136), the kthvalue cut
 *       (inference.py:114-122) and score voting (inference.py:123-157).
 *
 * Conventions: plain C, device pointers and sizes only (no torch / pybind types); every call
 * enqueues work on the caller's stream and returns without synchronising the host; the library
 * never allocates device memory -- the caller passes a workspace of *_workspace_bytes(); return
 * value 0 = OK, >0 = cudaError_t, <0 = PAA_ERR_*; paa_last_error() gives a thread-local message.
 * One caller thread per (device, stream); calls on different streams must use different
 * workspaces.  All floating tensors are float32.  Head tensors (and their gradients) are dense in one of two
 * memory layouts, the same for every tensor of a call (`head_layout`): NCHW-contiguous as PAAHead.forward produces
 * them (paa.py:90-108), or channels-last (NHWC: what a `torch.channels_last` / AMP pipeline hands over) -- in memory
 * the [N, H*W*a, ch] tensor that permute_and_flatten (rpn/utils.py:10-14) would build, consumed in place.
 */
#ifndef PAA_B200_H_
#define PAA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAA_ABI_VERSION 8
#define PAA_MAX_LEVELS 8
#define PAA_MAX_IMAGES 256      /* images per call (per rank) */
#define PAA_MAX_CANDIDATES 128  /* num_levels * topk must not exceed this */
#define PAA_MAX_PEERS 16        /* ranks in a peer-memory normaliser exchange */
#define PAA_PEER_BUFFER_DOUBLES 256  /* size of every rank's exchange buffer */

#define PAA_ERR_BAD_ARGUMENT (-1)
#define PAA_ERR_WORKSPACE    (-2)
#define PAA_ERR_EMPTY_TARGET (-3)   /* an image without GT: matcher.py:53-58 raises ValueError */
#define PAA_ERR_UNSUPPORTED  (-4)

#define PAA_LAYOUT_NCHW 0       /* head tensors [N, a*ch, H, W], W fastest */
#define PAA_LAYOUT_NHWC 1       /* same logical shape, channels fastest (torch.channels_last) */

/* One FPN level of the head outputs (paa.py:90-108) and its anchors
 * (anchor_generator.py:112-125).  hw = H*W; with anchors_per_loc = a the level holds hw*a anchors,
 * anchor i sits at location i / a with channel block i % a (rpn/utils.py:10-14). */
typedef struct PaaLevel {
    const float* box_cls;         /* [N, a*C, H, W] */
    const float* box_regression;  /* [N, a*4, H, W] */
    const float* iou_pred;        /* [N, a*1, H, W] or NULL (USE_IOU_PRED False) */
    const float* anchors;         /* [hw*a, 4] xyxy; image i reads anchors + i*anchor_image_stride */
    float* grad_box_cls;          /* same shapes as the inputs; all three NULL = no gradients */
    float* grad_box_regression;
    float* grad_iou_pred;
    int32_t hw;
    int32_t grid_w;               /* W of the level's H x W map (hw = H*W), or 0 if unknown: lets the IoU matching
                                     take 2-D patches of anchors per warp instead of runs of a row */
} PaaLevel;

typedef struct PaaLossArgs {
    int32_t num_images;           /* N on this rank */
    int32_t num_levels;           /* <= PAA_MAX_LEVELS */
    int32_t num_classes;          /* C = NUM_CLASSES - 1 (80) */
    int32_t anchors_per_loc;      /* 1 for every PAA config */
    int32_t topk;                 /* cfg.MODEL.PAA.TOPK (9) */
    int32_t use_iou_pred;         /* cfg.MODEL.PAA.USE_IOU_PRED */
    int32_t world_size;           /* WORLD_SIZE as read by loss.py:18-19 */
    int32_t loss_flavour;         /* PAA_LOSS_PAA: regression weight / BCE target = IoU(pred, target) (paa/loss.py:331-349);
                                     PAA_LOSS_ATSS: = centerness target (atss/loss.py:233-245,262-272), assignment by
                                     paa_atss_assign;
                                     PAA_LOSS_RETINANET: labels straight from the Matcher (-1 = ignored), smooth-L1 on
                                     the regression deltas (retinanet/loss.py:45-81), assignment by
                                     paa_retinanet_assign; use_iou_pred must be 0;
                                     PAA_LOSS_FCOS: anchor-free (fcos/loss.py): `anchors` are the locations as points
                                     (x, y, x, y), the regression maps are distances (l, t, r, b), IOULoss
                                     (layers/iou_loss.py) weighted by the centerness targets, BCE on the centerness
                                     map (passed as iou_pred), assignment by paa_fcos_assign */
    float gamma, alpha;           /* focal loss, cfg.MODEL.PAA.LOSS_GAMMA / LOSS_ALPHA */
    float iou_threshold;          /* Matcher high threshold; PAA: high == low (loss.py:38-40) */
    float reg_loss_weight;        /* cfg.MODEL.PAA.REG_LOSS_WEIGHT */
    float iou_loss_weight;        /* cfg.MODEL.PAA.IOU_LOSS_WEIGHT */
    float bg_iou_threshold;       /* PAA_LOSS_RETINANET: Matcher low threshold (BG_IOU_THRESHOLD); else unused */
    int64_t anchor_image_stride;  /* in floats; 0 when all images share the anchor tensors */
    PaaLevel levels[PAA_MAX_LEVELS];
    const float* gt_boxes;        /* device [sum G, 4] xyxy */
    const int64_t* gt_labels;     /* device [sum G], 1..C */
    int32_t gt_offsets[PAA_MAX_IMAGES + 1];  /* HOST values: image i owns GT [off[i], off[i+1]); see gt_offsets_dev */
    void* workspace;              /* device, >= paa_loss_workspace_bytes(), 256-byte aligned */
    size_t workspace_bytes;
    /* outputs (device) */
    double* normalisers;          /* [2] = {num_pos, sum of positives' IoU} of THIS rank after
                                     paa_assign; the caller all-reduces (SUM) it in place across
                                     ranks before paa_loss (loss.py:321,338) */
    float* losses;                /* [3] = loss_cls, loss_reg, loss_iou_pred */
    const float* grad_losses;     /* [3] upstream d/d(loss) or NULL for ones */
    /* optional debug / parity outputs (device, nullable) */
    int32_t* dbg_matched_idx;     /* [N, A]  Matcher result (-1 or GT index within the image) */
    int32_t* dbg_iou_labels;      /* [N, A]  IoU-based class label (0 = background) */
    float* dbg_combined_loss;     /* [N, A]  anchor score; only IoU-positive entries are defined */
    int32_t* dbg_cand_idx;        /* [sum G, num_levels*topk] sorted candidates (anchor index in image) */
    int32_t* dbg_cand_cnt;        /* [sum G] */
    int32_t* dbg_num_pos;         /* [sum G] length of the positive prefix */
    double* dbg_gmm;              /* [sum G, 8] = w0,w1,mu0,mu1,var0,var1,n_iter,converged */
    int32_t* dbg_paa_labels;      /* [N, A] final class label (0 = negative) */
    const float* teacher_combined_loss;  /* [N, A] if non-NULL, candidate selection and the GMM consume
                                            this instead of the kernel's own anchor scores
                                            (stage-wise parity protocol) */
    /* Optional peer-memory exchange of the two normalisers (replaces the caller's all-reduce between
     * paa_assign and paa_loss; loss.py:22-28,321,338).  peer_norm[r] is rank r's exchange buffer
     * (PAA_PEER_BUFFER_DOUBLES doubles, zeroed once, persistent across calls) as mapped into THIS process
     * (NVLink peer mapping / symmetric memory); peer_norm[rank] is this rank's own buffer.  paa_assign's last
     * kernel stores this rank's {num_pos, sum_iou} into every rank's buffer, paa_loss first waits (on its own
     * buffer only) until all world_size contributions of the step have arrived and leaves their sum, added
     * in rank order, in `normalisers`.  All ranks must make the same sequence of calls.  peer_norm[0] NULL
     * = disabled.  See peer_timeout_s / peer_status below for what happens when a rank never arrives. */
    int32_t rank;
    int32_t reserved2;
    double* peer_norm[PAA_MAX_PEERS];
    /* PAA_LOSS_RETINANET only */
    float box_code_weights[4];    /* BoxCoder weights wx, wy, ww, wh (box_coder.py:22-50; 10, 10, 5, 5) */
    float smooth_l1_beta;         /* cfg.MODEL.RETINANET.BBOX_REG_BETA */
    float reg_norm_weight;        /* cfg.MODEL.RETINANET.BBOX_REG_WEIGHT: loss_reg / max(1, num_pos * this) */
    /* PAA_LOSS_FCOS only */
    float fcos_strides[PAA_MAX_LEVELS];   /* cfg.MODEL.FCOS.FPN_STRIDES */
    float fcos_center_radius;     /* cfg.MODEL.FCOS.CENTER_SAMPLING_RADIUS (0 = every location inside the GT) */
    int32_t fcos_iou_loss_type;   /* PAA_IOU_LOSS_IOU / _LINEAR / _GIOU (cfg.MODEL.FCOS.IOU_LOSS_TYPE) */
    int32_t fcos_norm_reg_targets;/* cfg.MODEL.FCOS.NORM_REG_TARGETS: targets divided by the level's stride */
    /* PAA_LOSS_ATSS only: cfg.MODEL.ATSS.POSITIVE_TYPE (atss/loss.py:88-229) */
    int32_t atss_positive_type;   /* PAA_ATSS_POSITIVE_ATSS / _SSC / _IOU */
    /* Peer exchange only.  The wait for the other ranks' normalisers has no deadline when peer_timeout_s <= 0 --
     * the semantics of the all-reduce it replaces (loss.py:22-28): a rank that is late (checkpoint, data loader)
     * is waited for.  With a deadline, a contribution that does not arrive in time is an error, never data:
     * the waiting kernel writes {1, rank it waited for, step counter, 0} to peer_status -- four int32 in
     * host-mapped (pinned) memory, nullable -- and traps, so every later CUDA call of the process fails. */
    float peer_timeout_s;
    int32_t reserved3;
    int32_t* peer_status;
    /* Optional: the per-image GT ranges kept on the DEVICE ([num_images + 1] int32, same meaning as gt_offsets).
     * When set, every kernel of the step takes the ranges from here (the host copy in gt_offsets is only used to
     * check the arguments), the per-GT grids and the workspace are sized by gt_capacity (an upper bound on the
     * GTs of a call; use the same value for *_workspace_bytes) and the GT-list split of crowded images by
     * gt_per_image_capacity.  No kernel parameter or grid size then depends on the batch's GT counts: a CUDA graph
     * captured around paa_assign + paa_loss can be replayed after the caller has rewritten gt_offsets_dev, gt_boxes
     * and gt_labels in place with another batch of num_images images within the capacities (training feeds new
     * targets every step, paa.py:137-148).  The caller checks what the host path checks (an image without GT). */
    const int32_t* gt_offsets_dev;
    int32_t gt_capacity;
    int32_t gt_per_image_capacity;
    int32_t head_layout;          /* PAA_LAYOUT_NCHW (0) or PAA_LAYOUT_NHWC: layout of every box_cls / box_regression /
                                     iou_pred / grad_* tensor of the call */
    int32_t reserved4;
} PaaLossArgs;

typedef struct PaaPostArgs {
    int32_t num_images;
    int32_t num_levels;
    int32_t num_classes;          /* C = 80 foreground classes; labels are 1..C */
    int32_t anchors_per_loc;
    int32_t pre_nms_top_n;        /* cfg.MODEL.PAA.PRE_NMS_TOP_N (1000) per level */
    int32_t detections_per_img;   /* cfg.TEST.DETECTIONS_PER_IMG (100); <=0 disables the cut */
    int32_t score_voting;         /* cfg.MODEL.PAA.INFERENCE_SCORE_VOTING */
    int32_t skip_nms;             /* bbox_aug_enabled && !bbox_aug_vote (inference.py:96-97) */
    float pre_nms_thresh;         /* cfg.MODEL.PAA.INFERENCE_TH (0.05) */
    float nms_thresh;             /* cfg.MODEL.PAA.NMS_TH (0.6) */
    int64_t anchor_image_stride;
    PaaLevel levels[PAA_MAX_LEVELS];  /* grad_* unused */
    float image_wh[PAA_MAX_IMAGES][2];  /* HOST values: BoxList.size = (width, height) per image */
    void* workspace;
    size_t workspace_bytes;
    /* outputs (device).  cap = num_levels * pre_nms_top_n rows per image */
    float* out_boxes;             /* [N, cap, 4] */
    float* out_scores;            /* [N, cap] */
    int64_t* out_labels;          /* [N, cap] 1-based */
    int32_t* out_count;           /* [N] valid rows per image */
    /* optional debug outputs */
    float* dbg_pre_boxes;         /* [N, cap, 4] pre-NMS candidates, level-major */
    float* dbg_pre_scores;        /* [N, cap] */
    int32_t* dbg_pre_labels;      /* [N, cap] */
    int32_t* dbg_pre_count;       /* [N, num_levels] candidates kept per level */
    uint8_t* dbg_nms_keep;        /* [N, cap] 1 = survived NMS (before the detections_per_img cut) */
    /* How a selected (anchor, regression) pair becomes a box (0 keeps PAA / ATSS behaviour):
     *   PAA_DECODE_ATSS_BOX  rpn/atss/atss.py:68-96 ('BOX' BoxCoder: /10, /5, x2 = cx + (w-1)/2)
     *   PAA_DECODE_LEGACY    modeling/box_coder.py:51-95 (RetinaNet: weights, x2 = cx + w/2 - 1)
     *   PAA_DECODE_LTRB      rpn/fcos/inference.py:93-98 (FCOS: "anchors" are points (x, y, x, y), the four
     *                        regression channels are distances to the left / top / right / bottom edge) */
    int32_t box_decode;
    float decode_weights[4];      /* PAA_DECODE_LEGACY: wx, wy, ww, wh */
    float decode_clip;            /* PAA_DECODE_LEGACY: bbox_xform_clip (log(1000/16)) */
    int32_t head_layout;          /* PAA_LAYOUT_NCHW (0) or PAA_LAYOUT_NHWC */
} PaaPostArgs;
/* The plain RPN loss (paa_core/modeling/rpn/loss.py:98-137) over a GIVEN sample of anchors.  The Matcher result comes
 * from paa_retinanet_assign (unit GT labels; dbg_matched_idx), the sample from the caller's balanced sampler
 * (balanced_positive_negative_sampler.py: torch.randperm).  Enqueues memsets of the gradient tensors and one kernel. */
typedef struct PaaRpnArgs {
    int32_t num_images;
    int32_t num_levels;
    int32_t anchors_per_loc;
    int32_t head_layout;          /* PAA_LAYOUT_NCHW / PAA_LAYOUT_NHWC */
    int64_t anchor_image_stride;
    PaaLevel levels[PAA_MAX_LEVELS];  /* box_cls = objectness [N, a, H, W], box_regression [N, a*4, H, W], iou_pred NULL;
                                         grad_box_cls / grad_box_regression nullable (no gradients) */
    const float* gt_boxes;        /* device [sum G, 4] xyxy */
    int32_t gt_offsets[PAA_MAX_IMAGES + 1];   /* HOST values */
    const int32_t* matched_idx;   /* device [N, A]: GT index within the image of every anchor (>= 0 for the positives) */
    const int64_t* sampled;       /* device [n_pos + n_neg]: image * A + anchor of the sampled anchors, positives first */
    int32_t n_pos, n_neg;
    float box_code_weights[4];    /* BoxCoder weights (1, 1, 1, 1 for the RPN) */
    float smooth_l1_beta;         /* 1 / 9 */
    int32_t reserved;
    float* losses;                /* device [2] = objectness_loss, box_loss */
    const float* grad_losses;     /* device [2] upstream gradients or NULL for ones */
} PaaRpnArgs;
#define PAA_LOSS_PAA  0
#define PAA_LOSS_ATSS 1
#define PAA_LOSS_RETINANET 2
#define PAA_LOSS_FCOS 3
#define PAA_ATSS_POSITIVE_ATSS 0   /* topk nearest anchors per level, IoU threshold mean + std (the default) */
#define PAA_ATSS_POSITIVE_SSC  1   /* FCOS's rule on the anchor centres: inside by > 0.01, size range, min area */
#define PAA_ATSS_POSITIVE_IOU  2   /* Matcher(iou_threshold, bg_iou_threshold) labels; positives whose centre is not
                                      inside their GT, and anchors between the thresholds, are ignored (-1) */
#define PAA_IOU_LOSS_IOU    0
#define PAA_IOU_LOSS_LINEAR 1
#define PAA_IOU_LOSS_GIOU   2
#define PAA_DECODE_ATSS_BOX 0
#define PAA_DECODE_LEGACY   1
#define PAA_DECODE_LTRB     2

int paa_abi_version(void);
const char* paa_last_error(void);

/* Workspace sizes (bytes) for the given problem; pure host functions. */
size_t paa_loss_workspace_bytes(int num_images, int anchors_per_image, int num_gt_total,
                                int num_levels, int topk);
size_t paa_postprocess_workspace_bytes(int num_images, int anchors_per_image, int num_classes,
                                       int num_levels, int pre_nms_top_n);

/* Stages 1-4: IoU matching, anchor scores, per-GT top-k + GMM, PAA labels, this rank's
 * normalisers.  `stream` is a cudaStream_t. */
int paa_assign(const PaaLossArgs* args, void* stream);
/* Stage 5: the three losses and (if grad pointers are set) their gradients, using
 * args->normalisers as already reduced over ranks. */
int paa_loss(const PaaLossArgs* args, void* stream);
/* Both, for world_size == 1. */
int paa_assign_loss(const PaaLossArgs* args, void* stream);
/* Rescales gradients written by paa_loss by upstream grads that differ from the ones used then:
 * grad_x *= new_grad_losses[j] / old_grad_losses[j]; both are device [3]. */
int paa_rescale_grads(const PaaLossArgs* args, const float* old_grad_losses,
                      const float* new_grad_losses, void* stream);
int paa_rpn_loss(const PaaRpnArgs* args, void* stream);

int paa_postprocess(const PaaPostArgs* args, void* stream);

/* Stand-alone kernels behind the reference's `_C` entry points, for callers that use them
 * directly (boxlist_ml_nms, SigmoidFocalLoss). */
/* csrc/ml_nms.h:10-27: keep[n] (1 = kept) in input order; workspace >= paa_ml_nms_workspace_bytes(n). */
size_t paa_ml_nms_workspace_bytes(int n);
int paa_ml_nms(const float* boxes, const float* scores, const float* labels, int n, float thresh,
               uint8_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes, void* stream);
/* csrc/SigmoidFocalLoss.h:10-41 on [n, C] row-major logits with int32 targets. */
int paa_sigmoid_focal_loss_forward(const float* logits, const int32_t* targets, int n, int num_classes,
                                   float gamma, float alpha, float* losses, void* stream);
int paa_sigmoid_focal_loss_backward(const float* logits, const int32_t* targets, const float* d_losses,
                                    int n, int num_classes, float gamma, float alpha, float* d_logits,
                                    void* stream);

/* ATSS anchor assignment (rpn/atss/loss.py:139-197, POSITIVE_TYPE 'ATSS') in place of paa_assign: per GT and
 * level the `topk` anchors nearest to the GT centre, IoU threshold = mean + std of their IoUs, centre inside
 * the GT, conflicts to the larger IoU.  Same arguments / workspace / normalisers protocol as paa_assign
 * (normalisers = {num_pos, sum of centerness targets}); follow with paa_loss and loss_flavour = PAA_LOSS_ATSS.
 * dbg_cand_idx receives the candidates (level-major, nearest first), dbg_gmm[g*8] the GT's IoU threshold.
 * atss_positive_type selects the reference's other two rules instead (no candidates then): _SSC needs at most five
 * levels, _IOU reads iou_threshold / bg_iou_threshold (cfg.MODEL.ATSS.FG_IOU_THRESHOLD / BG_IOU_THRESHOLD). */
int paa_atss_assign(const PaaLossArgs* args, void* stream);

/* RetinaNet anchor labelling (rpn/loss.py:41-88 as used by rpn/retinanet/loss.py:45-56) in place of paa_assign:
 * boxlist_iou + Matcher(iou_threshold, bg_iou_threshold, allow_low_quality_matches) per image; label = class of
 * the matched GT, 0 below the low threshold, -1 (ignored by the focal loss) between the thresholds.  Works for any
 * anchors_per_loc (9 in RetinaNet).  normalisers[0] = num_pos of this rank (the reference does not reduce it over
 * ranks); follow with paa_loss and loss_flavour = PAA_LOSS_RETINANET, which returns losses = {cls, reg, 0}.
 * dbg_matched_idx receives the Matcher result (-2 = between thresholds), dbg_iou_labels the labels. */
int paa_retinanet_assign(const PaaLossArgs* args, void* stream);

/* FCOS target assignment (rpn/fcos/loss.py:105-201) in place of paa_assign: every location takes, among the GTs
 * that contain it (or whose centre region of fcos_center_radius * stride contains it) and whose largest distance
 * max(l, t, r, b) lies in the level's size range ((-1, 64), (64, 128), (128, 256), (256, 512), (512, inf): at most
 * five levels), the one with the smallest area.  anchors_per_loc must be 1 and levels[l].anchors hold the points
 * (x, y, x, y).  normalisers = {num_pos, sum of centerness targets}, exchanged over ranks like PAA's; follow with
 * paa_loss and loss_flavour = PAA_LOSS_FCOS.  dbg_matched_idx receives the chosen GT (-1 = background). */
int paa_fcos_assign(const PaaLossArgs* args, void* stream);

/* ---- operators on either side of the path ----------------------------------------- */
/* AnchorGenerator.grid_anchors (anchor_generator.py:73-95) for one level: out[(y*W + x)*a + k] =
 * cell_anchors[k] + (x*stride, y*stride, x*stride, y*stride), float32 like torch.arange(0, W*stride, stride).
 * cell_anchors [a, 4] and out [H*W*a, 4] are device pointers (out 16-byte aligned). */
int paa_grid_anchors(const float* cell_anchors, int anchors_per_loc, int grid_h, int grid_w, float stride,
                     float* out_anchors, void* stream);
/* AnchorGenerator.add_visibility_to (anchor_generator.py:97-110): out[i] = anchor i lies inside the image
 * grown by straddle_thresh (all ones when straddle_thresh < 0). */
int paa_anchor_visibility(const float* anchors, int64_t n, float image_w, float image_h, float straddle_thresh,
                          uint8_t* out, void* stream);
/* boxlist_iou (structures/boxlist_ops.py:81-116): out [n1, n2] row-major, "+1" convention, the reference's
 * float32 operation order (bit-exact).  Callers check the BoxList sizes (boxlist_ops.py:95-97). */
int paa_boxlist_iou(const float* boxes1, int n1, const float* boxes2, int n2, float* out, void* stream);

/* merge_result_from_multi_scales for ONE image (engine/bbox_aug_vote.py:140-310): the detections pooled from
 * all test-time augmentations (boxes [n,4], scores [n], labels [n] as float, n <= 65535) are merged class by
 * class -- mode 0 'nms' (per-class NMS at nms_thresh), 1 'vote' (bbox_vote), 2 'soft-vote' (soft_bbox_vote,
 * decayed members kept while their score stays >= soft_score_thresh) -- concatenated in class order, and cut to
 * the rows scoring >= the max_detections-th best when there are more (max_detections <= 0: no cut).  Outputs have
 * room for n rows (2n for soft-vote); out_count receives the number of rows. */
size_t paa_box_vote_workspace_bytes(int n);
int paa_box_vote(const float* boxes, const float* scores, const float* labels, int n, int mode, float vote_thresh,
                 float nms_thresh, float soft_score_thresh, int max_detections, float* out_boxes, float* out_scores,
                 int64_t* out_labels, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream);

/* Measurement aid (not on the reference's interface): while enabled, every launch of the chosen kernel
 * is bracketed by CUDA events on its own stream; paa_kernel_timing_end waits for them and returns the
 * summed device time and the number of launches.  Do not enable during CUDA-graph capture. */
#define PAA_KERNEL_PASS1        1   /* iou_match_kernel: every anchor's best GT, every GT's maximal IoU */
#define PAA_KERNEL_MATCH_SCORE  2
#define PAA_KERNEL_SELECT_GMM   3
#define PAA_KERNEL_FINAL_LOSS   4   /* bulk_focal_kernel: every logit read once, its gradient written once */
#define PAA_KERNEL_POSITIVE_TERMS 5 /* positive_terms_kernel */
#define PAA_KERNEL_FINISH_LOSS  6   /* finish_loss_kernel */
#define PAA_KERNEL_NORM_WAIT    7   /* norm_wait_kernel (peer exchange only) */
#define PAA_KERNEL_PREP_STEP    8   /* prep_step_kernel: workspace prefix cleared, GT ranges to device memory */
#define PAA_KERNEL_POST_CANDIDATES 10
#define PAA_KERNEL_POST_FILTER  11
#define PAA_KERNEL_POST_SELECT  12
#define PAA_KERNEL_POST_RANK    13
#define PAA_KERNEL_POST_NMS_MASK 14
#define PAA_KERNEL_POST_NMS_SCAN 15
#define PAA_KERNEL_POST_FINISH  16
#define PAA_KERNEL_POST_VOTE    17
#define PAA_KERNEL_POST_THRESHOLD 18
#define PAA_KERNEL_POST_SEGMENTS 19
#define PAA_KERNEL_POST_NMS_RUNS 20  /* post_nms_runs_kernel: greedy NMS of every label run by one warp, no mask in memory */
/* Self-test of the branch-free float32 square root / reciprocal the mixture fit uses in place of the IEEE library
 * routines: out [4, n] = {fast sqrt(x), __fsqrt_rn(x), fast 1/x, __fdiv_rn(1, x)} for device arrays x [n]. */
int paa_selftest_roots(const float* x, int n, float* out, void* stream);
int paa_kernel_timing_begin(int kernel_id);
int paa_kernel_timing_end(float* total_ms, int32_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* PAA_B200_H_ */
