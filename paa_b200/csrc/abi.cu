// extern "C" entry points of libpaa_b200.so (see include/paa_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "post.h"

namespace paa {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// ---- per-kernel event timing (measurement aid) ------------------------------------------------
static const int kMaxTimed = 4096;
static int g_timed_kernel = 0;
static int g_timed_count = 0;
static cudaEvent_t g_ev0[kMaxTimed], g_ev1[kMaxTimed];
static int g_ev_created = 0;

KernelTimer::KernelTimer(int kernel_id, cudaStream_t stream) : stream_(stream), slot_(-1) {
    if (kernel_id != g_timed_kernel || g_timed_count >= kMaxTimed) return;
    slot_ = g_timed_count++;
    while (g_ev_created <= slot_) {
        cudaEventCreate(&g_ev0[g_ev_created]);
        cudaEventCreate(&g_ev1[g_ev_created]);
        ++g_ev_created;
    }
    cudaEventRecord(g_ev0[slot_], stream_);
}

KernelTimer::~KernelTimer() {
    if (slot_ >= 0) cudaEventRecord(g_ev1[slot_], stream_);
}

static int tiles_of(int n_anchor) { return (n_anchor + PAA_TILE - 1) / PAA_TILE; }

// Validates the caller's description of the head tensors and builds the kernel-side view.
static int build_geometry(int num_images, int num_levels, int num_classes, int anchors_per_loc,
                          long long anchor_image_stride, const PaaLevel* levels, bool need_iou,
                          int head_layout, Geometry* geo) {
    if (head_layout != PAA_LAYOUT_NCHW && head_layout != PAA_LAYOUT_NHWC) {
        set_error("head_layout=%d (PAA_LAYOUT_NCHW or PAA_LAYOUT_NHWC)", head_layout);
        return PAA_ERR_BAD_ARGUMENT;
    }
    if (num_images < 1 || num_images > PAA_MAX_IMAGES) {
        set_error("num_images=%d outside [1, %d]", num_images, PAA_MAX_IMAGES);
        return PAA_ERR_BAD_ARGUMENT;
    }
    if (num_levels < 1 || num_levels > PAA_MAX_LEVELS) {
        set_error("num_levels=%d outside [1, %d]", num_levels, PAA_MAX_LEVELS);
        return PAA_ERR_BAD_ARGUMENT;
    }
    if (num_classes < 1 || anchors_per_loc < 1) {
        set_error("num_classes=%d / anchors_per_loc=%d must be positive", num_classes, anchors_per_loc);
        return PAA_ERR_BAD_ARGUMENT;
    }
    memset(geo, 0, sizeof(*geo));
    geo->num_levels = num_levels;
    geo->num_images = num_images;
    geo->C = num_classes;
    geo->apl = anchors_per_loc;
    geo->anchor_image_stride = anchor_image_stride;
    geo->nhwc = head_layout == PAA_LAYOUT_NHWC ? 1 : 0;
    int a_off = 0, t_off = 0;
    for (int l = 0; l < num_levels; ++l) {
        const PaaLevel& s = levels[l];
        if (s.hw < 1 || !s.box_cls || !s.box_regression || !s.anchors || (need_iou && !s.iou_pred)) {
            set_error("level %d: null tensor or hw=%d", l, s.hw);
            return PAA_ERR_BAD_ARGUMENT;
        }
        LevelView& v = geo->lv[l];
        v.cls = s.box_cls;
        v.reg = s.box_regression;
        v.iou = s.iou_pred;
        v.anchors = s.anchors;
        v.g_cls = s.grad_box_cls;
        v.g_reg = s.grad_box_regression;
        v.g_iou = s.grad_iou_pred;
        v.hw = s.hw;
        v.grid_w = (s.grid_w > 0 && s.hw % s.grid_w == 0) ? s.grid_w : 0;
        v.n_anchor = s.hw * anchors_per_loc;
        v.a_off = a_off;
        v.tile_off = t_off;
        a_off += v.n_anchor;
        t_off += tiles_of(v.n_anchor);
    }
    geo->A = a_off;
    geo->tiles_per_image = t_off;
    return 0;
}

static int tiles_upper_bound(int anchors_per_image, int num_levels) {
    return (anchors_per_image + PAA_TILE - 1) / PAA_TILE + num_levels;
}

struct LossPlan {
    PeerExchange px;
    Geometry geo;
    GtOffsets go;          // host values (argument checks; uploaded by prep_step_kernel unless gt_offsets_dev is set)
    LossScalars sc;
    LossDebug dbg;
    LossWorkspace ws;
    int sumG;              // GT capacity of the call: sizes the per-GT grids and the workspace
};

static int plan_loss(const PaaLossArgs* a, LossPlan* p) {
    if (!a) {
        set_error("null PaaLossArgs");
        return PAA_ERR_BAD_ARGUMENT;
    }
    int rc = build_geometry(a->num_images, a->num_levels, a->num_classes, a->anchors_per_loc,
                            a->anchor_image_stride, a->levels, a->use_iou_pred != 0, a->head_layout, &p->geo);
    if (rc) return rc;
    if (a->topk < 1 || a->topk > 32 || a->num_levels * a->topk > PAA_MAX_CANDIDATES) {
        set_error("topk=%d unsupported (1..32, num_levels*topk <= %d)", a->topk, PAA_MAX_CANDIDATES);
        return PAA_ERR_UNSUPPORTED;
    }
    if (a->num_classes > 1024) {
        set_error("num_classes=%d unsupported (<= 1024)", a->num_classes);
        return PAA_ERR_UNSUPPORTED;
    }
    if (a->world_size < 1) {
        set_error("world_size=%d", a->world_size);
        return PAA_ERR_BAD_ARGUMENT;
    }
    if (a->gt_offsets[0] != 0) {
        set_error("gt_offsets[0] must be 0");
        return PAA_ERR_BAD_ARGUMENT;
    }
    for (int i = 0; i < a->num_images; ++i) {
        p->go.v[i] = a->gt_offsets[i];
        if (a->gt_offsets[i + 1] <= a->gt_offsets[i]) {
            // matcher.py:53-58: "No ground-truth boxes available for one of the images during training"
            set_error("No ground-truth boxes available for one of the images during training (image %d)", i);
            return PAA_ERR_EMPTY_TARGET;
        }
    }
    p->go.v[a->num_images] = a->gt_offsets[a->num_images];
    memset(p->go.by_load, 0, sizeof(p->go.by_load));       // the order by load is worked out on the device
    // Capacities.  Without gt_offsets_dev the host offsets ARE the step's offsets; with it they only describe the
    // batch the call was planned on, and everything sized here must hold for every batch the caller will put into
    // gt_offsets_dev / gt_boxes / gt_labels before replaying a graph captured around the call.
    int gmax = 1;
    for (int i = 0; i < a->num_images; ++i)
        gmax = gmax > a->gt_offsets[i + 1] - a->gt_offsets[i] ? gmax : a->gt_offsets[i + 1] - a->gt_offsets[i];
    p->sumG = a->gt_offsets[a->num_images];
    if (a->gt_offsets_dev) {
        if (a->gt_capacity < p->sumG || a->gt_per_image_capacity < gmax) {
            set_error("gt_offsets_dev: gt_capacity=%d / gt_per_image_capacity=%d below this batch's %d / %d",
                      a->gt_capacity, a->gt_per_image_capacity, p->sumG, gmax);
            return PAA_ERR_BAD_ARGUMENT;
        }
        p->sumG = a->gt_capacity;
        gmax = a->gt_per_image_capacity;
    }
    p->sc.gt_capacity = p->sumG;
    p->sc.gt_parts = gt_parts_for(gmax);
    if (!a->gt_boxes || !a->gt_labels || !a->workspace || !a->normalisers || !a->losses) {
        set_error("null gt_boxes / gt_labels / workspace / normalisers / losses");
        return PAA_ERR_BAD_ARGUMENT;
    }
    p->ws = carve_loss_workspace(a->workspace, a->num_images, p->geo.A, p->sumG, p->geo.tiles_per_image,
                                 loss_grid_blocks(a->num_images, p->geo.tiles_per_image), a->num_levels);
    if (p->ws.total_bytes > a->workspace_bytes) {
        set_error("workspace too small: need %zu bytes, got %zu", p->ws.total_bytes, a->workspace_bytes);
        return PAA_ERR_WORKSPACE;
    }
    if ((reinterpret_cast<uintptr_t>(a->workspace) & 255u) != 0) {
        set_error("workspace must be 256-byte aligned");
        return PAA_ERR_BAD_ARGUMENT;
    }
    p->sc.gamma = a->gamma;
    p->sc.alpha = a->alpha;
    p->sc.iou_threshold = a->iou_threshold;
    p->sc.reg_loss_weight = a->reg_loss_weight;
    p->sc.iou_loss_weight = a->iou_loss_weight;
    p->sc.topk = a->topk;
    p->sc.use_iou_pred = a->use_iou_pred;
    p->sc.world_size = a->world_size;
    p->sc.flavour = a->loss_flavour;
    if (a->loss_flavour != PAA_LOSS_PAA && a->loss_flavour != PAA_LOSS_ATSS &&
        a->loss_flavour != PAA_LOSS_RETINANET && a->loss_flavour != PAA_LOSS_FCOS) {
        set_error("loss_flavour=%d", a->loss_flavour);
        return PAA_ERR_BAD_ARGUMENT;
    }
    p->sc.num_images = a->num_images;
    p->sc.bg_threshold = a->bg_iou_threshold;
    for (int k = 0; k < 4; ++k) p->sc.code_w[k] = a->box_code_weights[k];
    p->sc.beta = a->smooth_l1_beta;
    p->sc.reg_norm_weight = a->reg_norm_weight;
    p->sc.atss_type = a->loss_flavour == PAA_LOSS_ATSS ? a->atss_positive_type : PAA_ATSS_POSITIVE_ATSS;
    if (a->loss_flavour == PAA_LOSS_ATSS) {
        if (a->atss_positive_type < PAA_ATSS_POSITIVE_ATSS || a->atss_positive_type > PAA_ATSS_POSITIVE_IOU) {
            set_error("POSITIVE_TYPE %d not implemented", a->atss_positive_type);      // atss/loss.py:227-228
            return PAA_ERR_UNSUPPORTED;
        }
        if (a->atss_positive_type == PAA_ATSS_POSITIVE_SSC && a->num_levels > 5) {
            set_error("ATSS POSITIVE_TYPE 'SSC' lists five size ranges (atss/loss.py:89), got %d levels", a->num_levels);
            return PAA_ERR_BAD_ARGUMENT;
        }
        if (a->atss_positive_type == PAA_ATSS_POSITIVE_IOU && !(a->bg_iou_threshold <= a->iou_threshold)) {
            set_error("ATSS POSITIVE_TYPE 'IoU': bg_iou_threshold=%g > iou_threshold=%g", a->bg_iou_threshold,
                      a->iou_threshold);
            return PAA_ERR_BAD_ARGUMENT;
        }
    }
    p->sc.fcos_iou_type = a->fcos_iou_loss_type;
    p->sc.fcos_norm = a->fcos_norm_reg_targets;
    for (int l = 0; l < PAA_MAX_LEVELS; ++l) {
        p->sc.fcos_stride[l] = a->fcos_strides[l];
        // `center - stride * radius` with a Python float product rounded to float32 (fcos/loss.py:74-78)
        p->sc.fcos_radius[l] = (float)((double)a->fcos_strides[l] * (double)a->fcos_center_radius);
    }
    if (a->loss_flavour == PAA_LOSS_FCOS) {
        if (!a->use_iou_pred || a->anchors_per_loc != 1 || a->num_levels > 5) {
            // object_sizes_of_interest lists five levels (fcos/loss.py:106-112)
            set_error("PAA_LOSS_FCOS needs the centerness map (use_iou_pred), anchors_per_loc = 1 and <= 5 levels");
            return PAA_ERR_BAD_ARGUMENT;
        }
        if (a->fcos_iou_loss_type < PAA_IOU_LOSS_IOU || a->fcos_iou_loss_type > PAA_IOU_LOSS_GIOU ||
            !(a->fcos_center_radius >= 0.0f)) {
            set_error("PAA_LOSS_FCOS: fcos_iou_loss_type=%d / fcos_center_radius=%g", a->fcos_iou_loss_type,
                      a->fcos_center_radius);
            return PAA_ERR_BAD_ARGUMENT;
        }
        for (int l = 0; l < a->num_levels; ++l)
            if (!(a->fcos_strides[l] > 0.0f)) {
                set_error("PAA_LOSS_FCOS: fcos_strides[%d]=%g", l, a->fcos_strides[l]);
                return PAA_ERR_BAD_ARGUMENT;
            }
    }
    if (a->loss_flavour == PAA_LOSS_RETINANET) {
        if (a->use_iou_pred) {
            set_error("PAA_LOSS_RETINANET has no third head: use_iou_pred must be 0");
            return PAA_ERR_BAD_ARGUMENT;
        }
        if (!(a->smooth_l1_beta > 0.0f) || !(a->bg_iou_threshold <= a->iou_threshold)) {
            // matcher.py:35 asserts low <= high
            set_error("PAA_LOSS_RETINANET: smooth_l1_beta=%g must be positive and bg_iou_threshold=%g <= "
                      "iou_threshold=%g", a->smooth_l1_beta, a->bg_iou_threshold, a->iou_threshold);
            return PAA_ERR_BAD_ARGUMENT;
        }
    }
    {
        // test hook: a smaller pool forces the overflow path of the candidate selection
        const char* e = getenv("PAA_SEG_CAP");
        int cap = e ? atoi(e) : kSegCap;
        p->sc.seg_cap = cap < 1 ? 1 : (cap > kSegCap ? kSegCap : cap);
    }
    memset(&p->px, 0, sizeof(p->px));
    if (a->peer_norm[0] != nullptr && a->world_size > 1 && a->loss_flavour != PAA_LOSS_RETINANET) {
        if (a->world_size > PAA_MAX_PEERS || a->rank < 0 || a->rank >= a->world_size) {
            set_error("peer exchange: rank %d / world_size %d unsupported (<= %d ranks)", a->rank, a->world_size,
                      PAA_MAX_PEERS);
            return PAA_ERR_UNSUPPORTED;
        }
        for (int r = 0; r < a->world_size; ++r) {
            if (!a->peer_norm[r]) {
                set_error("peer exchange: peer_norm[%d] is null", r);
                return PAA_ERR_BAD_ARGUMENT;
            }
            p->px.buf[r] = a->peer_norm[r];
        }
        p->px.rank = a->rank;
        p->px.world = a->world_size;
        p->px.status = a->peer_status;
        p->px.timeout_ns = a->peer_timeout_s > 0.0f ? (unsigned long long)((double)a->peer_timeout_s * 1e9) : 0ull;
    }
    p->dbg.matched_idx = a->dbg_matched_idx;
    p->dbg.iou_labels = a->dbg_iou_labels;
    p->dbg.combined_loss = a->dbg_combined_loss;
    p->dbg.cand_idx = a->dbg_cand_idx;
    p->dbg.cand_cnt = a->dbg_cand_cnt;
    p->dbg.num_pos = a->dbg_num_pos;
    p->dbg.gmm = a->dbg_gmm;
    p->dbg.paa_labels = a->dbg_paa_labels;
    return 0;
}

}  // namespace paa

using namespace paa;

#ifdef PAA_TRACE
namespace paa {
int trace_set_assign(unsigned long long* p);
int trace_set_loss(unsigned long long* p);
int trace_set_post(unsigned long long* p);
}
#endif

extern "C" {

int paa_abi_version(void) { return PAA_ABI_VERSION; }

const char* paa_last_error(void) { return g_error; }

size_t paa_loss_workspace_bytes(int num_images, int anchors_per_image, int num_gt_total, int num_levels,
                                int topk) {
    (void)topk;
    if (num_images < 1 || anchors_per_image < 1 || num_levels < 1) return 0;
    const int tiles = tiles_upper_bound(anchors_per_image, num_levels);
    LossWorkspace w = carve_loss_workspace(nullptr, num_images, anchors_per_image, num_gt_total, tiles,
                                           loss_grid_blocks(num_images, tiles), num_levels);
    return w.total_bytes;
}

int paa_assign(const PaaLossArgs* args, void* stream_) {
    LossPlan p;
    int rc = plan_loss(args, &p);
    if (rc) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    // clears the zeroed prefix and the coarse levels' best-GT keys (merged with atomicMax when their GT list is cut
    // into parts), uploads the GT ranges
    if ((rc = launch_prep_step(p.geo, p.go, args->gt_offsets_dev, p.sc, p.ws, args->workspace, true, false, stream)))
        return rc;
    if ((rc = launch_iou_match(p.geo, args->gt_boxes, p.sc, p.ws, stream))) return rc;
    const float* score_src = args->teacher_combined_loss ? args->teacher_combined_loss : p.ws.score;
    if ((rc = launch_match_score(p.geo, args->gt_boxes, args->gt_labels, p.sc, p.ws,
                                 args->teacher_combined_loss, p.dbg, stream)))
        return rc;
    if ((rc = launch_select_gmm(p.geo, args->gt_boxes, args->gt_labels, p.sc, p.ws, score_src,
                                args->normalisers, p.px, p.dbg, stream)))
        return rc;
    if (args->dbg_paa_labels)
        PAA_CUDA_CHECK(cudaMemcpyAsync(args->dbg_paa_labels, p.ws.paa_label,
                                       sizeof(int) * (size_t)args->num_images * p.geo.A,
                                       cudaMemcpyDeviceToDevice, stream));
    return 0;
}

int paa_atss_assign(const PaaLossArgs* args, void* stream_) {
    LossPlan p;
    int rc = plan_loss(args, &p);
    if (rc) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (p.sc.atss_type == PAA_ATSS_POSITIVE_SSC) {
        if ((rc = launch_prep_step(p.geo, p.go, args->gt_offsets_dev, p.sc, p.ws, args->workspace, false, false,
                                   stream)))
            return rc;
        rc = launch_fcos_assign(p.geo, args->gt_boxes, args->gt_labels, p.sc, p.ws, args->normalisers, p.px,
                                p.dbg, stream, /*ssc=*/true);
    } else if (p.sc.atss_type == PAA_ATSS_POSITIVE_IOU) {
        if ((rc = launch_prep_step(p.geo, p.go, args->gt_offsets_dev, p.sc, p.ws, args->workspace, true, false,
                                   stream)))
            return rc;
        rc = launch_retinanet_assign(p.geo, args->gt_boxes, args->gt_labels, p.sc, p.ws, args->normalisers,
                                     p.dbg, stream, /*atss_iou=*/true, &p.px);
    } else {
        for (int l = 0; l < p.geo.num_levels; ++l)
            if (p.geo.lv[l].n_anchor < args->topk) {
                // torch.topk(k) on fewer than k anchors: "selected index k out of range" (atss/loss.py:159)
                set_error("selected index k out of range: level %d has %d anchors, TOPK is %d", l,
                          p.geo.lv[l].n_anchor, args->topk);
                return PAA_ERR_UNSUPPORTED;
            }
        if ((rc = launch_prep_step(p.geo, p.go, args->gt_offsets_dev, p.sc, p.ws, args->workspace, false, true,
                                   stream)))
            return rc;
        rc = launch_atss_assign(p.geo, args->gt_boxes, args->gt_labels, p.sc, p.ws, args->normalisers,
                                p.px, p.dbg, stream);
    }
    if (rc) return rc;
    if (args->dbg_paa_labels)
        PAA_CUDA_CHECK(cudaMemcpyAsync(args->dbg_paa_labels, p.ws.paa_label,
                                       sizeof(int) * (size_t)args->num_images * p.geo.A,
                                       cudaMemcpyDeviceToDevice, stream));
    return 0;
}

int paa_retinanet_assign(const PaaLossArgs* args, void* stream_) {
    LossPlan p;
    int rc = plan_loss(args, &p);
    if (rc) return rc;
    if (args->loss_flavour != PAA_LOSS_RETINANET) {
        set_error("paa_retinanet_assign needs loss_flavour = PAA_LOSS_RETINANET");
        return PAA_ERR_BAD_ARGUMENT;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if ((rc = launch_prep_step(p.geo, p.go, args->gt_offsets_dev, p.sc, p.ws, args->workspace, true, false, stream)))
        return rc;
    if ((rc = launch_retinanet_assign(p.geo, args->gt_boxes, args->gt_labels, p.sc, p.ws, args->normalisers,
                                      p.dbg, stream)))
        return rc;
    if (args->dbg_paa_labels)
        PAA_CUDA_CHECK(cudaMemcpyAsync(args->dbg_paa_labels, p.ws.paa_label,
                                       sizeof(int) * (size_t)args->num_images * p.geo.A,
                                       cudaMemcpyDeviceToDevice, stream));
    return 0;
}

int paa_fcos_assign(const PaaLossArgs* args, void* stream_) {
    LossPlan p;
    int rc = plan_loss(args, &p);
    if (rc) return rc;
    if (args->loss_flavour != PAA_LOSS_FCOS) {
        set_error("paa_fcos_assign needs loss_flavour = PAA_LOSS_FCOS");
        return PAA_ERR_BAD_ARGUMENT;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if ((rc = launch_prep_step(p.geo, p.go, args->gt_offsets_dev, p.sc, p.ws, args->workspace, false, false, stream)))
        return rc;
    if ((rc = launch_fcos_assign(p.geo, args->gt_boxes, args->gt_labels, p.sc, p.ws, args->normalisers, p.px,
                                 p.dbg, stream)))
        return rc;
    if (args->dbg_paa_labels)
        PAA_CUDA_CHECK(cudaMemcpyAsync(args->dbg_paa_labels, p.ws.paa_label,
                                       sizeof(int) * (size_t)args->num_images * p.geo.A,
                                       cudaMemcpyDeviceToDevice, stream));
    return 0;
}

int paa_loss(const PaaLossArgs* args, void* stream_) {
    LossPlan p;
    int rc = plan_loss(args, &p);
    if (rc) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (p.px.world > 1 && (rc = launch_norm_wait(p.px, args->normalisers, stream))) return rc;
    bool write_grads = false;
    for (int l = 0; l < args->num_levels; ++l)
        write_grads = write_grads || args->levels[l].grad_box_cls || args->levels[l].grad_box_regression ||
                      args->levels[l].grad_iou_pred;
    return launch_final_loss(p.geo, args->gt_boxes, args->gt_labels, p.sc, p.ws, args->normalisers, args->grad_losses,
                             args->losses, write_grads, stream);
}

int paa_assign_loss(const PaaLossArgs* args, void* stream) {
    int rc = paa_assign(args, stream);
    if (rc) return rc;
    return paa_loss(args, stream);
}

int paa_rescale_grads(const PaaLossArgs* args, const float* old_grad_losses, const float* new_grad_losses,
                      void* stream_) {
    LossPlan p;
    int rc = plan_loss(args, &p);
    if (rc) return rc;
    if (!old_grad_losses || !new_grad_losses) {
        set_error("null grad_losses");
        return PAA_ERR_BAD_ARGUMENT;
    }
    return launch_rescale_grads(p.geo, old_grad_losses, new_grad_losses, static_cast<cudaStream_t>(stream_));
}

int paa_sigmoid_focal_loss_forward(const float* logits, const int32_t* targets, int n, int num_classes,
                                   float gamma, float alpha, float* losses, void* stream) {
    if (n < 0 || num_classes < 1 || (n > 0 && (!logits || !targets || !losses))) {
        set_error("bad arguments to paa_sigmoid_focal_loss_forward");
        return PAA_ERR_BAD_ARGUMENT;
    }
    return launch_focal_forward(logits, targets, n, num_classes, gamma, alpha, losses,
                                static_cast<cudaStream_t>(stream));
}

int paa_sigmoid_focal_loss_backward(const float* logits, const int32_t* targets, const float* d_losses, int n,
                                    int num_classes, float gamma, float alpha, float* d_logits, void* stream) {
    if (n < 0 || num_classes < 1 || (n > 0 && (!logits || !targets || !d_losses || !d_logits))) {
        set_error("bad arguments to paa_sigmoid_focal_loss_backward");
        return PAA_ERR_BAD_ARGUMENT;
    }
    return launch_focal_backward(logits, targets, d_losses, n, num_classes, gamma, alpha, d_logits,
                                 static_cast<cudaStream_t>(stream));
}

size_t paa_postprocess_workspace_bytes(int num_images, int anchors_per_image, int num_classes, int num_levels,
                                       int pre_nms_top_n) {
    return post_workspace_bytes(num_images, anchors_per_image, num_classes, num_levels, pre_nms_top_n);
}

int paa_rpn_loss(const PaaRpnArgs* a, void* stream) {
    if (!a) {
        set_error("null PaaRpnArgs");
        return PAA_ERR_BAD_ARGUMENT;
    }
    Geometry geo;
    int rc = build_geometry(a->num_images, a->num_levels, 1, a->anchors_per_loc, a->anchor_image_stride, a->levels,
                            false, a->head_layout, &geo);
    if (rc) return rc;
    if (a->n_pos < 0 || a->n_neg < 0 || !a->losses || !a->gt_boxes || !a->matched_idx ||
        ((a->n_pos + a->n_neg) > 0 && !a->sampled)) {
        set_error("paa_rpn_loss: null pointer or negative sample size (n_pos=%d, n_neg=%d)", a->n_pos, a->n_neg);
        return PAA_ERR_BAD_ARGUMENT;
    }
    if (!(a->smooth_l1_beta > 0.0f)) {
        set_error("paa_rpn_loss: smooth_l1_beta=%g must be positive", (double)a->smooth_l1_beta);
        return PAA_ERR_BAD_ARGUMENT;
    }
    GtOffsets go;
    memset(&go, 0, sizeof(go));
    for (int i = 0; i <= a->num_images; ++i) go.v[i] = a->gt_offsets[i];
    return launch_rpn_loss(geo, go, a->gt_boxes, a->matched_idx, reinterpret_cast<const long long*>(a->sampled),
                           a->n_pos, a->n_neg, a->box_code_weights, a->smooth_l1_beta, a->grad_losses, a->losses,
                           static_cast<cudaStream_t>(stream));
}

int paa_postprocess(const PaaPostArgs* args, void* stream) {
    if (!args) {
        set_error("null PaaPostArgs");
        return PAA_ERR_BAD_ARGUMENT;
    }
    Geometry geo;
    int rc = build_geometry(args->num_images, args->num_levels, args->num_classes, args->anchors_per_loc,
                            args->anchor_image_stride, args->levels, false, args->head_layout, &geo);
    if (rc) return rc;
    return run_postprocess(geo, args, static_cast<cudaStream_t>(stream));
}

int paa_selftest_roots(const float* x, int n, float* out, void* stream) {
    if (n < 0 || (n > 0 && (!x || !out))) {
        set_error("bad arguments to paa_selftest_roots");
        return PAA_ERR_BAD_ARGUMENT;
    }
    return launch_selftest_roots(x, n, out, static_cast<cudaStream_t>(stream));
}

int paa_kernel_timing_begin(int kernel_id) {
    g_timed_kernel = kernel_id;
    g_timed_count = 0;
    return 0;
}

int paa_kernel_timing_end(float* total_ms, int32_t* launches) {
    float total = 0.0f;
    for (int i = 0; i < g_timed_count; ++i) {
        PAA_CUDA_CHECK(cudaEventSynchronize(g_ev1[i]));
        float ms = 0.0f;
        PAA_CUDA_CHECK(cudaEventElapsedTime(&ms, g_ev0[i], g_ev1[i]));
        total += ms;
    }
    if (total_ms) *total_ms = total;
    if (launches) *launches = g_timed_count;
    g_timed_kernel = 0;
    g_timed_count = 0;
    return 0;
}

#ifdef PAA_TRACE
// measurement build only (tools/step_trace.py): device buffer of 4 x 8 timestamps per kernel slot, see common.cuh
int paa_trace_set(unsigned long long* device_buffer) {
    int rc = paa::trace_set_assign(device_buffer);
    if (!rc) rc = paa::trace_set_loss(device_buffer);
    if (!rc) rc = paa::trace_set_post(device_buffer);
    return rc;
}
#endif

size_t paa_ml_nms_workspace_bytes(int n) { return ml_nms_workspace_bytes(n); }

int paa_ml_nms(const float* boxes, const float* scores, const float* labels, int n, float thresh,
               uint8_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes, void* stream) {
    return run_ml_nms(boxes, scores, labels, n, thresh, keep, num_keep, workspace, workspace_bytes,
                      static_cast<cudaStream_t>(stream));
}

size_t paa_box_vote_workspace_bytes(int n) { return box_vote_workspace_bytes(n); }

int paa_box_vote(const float* boxes, const float* scores, const float* labels, int n, int mode, float vote_thresh,
                 float nms_thresh, float soft_score_thresh, int max_detections, float* out_boxes, float* out_scores,
                 int64_t* out_labels, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
    return run_box_vote(boxes, scores, labels, n, mode, vote_thresh, nms_thresh, soft_score_thresh, max_detections,
                        out_boxes, out_scores, reinterpret_cast<long long*>(out_labels), out_count, workspace,
                        workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
