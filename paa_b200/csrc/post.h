// Internal declarations of the post-processing path.  Not a public header.
#pragma once
#include "common.cuh"

namespace paa {

size_t post_workspace_bytes(int num_images, int anchors_per_image, int num_classes, int num_levels,
                            int pre_nms_top_n);
int run_postprocess(const Geometry& geo, const PaaPostArgs* args, cudaStream_t stream);

size_t ml_nms_workspace_bytes(int n);
int run_ml_nms(const float* boxes, const float* scores, const float* labels, int n, float thresh,
               uint8_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes, cudaStream_t stream);

size_t box_vote_workspace_bytes(int n);
int run_box_vote(const float* boxes, const float* scores, const float* labels, int n, int mode, float vote_thresh,
                 float nms_thresh, float soft_score_thresh, int max_detections, float* out_boxes, float* out_scores,
                 long long* out_labels, int32_t* out_count, void* workspace, size_t workspace_bytes,
                 cudaStream_t stream);

}  // namespace paa
