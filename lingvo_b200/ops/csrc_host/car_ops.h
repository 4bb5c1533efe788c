// 3-D detection geometry (SURVEY §2.7 "car ops"): rotated-box IoU, 3-D NMS,
// point → pillar grid assignment, farthest-point sampling. Native re-designs of
// `tasks/car/ops/{pairwise_iou_op,nms_3d_op,point_grid_op,sampling_ops}.cc`.
#pragma once

#include <cstdint>
#include <vector>

namespace lbh {

// Boxes are 7-DOF: (x, y, z, dx, dy, dz, heading about +z).
// iou[n, m] = volume(A_n ∩ B_m) / volume(A_n ∪ B_m): exact convex-polygon clipping
// in the ground plane × overlap of the z extents.
void PairwiseIou3D(const float* a, int n, const float* b, int m, float* iou);

// Greedy per-class NMS. scores [n, num_classes]; a box is kept for class c if its
// score > score_thresh[c] and its IoU with every already-kept box of that class is
// <= nms_iou_thresh[c]. Returns for each class up to max_boxes_per_class indices
// (padded with -1) in decreasing score order.
std::vector<int32_t> NonMaxSuppression3D(const float* boxes, const float* scores, int n,
                                         int num_classes, const std::vector<float>& nms_iou_thresh,
                                         const std::vector<float>& score_thresh,
                                         int max_boxes_per_class);

// Assigns points to an (nx, ny) pillar grid over [x0,x1)×[y0,y1); keeps at most
// `max_pillars` non-empty pillars and `points_per_pillar` points each (first come).
// Outputs: pillar_points [P, K, dims] (zero padded), pillar_xy [P, 2] grid indices,
// pillar_count [P]; returns number of occupied pillars.
int PointsToPillars(const float* points, int n, int dims, float x0, float x1, float y0, float y1,
                    int nx, int ny, int max_pillars, int points_per_pillar, float* pillar_points,
                    int32_t* pillar_xy, int32_t* pillar_count);

// Farthest-point sampling: indices of `k` points, starting from point 0.
std::vector<int32_t> FarthestPointSample(const float* xyz, int n, int k);

// Average precision of 3-D detections of ONE class (re-design of
// `average_precision_3d_op.cc`). Predictions are visited by decreasing score; each is
// matched to the unmatched ground-truth box of the same image with the highest
// IoU >= iou_threshold. Ground-truth `ignore`: 0 = normal, 1 = ignore the first match
// (matched prediction is neither TP nor FP, box then counts as consumed), 2 = ignore every
// match. Prediction `ignore` 1 = the prediction never counts as FP (KITTI "too small").
// algorithm "KITTI": mean of the interpolated precision at num_recall_points+1 equally
// spaced recalls; "VOC": area under the interpolated PR curve.
struct ApResult {
  float average_precision = 0.f;
  std::vector<float> precision_recall;   // [num_recall_points, 2] (precision, recall), recall descending
  std::vector<float> score_and_hit;      // [M, 2]
};
ApResult AveragePrecision3D(float iou_threshold, const float* gt_bbox, const int32_t* gt_imageid,
                            const int32_t* gt_ignore, int n, const float* pd_bbox,
                            const int32_t* pd_imageid, const int32_t* pd_ignore,
                            const float* pd_score, int m, int num_recall_points, bool kitti);

// Same protocol for axis-aligned image boxes (ymin, xmin, ymax, xmax) — re-design of
// `image_metrics.cc`.
float Iou2D(const float* a, const float* b);
ApResult AveragePrecision2D(float iou_threshold, const float* gt_bbox, const int32_t* gt_imageid,
                            const int32_t* gt_ignore, int n, const float* pd_bbox,
                            const int32_t* pd_imageid, const int32_t* pd_ignore,
                            const float* pd_score, int m, int num_recall_points, bool kitti);

// Point-cloud sampling for one scene (re-design of `ps_utils.cc` / `sampling_ops.cc`):
// picks `num_centers` centres (uniform without replacement, or farthest-point; only
// un-padded points with z in [center_z_min, center_z_max] qualify) and for each centre up
// to `num_neighbors` un-padded points within `max_dist` (the `closest` ones, or a uniform
// sample). The first `num_seeded` points count as already-chosen centres for the
// farthest-point criterion but are never emitted as centres or neighbours.
// use_hash: answer ball queries from a uniform grid of edge max_dist instead of a scan
// (picked automatically for large scenes).
struct SampleOptions {
  bool farthest = true;
  bool closest = true;
  bool use_hash = false;
  int num_centers = 0;
  int num_neighbors = 0;
  float center_z_min = -3.4e38f, center_z_max = 3.4e38f;
  float max_dist = 3.4e38f;
  int64_t seed = -1;
};
struct SampleResult {
  std::vector<int32_t> center;           // [M]
  std::vector<float> center_padding;     // [M]
  std::vector<int32_t> indices;          // [M, K]
  std::vector<float> indices_padding;    // [M, K]
};
SampleResult SamplePoints(const float* points, const float* padding, int n, int dims,
                          int num_seeded, const SampleOptions& o);

}  // namespace lbh
