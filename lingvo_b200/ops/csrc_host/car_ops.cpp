#include "car_ops.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <unordered_map>

namespace lbh {
namespace {

struct P2 {
  float x, y;
};

// Corners of the box footprint, counter-clockwise.
void Corners(const float* b, P2 out[4]) {
  const float c = std::cos(b[6]), s = std::sin(b[6]);
  const float hx = b[3] * 0.5f, hy = b[4] * 0.5f;
  const float lx[4] = {hx, -hx, -hx, hx}, ly[4] = {hy, hy, -hy, -hy};
  for (int i = 0; i < 4; ++i) out[i] = {b[0] + lx[i] * c - ly[i] * s, b[1] + lx[i] * s + ly[i] * c};
}

inline float Cross(const P2& o, const P2& a, const P2& b) {
  return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x);
}

// Sutherland–Hodgman: clip polygon `poly` (size n) by the half-plane left of a→b.
int ClipEdge(const P2* poly, int n, const P2& a, const P2& b, P2* out) {
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const P2& cur = poly[i];
    const P2& nxt = poly[(i + 1) % n];
    const float dc = Cross(a, b, cur), dn = Cross(a, b, nxt);
    if (dc >= 0) out[m++] = cur;
    if ((dc >= 0) != (dn >= 0)) {
      const float t = dc / (dc - dn);
      out[m++] = {cur.x + t * (nxt.x - cur.x), cur.y + t * (nxt.y - cur.y)};
    }
  }
  return m;
}

float IntersectionArea(const float* a, const float* b) {
  P2 pa[4], pb[4], buf1[16], buf2[16];
  Corners(a, pa);
  Corners(b, pb);
  int n = 4;
  std::memcpy(buf1, pa, sizeof(pa));
  P2* src = buf1;
  P2* dst = buf2;
  for (int e = 0; e < 4 && n > 0; ++e) {
    n = ClipEdge(src, n, pb[e], pb[(e + 1) % 4], dst);
    std::swap(src, dst);
  }
  float area = 0.f;
  for (int i = 0; i < n; ++i) {
    const P2& p = src[i];
    const P2& q = src[(i + 1) % n];
    area += p.x * q.y - q.x * p.y;
  }
  return std::fabs(area) * 0.5f;
}

float Iou3D(const float* a, const float* b) {
  const float za0 = a[2] - a[5] * 0.5f, za1 = a[2] + a[5] * 0.5f;
  const float zb0 = b[2] - b[5] * 0.5f, zb1 = b[2] + b[5] * 0.5f;
  const float h = std::min(za1, zb1) - std::max(za0, zb0);
  if (h <= 0) return 0.f;
  // cheap reject on circumscribed circles
  const float ra = 0.5f * std::hypot(a[3], a[4]), rb = 0.5f * std::hypot(b[3], b[4]);
  if (std::hypot(a[0] - b[0], a[1] - b[1]) > ra + rb) return 0.f;
  const float inter = IntersectionArea(a, b) * h;
  const float va = a[3] * a[4] * a[5], vb = b[3] * b[4] * b[5];
  const float uni = va + vb - inter;
  return uni > 0 ? inter / uni : 0.f;
}

}  // namespace

void PairwiseIou3D(const float* a, int n, const float* b, int m, float* iou) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) iou[static_cast<size_t>(i) * m + j] = Iou3D(a + 7 * i, b + 7 * j);
}

std::vector<int32_t> NonMaxSuppression3D(const float* boxes, const float* scores, int n,
                                         int num_classes, const std::vector<float>& nms_iou,
                                         const std::vector<float>& score_thresh, int max_boxes) {
  std::vector<int32_t> out(static_cast<size_t>(num_classes) * max_boxes, -1);
  std::vector<int> order(n);
  for (int c = 0; c < num_classes; ++c) {
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
      return scores[static_cast<size_t>(x) * num_classes + c] >
             scores[static_cast<size_t>(y) * num_classes + c];
    });
    std::vector<int> kept;
    for (int idx : order) {
      if (static_cast<int>(kept.size()) >= max_boxes) break;
      if (scores[static_cast<size_t>(idx) * num_classes + c] <= score_thresh[c]) break;
      bool ok = true;
      for (int k : kept) {
        if (Iou3D(boxes + 7 * idx, boxes + 7 * k) > nms_iou[c]) {
          ok = false;
          break;
        }
      }
      if (ok) kept.push_back(idx);
    }
    for (size_t k = 0; k < kept.size(); ++k) out[static_cast<size_t>(c) * max_boxes + k] = kept[k];
  }
  return out;
}

int PointsToPillars(const float* points, int n, int dims, float x0, float x1, float y0, float y1,
                    int nx, int ny, int max_pillars, int k, float* pillar_points,
                    int32_t* pillar_xy, int32_t* pillar_count) {
  std::unordered_map<int64_t, int> slot;
  slot.reserve(static_cast<size_t>(max_pillars) * 2);
  const float sx = nx / (x1 - x0), sy = ny / (y1 - y0);
  int used = 0;
  for (int i = 0; i < n; ++i) {
    const float* p = points + static_cast<size_t>(i) * dims;
    if (!(p[0] >= x0 && p[0] < x1 && p[1] >= y0 && p[1] < y1)) continue;
    const int gx = std::min(static_cast<int>((p[0] - x0) * sx), nx - 1);
    const int gy = std::min(static_cast<int>((p[1] - y0) * sy), ny - 1);
    const int64_t key = static_cast<int64_t>(gx) * ny + gy;
    auto it = slot.find(key);
    int s;
    if (it == slot.end()) {
      if (used >= max_pillars) continue;
      s = used++;
      slot.emplace(key, s);
      pillar_xy[2 * s] = gx;
      pillar_xy[2 * s + 1] = gy;
      pillar_count[s] = 0;
    } else {
      s = it->second;
    }
    if (pillar_count[s] >= k) continue;
    std::memcpy(pillar_points + (static_cast<size_t>(s) * k + pillar_count[s]) * dims, p,
                sizeof(float) * dims);
    ++pillar_count[s];
  }
  return used;
}

std::vector<int32_t> FarthestPointSample(const float* xyz, int n, int k) {
  std::vector<int32_t> out;
  if (n <= 0) return out;
  std::vector<float> dist(n, std::numeric_limits<float>::max());
  int cur = 0;
  for (int s = 0; s < k && s < n; ++s) {
    out.push_back(cur);
    const float* c = xyz + 3 * cur;
    int best = 0;
    float best_d = -1.f;
    for (int i = 0; i < n; ++i) {
      const float* p = xyz + 3 * i;
      const float d = (p[0] - c[0]) * (p[0] - c[0]) + (p[1] - c[1]) * (p[1] - c[1]) +
                      (p[2] - c[2]) * (p[2] - c[2]);
      dist[i] = std::min(dist[i], d);
      if (dist[i] > best_d) {
        best_d = dist[i];
        best = i;
      }
    }
    cur = best;
  }
  return out;
}

ApResult AveragePrecision3D(float iou_threshold, const float* gt_bbox, const int32_t* gt_imageid,
                            const int32_t* gt_ignore, int n, const float* pd_bbox,
                            const int32_t* pd_imageid, const int32_t* pd_ignore,
                            const float* pd_score, int m, int num_recall_points, bool kitti) {
  ApResult res;
  res.score_and_hit.assign(static_cast<size_t>(m) * 2, 0.f);
  res.precision_recall.assign(static_cast<size_t>(num_recall_points) * 2, 0.f);
  // ground truth grouped by image
  std::unordered_map<int32_t, std::vector<int>> by_image;
  int num_valid_gt = 0;
  for (int i = 0; i < n; ++i) {
    by_image[gt_imageid[i]].push_back(i);
    if (gt_ignore[i] == 0) ++num_valid_gt;
  }
  std::vector<int> order(m);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(),
                   [&](int a, int b) { return pd_score[a] > pd_score[b]; });
  std::vector<char> used(n, 0);
  // running (tp, fp) after each counted prediction
  std::vector<float> tps, fps;
  int tp = 0, fp = 0;
  for (int rank = 0; rank < m; ++rank) {
    const int j = order[rank];
    res.score_and_hit[2 * j] = pd_score[j];
    int best = -1;
    float best_iou = iou_threshold;
    bool touches_ignore_all = false;
    auto it = by_image.find(pd_imageid[j]);
    if (it != by_image.end()) {
      for (int g : it->second) {
        const float iou = Iou3D(pd_bbox + 7 * j, gt_bbox + 7 * g);
        if (iou < iou_threshold) continue;
        if (gt_ignore[g] == 2) {
          touches_ignore_all = true;
          continue;
        }
        if (used[g]) continue;
        // prefer real boxes over "ignore first match" boxes at equal footing, then IoU
        const bool better = best < 0 || (gt_ignore[best] != 0 && gt_ignore[g] == 0) ||
                            ((gt_ignore[best] != 0) == (gt_ignore[g] != 0) && iou > best_iou);
        if (better) {
          best = g;
          best_iou = iou;
        }
      }
    }
    if (best >= 0) {
      used[best] = 1;
      if (gt_ignore[best] == 0) {
        ++tp;
        res.score_and_hit[2 * j + 1] = 1.f;
        tps.push_back(static_cast<float>(tp));
        fps.push_back(static_cast<float>(fp));
      }
      continue;   // matched an "ignore first" box: neither TP nor FP
    }
    if (touches_ignore_all || pd_ignore[j] == 1) continue;
    ++fp;
    tps.push_back(static_cast<float>(tp));
    fps.push_back(static_cast<float>(fp));
  }
  if (num_valid_gt == 0 || tps.empty()) return res;
  const size_t k = tps.size();
  std::vector<float> prec(k), rec(k);
  for (size_t i = 0; i < k; ++i) {
    prec[i] = tps[i] / (tps[i] + fps[i]);
    rec[i] = tps[i] / static_cast<float>(num_valid_gt);
  }
  // interpolated precision: max precision at recall >= r
  for (size_t i = k - 1; i-- > 0;) prec[i] = std::max(prec[i], prec[i + 1]);
  auto prec_at = [&](float r) {
    auto lo = std::lower_bound(rec.begin(), rec.end(), r - 1e-7f);
    return lo == rec.end() ? 0.f : prec[lo - rec.begin()];
  };
  if (kitti) {
    double sum = 0;
    for (int i = 0; i <= num_recall_points; ++i)
      sum += prec_at(static_cast<float>(i) / num_recall_points);
    res.average_precision = static_cast<float>(sum / (num_recall_points + 1));
  } else {
    double area = 0, prev_r = 0;
    for (size_t i = 0; i < k; ++i) {
      if (i + 1 < k && rec[i + 1] == rec[i]) continue;
      area += (rec[i] - prev_r) * prec[i];
      prev_r = rec[i];
    }
    res.average_precision = static_cast<float>(area);
  }
  for (int i = 0; i < num_recall_points; ++i) {       // recall descending
    const float r = static_cast<float>(num_recall_points - i) / num_recall_points;
    res.precision_recall[2 * i] = prec_at(r);
    res.precision_recall[2 * i + 1] = prec_at(r) > 0.f ? r : 0.f;
  }
  return res;
}

}  // namespace lbh
