#include "car_ops.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <random>
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

namespace {

// Shared matcher / PR-curve integration; `iou(j, g)` scores prediction j against box g.
template <class IouFn>
ApResult AveragePrecisionImpl(float iou_threshold, IouFn iou_of, const int32_t* gt_imageid,
                              const int32_t* gt_ignore, int n, const int32_t* pd_imageid,
                              const int32_t* pd_ignore, const float* pd_score, int m,
                              int num_recall_points, bool kitti) {
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
        const float iou = iou_of(j, g);
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
      if (i > 0 && rec[i] == rec[i - 1]) continue;   // first entry at a recall has the max precision
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

}  // namespace

ApResult AveragePrecision3D(float iou_threshold, const float* gt_bbox, const int32_t* gt_imageid,
                            const int32_t* gt_ignore, int n, const float* pd_bbox,
                            const int32_t* pd_imageid, const int32_t* pd_ignore,
                            const float* pd_score, int m, int num_recall_points, bool kitti) {
  return AveragePrecisionImpl(
      iou_threshold, [&](int j, int g) { return Iou3D(pd_bbox + 7 * j, gt_bbox + 7 * g); },
      gt_imageid, gt_ignore, n, pd_imageid, pd_ignore, pd_score, m, num_recall_points, kitti);
}

float Iou2D(const float* a, const float* b) {
  const float ih = std::min(a[2], b[2]) - std::max(a[0], b[0]);
  const float iw = std::min(a[3], b[3]) - std::max(a[1], b[1]);
  if (ih <= 0.f || iw <= 0.f) return 0.f;
  const float inter = ih * iw;
  const float ua = std::max(a[2] - a[0], 0.f) * std::max(a[3] - a[1], 0.f) +
                   std::max(b[2] - b[0], 0.f) * std::max(b[3] - b[1], 0.f) - inter;
  return ua > 0.f ? inter / ua : 0.f;
}

ApResult AveragePrecision2D(float iou_threshold, const float* gt_bbox, const int32_t* gt_imageid,
                            const int32_t* gt_ignore, int n, const float* pd_bbox,
                            const int32_t* pd_imageid, const int32_t* pd_ignore,
                            const float* pd_score, int m, int num_recall_points, bool kitti) {
  return AveragePrecisionImpl(
      iou_threshold, [&](int j, int g) { return Iou2D(pd_bbox + 4 * j, gt_bbox + 4 * g); },
      gt_imageid, gt_ignore, n, pd_imageid, pd_ignore, pd_score, m, num_recall_points, kitti);
}

// ------------------------------------------------------------- point sampling ----
namespace {

inline float Dist2(const float* a, const float* b) {
  const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  return dx * dx + dy * dy + dz * dz;
}

// Uniform spatial hash over the valid points: cell edge = max_dist, so a ball query only
// has to visit the 27 cells around the centre.
class CellGrid {
 public:
  CellGrid(const float* pts, int dims, const std::vector<int>& ids, float cell)
      : pts_(pts), dims_(dims), inv_(1.f / cell) {
    cells_.reserve(ids.size());
    for (int i : ids) cells_[Key(pts_ + static_cast<size_t>(i) * dims_, 0, 0, 0)].push_back(i);
  }
  template <class F>
  void ForEachNear(const float* c, F&& f) const {
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = -1; dz <= 1; ++dz) {
          auto it = cells_.find(Key(c, dx, dy, dz));
          if (it == cells_.end()) continue;
          for (int i : it->second) f(i);
        }
  }

 private:
  uint64_t Key(const float* p, int dx, int dy, int dz) const {
    const int64_t x = static_cast<int64_t>(std::floor(p[0] * inv_)) + dx;
    const int64_t y = static_cast<int64_t>(std::floor(p[1] * inv_)) + dy;
    const int64_t z = static_cast<int64_t>(std::floor(p[2] * inv_)) + dz;
    return (static_cast<uint64_t>(x & 0x1FFFFF) << 42) | (static_cast<uint64_t>(y & 0x1FFFFF) << 21) |
           static_cast<uint64_t>(z & 0x1FFFFF);
  }
  const float* pts_;
  int dims_;
  float inv_;
  std::unordered_map<uint64_t, std::vector<int>> cells_;
};

}  // namespace

SampleResult SamplePoints(const float* pts, const float* padding, int n, int dims, int num_seeded,
                          const SampleOptions& o) {
  SampleResult r;
  const int m = o.num_centers, k = o.num_neighbors;
  r.center.assign(m, 0);
  r.center_padding.assign(m, 1.f);
  r.indices.assign(static_cast<size_t>(m) * k, 0);
  r.indices_padding.assign(static_cast<size_t>(m) * k, 1.f);
  if (n <= 0 || m <= 0) return r;
  num_seeded = std::max(0, std::min(num_seeded, n));
  std::mt19937_64 rng(o.seed >= 0 ? static_cast<uint64_t>(o.seed) : std::random_device{}());
  auto P = [&](int i) { return pts + static_cast<size_t>(i) * dims; };

  // --- centre selection -------------------------------------------------------
  std::vector<int> cand;                  // may become a centre
  std::vector<int> members;               // may become a neighbour
  for (int i = num_seeded; i < n; ++i) {
    if (padding[i] > 0.5f) continue;
    members.push_back(i);
    const float z = P(i)[2];
    if (z >= o.center_z_min && z <= o.center_z_max) cand.push_back(i);
  }
  std::vector<int> centers;
  if (!o.farthest) {
    std::shuffle(cand.begin(), cand.end(), rng);
    centers.assign(cand.begin(), cand.begin() + std::min<size_t>(m, cand.size()));
  } else if (!cand.empty()) {
    // min squared distance to everything chosen so far; unseeded runs start at a random point
    std::vector<float> far(cand.size(), std::numeric_limits<float>::max());
    int next = static_cast<int>(rng() % cand.size());
    if (num_seeded > 0) {
      float best = -1.f;
      for (size_t c = 0; c < cand.size(); ++c) {
        for (int s = 0; s < num_seeded; ++s) far[c] = std::min(far[c], Dist2(P(cand[c]), P(s)));
        if (far[c] > best) {
          best = far[c];
          next = static_cast<int>(c);
        }
      }
    }
    const int want = std::min<int>(m, static_cast<int>(cand.size()));
    for (int s = 0; s < want; ++s) {
      const int cur = cand[next];
      centers.push_back(cur);
      far[next] = -1.f;                    // never picked twice
      float best = -1.f;
      for (size_t c = 0; c < cand.size(); ++c) {
        if (far[c] < 0.f) continue;
        far[c] = std::min(far[c], Dist2(P(cand[c]), P(cur)));
        if (far[c] > best) {
          best = far[c];
          next = static_cast<int>(c);
        }
      }
      if (best < 0.f) break;
    }
  }

  // --- neighbourhoods -----------------------------------------------------------
  const bool bounded = o.max_dist > 0.f && o.max_dist < 1e18f;
  const float max_d2 = bounded ? o.max_dist * o.max_dist : std::numeric_limits<float>::max();
  const bool hash = bounded && (o.use_hash || members.size() > 2048);
  std::unique_ptr<CellGrid> grid;
  if (hash) grid.reset(new CellGrid(pts, dims, members, o.max_dist));
  std::vector<std::pair<float, int>> near;
  for (size_t ci = 0; ci < centers.size(); ++ci) {
    const int c = centers[ci];
    r.center[ci] = c;
    r.center_padding[ci] = 0.f;
    near.clear();
    auto visit = [&](int i) {
      const float d = Dist2(P(i), P(c));
      if (d <= max_d2) near.emplace_back(d, i);
    };
    if (hash) {
      grid->ForEachNear(P(c), visit);
    } else {
      for (int i : members) visit(i);
    }
    const int take = std::min<int>(k, static_cast<int>(near.size()));
    if (o.closest) {
      std::partial_sort(near.begin(), near.begin() + take, near.end());
    } else {
      // uniform without replacement: partial Fisher–Yates (order made canonical first so the
      // hash and brute-force paths draw the same sample for a given seed)
      std::sort(near.begin(), near.end(),
                [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.second < b.second; });
      for (int j = 0; j < take; ++j) {
        const size_t pick = j + rng() % (near.size() - j);
        std::swap(near[j], near[pick]);
      }
    }
    for (int j = 0; j < take; ++j) {
      r.indices[ci * k + j] = near[j].second;
      r.indices_padding[ci * k + j] = 0.f;
    }
  }
  return r;
}

}  // namespace lbh
