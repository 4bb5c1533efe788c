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

}  // namespace lbh
