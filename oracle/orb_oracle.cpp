// orb_oracle.cpp -- CPU ORACLE (test infrastructure only; see orb_oracle.h for scope + pin).
// Build: g++ -O3 -march=native -ffp-contract=off -std=c++17 -shared -fPIC (oracle/Makefile).
// Every function cites the reference file:line (relative to /root/reference) it follows.
#include "orb_oracle.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

namespace {

using KP = orc_keypoint;

// ---- OpenCV scalar helpers ---------------------------------------
// cvRound == SSE cvtss2si / cvtsd2si == round-half-to-even in the default rounding mode.
inline int cvRoundF(float v) { return (int)lrintf(v); }
inline int cvRoundD(double v) { return (int)lrint(v); }
inline int cvFloorF(float v) { int i = (int)v; return i - (i > v); }
inline int cvFloorD(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeilD(double v) { int i = (int)v; return i + (i < v); }
inline int reflect101(int p, int n) {
  // cv::borderInterpolate(BORDER_REFLECT_101): -i -> i, n-1+i -> n-1-i (repeated if needed)
  if (n == 1) return 0;
  while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * (n - 1) - p; }
  return p;
}

const int PATCH_SIZE = 31;       // orb_extractor.cpp:13
const int HALF_PATCH_SIZE = 15;  // :14
const int EDGE_THRESHOLD = 19;   // :15

const signed char kPattern[1024] = {
#include "orb_pattern_31.inc"
};

struct Plane {
  int w = 0, h = 0, stride = 0;
  std::vector<uint8_t> data;
  void alloc(int w_, int h_) { w = w_; h = h_; stride = w_; data.assign((size_t)w_ * h_, 0); }
  uint8_t* row(int y) { return data.data() + (size_t)y * stride; }
  const uint8_t* row(int y) const { return data.data() + (size_t)y * stride; }
};

// ---- cv::resize(INTER_LINEAR), u8 C1 (Appendix A.1; call site orb_extractor.cpp:1064) ----
void resize_linear(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                   int dstride) {
  const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
  // OpenCV: INTER_LINEAR with an exact 2x2 decimation is executed as INTER_AREA (fast path).
  {
    const int isx = (int)lrint(scale_x), isy = (int)lrint(scale_y);
    const bool fast = std::abs(scale_x - isx) < DBL_EPSILON && std::abs(scale_y - isy) < DBL_EPSILON;
    if (fast && isx == 2 && isy == 2) {
      for (int y = 0; y < dh; ++y) {
        const uint8_t* s0 = src + (size_t)(2 * y) * sstride;
        const uint8_t* s1 = s0 + sstride;
        for (int x = 0; x < dw; ++x)
          dst[(size_t)y * dstride + x] = (uint8_t)((s0[2 * x] + s0[2 * x + 1] + s1[2 * x] + s1[2 * x + 1] + 2) >> 2);
      }
      return;
    }
  }
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> xa0(dw), xa1(dw), ya0(dh), ya1(dh);
  auto coeffs = [](int d, double scale, int sn, int& ofs, short& c0, short& c1) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = cvFloorF(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= sn - 1) { s = sn - 1; f = 0.f; }
    ofs = s;
    c0 = (short)cvRoundF((1.f - f) * 2048.f);
    c1 = (short)cvRoundF(f * 2048.f);
  };
  for (int x = 0; x < dw; ++x) coeffs(x, scale_x, sw, xofs[x], xa0[x], xa1[x]);
  for (int y = 0; y < dh; ++y) coeffs(y, scale_y, sh, yofs[y], ya0[y], ya1[y]);
  std::vector<int> r0(dw), r1(dw);
  auto hrow = [&](int sy, std::vector<int>& out) {
    const uint8_t* S = src + (size_t)sy * sstride;
    for (int x = 0; x < dw; ++x) {
      const int sx = xofs[x];
      const int sx1 = std::min(sx + 1, sw - 1);
      out[x] = S[sx] * xa0[x] + S[sx1] * xa1[x];
    }
  };
  for (int y = 0; y < dh; ++y) {
    const int sy = yofs[y];
    hrow(sy, r0);
    hrow(std::min(sy + 1, sh - 1), r1);
    const int b0 = ya0[y], b1 = ya1[y];
    uint8_t* D = dst + (size_t)y * dstride;
    for (int x = 0; x < dw; ++x)
      D[x] = (uint8_t)((((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2);
  }
}

// ---- cv::copyMakeBorder(BORDER_REFLECT_101 [+ISOLATED]) (orb_extractor.cpp:1066,1071) ----
void border_reflect101(uint8_t* buf, int w, int h, int stride, int b) {
  for (int y = 0; y < h + 2 * b; ++y) {
    const int sy = reflect101(y - b, h) + b;
    uint8_t* D = buf + (size_t)y * stride;
    const uint8_t* S = buf + (size_t)sy * stride;
    for (int x = 0; x < w + 2 * b; ++x) {
      if (y >= b && y < h + b && x >= b && x < w + b) continue;
      D[x] = S[reflect101(x - b, w) + b];
    }
  }
}

// ---- cv::GaussianBlur(7x7, sigma 2, REFLECT_101), u8 fixed point (Appendix A.3) ----------
void gaussian7x7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
  static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
  std::vector<uint16_t> H((size_t)w * h);
  std::vector<uint8_t> prow((size_t)w + 6);
  for (int y = 0; y < h; ++y) {  // horizontal pass on a REFLECT_101-extended row
    const uint8_t* S = src + (size_t)y * sstride;
    for (int i = 0; i < 3; ++i) { prow[i] = S[reflect101(i - 3, w)]; prow[w + 3 + i] = S[reflect101(w + i, w)]; }
    std::memcpy(prow.data() + 3, S, (size_t)w);
    uint16_t* Hr = &H[(size_t)y * w];
    const uint8_t* P = prow.data();
    for (int x = 0; x < w; ++x)
      Hr[x] = (uint16_t)(k[0] * P[x] + k[1] * P[x + 1] + k[2] * P[x + 2] + k[3] * P[x + 3] + k[4] * P[x + 4] +
                         k[5] * P[x + 5] + k[6] * P[x + 6]);  // <= 65280
  }
  for (int y = 0; y < h; ++y) {  // vertical pass, single rounding
    const uint16_t* r[7];
    for (int j = 0; j < 7; ++j) r[j] = &H[(size_t)reflect101(y + j - 3, h) * w];
    uint8_t* D = dst + (size_t)y * dstride;
    for (int x = 0; x < w; ++x) {
      const uint32_t acc = (uint32_t)k[0] * r[0][x] + (uint32_t)k[1] * r[1][x] + (uint32_t)k[2] * r[2][x] +
                           (uint32_t)k[3] * r[3][x] + (uint32_t)k[4] * r[4][x] + (uint32_t)k[5] * r[5][x] +
                           (uint32_t)k[6] * r[6][x];
      D[x] = (uint8_t)((acc + 32768u) >> 16);
    }
  }
}

// ---- cv::FAST(img, kps, threshold, nonmaxSuppression) TYPE_9_16 (Appendix A.2) -----------
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// corner score = max threshold for which the pixel is still a FAST-9 corner (0 if none at 1)
inline int fast_score(const uint8_t* p, int stride) {
  const int v = p[0];
  int d[25];
  for (int k = 0; k < 16; ++k) d[k] = v - p[kRingDy[k] * stride + kRingDx[k]];
  for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
  int best = INT_MIN;  // max over arcs of min(d) (dark) and of min(-d) (bright)
  for (int k = 0; k < 16; ++k) {
    int mn = d[k], mx = d[k];
    for (int i = 1; i < 9; ++i) { mn = std::min(mn, d[k + i]); mx = std::max(mx, d[k + i]); }
    best = std::max(best, std::max(mn, -mx));
  }
  return best - 1;
}

void fast9(const uint8_t* img, int w, int h, int stride, int threshold, bool nms, std::vector<KP>& out) {
  out.clear();
  threshold = std::min(std::max(threshold, 0), 255);
  if (w < 7 || h < 7) return;
  std::vector<int> score((size_t)w * h, 0);
  // OpenCV's scalar structure: 512-entry class table (1 = darker than v-t, 2 = brighter than
  // v+t), opposite-pair early rejects, then the exact score only for pixels that survive.
  uint8_t tab[512];
  for (int i = -255; i <= 255; ++i) tab[i + 255] = (uint8_t)(i < -threshold ? 1 : i > threshold ? 2 : 0);
  int ofs[16];
  for (int k = 0; k < 16; ++k) ofs[k] = kRingDy[k] * stride + kRingDx[k];
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x) {
      const uint8_t* p = img + (size_t)y * stride + x;
      const uint8_t* t = tab + 255 - p[0];
      int dd = t[p[ofs[0]]] | t[p[ofs[8]]];
      if (dd == 0) continue;
      dd &= t[p[ofs[2]]] | t[p[ofs[10]]];
      dd &= t[p[ofs[4]]] | t[p[ofs[12]]];
      dd &= t[p[ofs[6]]] | t[p[ofs[14]]];
      if (dd == 0) continue;
      dd &= t[p[ofs[1]]] | t[p[ofs[9]]];
      dd &= t[p[ofs[3]]] | t[p[ofs[11]]];
      dd &= t[p[ofs[5]]] | t[p[ofs[13]]];
      dd &= t[p[ofs[7]]] | t[p[ofs[15]]];
      if (dd == 0) continue;
      const int s = fast_score(p, stride);
      // corner at `threshold` <=> score >= threshold (and cv stores score only for corners)
      if (s >= threshold && s >= 0) {
        // a pixel with all ring pixels equal to v has s = -1; threshold 0 corners need s>=0
        score[(size_t)y * w + x] = nms ? s : 1;
        if (!nms) out.push_back(KP{(float)x, (float)y, 7.f, -1.f, 0.f, 0, -1});
      }
    }
  if (!nms) return;
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x) {
      const int s = score[(size_t)y * w + x];
      if (s == 0 && threshold > 0) continue;
      if (s < threshold) continue;
      const int* c = &score[(size_t)y * w + x];
      if (s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] && s > c[w - 1] &&
          s > c[w] && s > c[w + 1])
        out.push_back(KP{(float)x, (float)y, 7.f, -1.f, (float)s, 0, -1});
    }
}

// ---- cv::fastAtan2 (Appendix A.4; call site orb_extractor.cpp:44) -------------------------
float fast_atan2(float y, float x) {
  const float scale = (float)(180 / 3.141592653589793238462643383279502884);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  const float ax = std::fabs(x), ay = std::fabs(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// ---- OrbMatcher::DescriptorDistance (orb_matcher.cpp:1630-1646) ---------------------------
inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; ++i) {
    uint32_t x, y;
    std::memcpy(&x, a + 4 * i, 4);
    std::memcpy(&y, b + 4 * i, 4);
    uint32_t v = x ^ y;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

// ---- IC_Angle (orb_extractor.cpp:18-45) ---------------------------------------------------
float ic_angle(const uint8_t* img, int stride, float px, float py, const std::vector<int>& u_max) {
  int m_01 = 0, m_10 = 0;
  const uint8_t* center = img + (ptrdiff_t)cvRoundF(py) * stride + cvRoundF(px);
  for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
  for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
    int v_sum = 0;
    const int d = u_max[v];
    for (int u = -d; u <= d; ++u) {
      const int val_plus = center[u + v * stride], val_minus = center[u - v * stride];
      v_sum += (val_plus - val_minus);
      m_10 += u * (val_plus + val_minus);
    }
    m_01 += v * v_sum;
  }
  return fast_atan2((float)m_01, (float)m_10);
}

// ---- computeOrbDescriptor (orb_extractor.cpp:48-88) ---------------------------------------
const float factorPI = (float)(3.141592653589793238462643383279502884 / 180.f);  // :48
// `img` is the blurred CLONE of the level: a continuous w x h buffer (step == w), so a sample
// whose column falls outside [0,w) reads the neighbouring row exactly as the reference's
// center[iy*step + ix] does.  The pattern reaches 18 px but keypoints may sit 16 px from the edge;
// a sample before/after the whole buffer is an out-of-allocation read in the reference (UB):
// the oracle defines it as 0 (header item 6).
void orb_descriptor(const KP& kpt, const uint8_t* img, int step, int rows, uint8_t* desc) {
  const float angle = (float)kpt.angle * factorPI;
  const float a = cosf(angle), b = sinf(angle);  // cos/sin on a float => cosf/sinf (:54)
  const ptrdiff_t center = (ptrdiff_t)cvRoundF(kpt.y) * step + cvRoundF(kpt.x);
  const ptrdiff_t total = (ptrdiff_t)step * rows;
  const signed char* pat = kPattern;
  auto get = [&](int idx) -> int {
    const float x = (float)pat[2 * idx], y = (float)pat[2 * idx + 1];
    const ptrdiff_t lin = center + (ptrdiff_t)cvRoundF(x * b + y * a) * step + cvRoundF(x * a - y * b);
    return (lin < 0 || lin >= total) ? 0 : img[lin];
  };
  for (int i = 0; i < 32; ++i, pat += 32) {
    int val = 0;
    for (int k = 0; k < 8; ++k) {
      const int t0 = get(2 * k), t1 = get(2 * k + 1);
      val |= (t0 < t1) << k;
    }
    desc[i] = (uint8_t)val;
  }
}

// ---- ExtractorNode / DistributeOctTree (orb_extractor.cpp:422-704) ------------------------
struct Node {
  std::vector<KP> keys;
  int ULx = 0, ULy = 0, URx = 0, URy = 0, BLx = 0, BLy = 0, BRx = 0, BRy = 0;
  std::list<Node>::iterator lit;
  bool noMore = false;
  long seq = 0;  // creation sequence: the oracle's tie-break for :625 (see header, item 1)
};

void divide_node(const Node& p, Node& n1, Node& n2, Node& n3, Node& n4) {  // :422-478
  const int halfX = (int)std::ceil(static_cast<float>(p.URx - p.ULx) / 2);
  const int halfY = (int)std::ceil(static_cast<float>(p.BRy - p.ULy) / 2);
  n1.ULx = p.ULx; n1.ULy = p.ULy;
  n1.URx = p.ULx + halfX; n1.URy = p.ULy;
  n1.BLx = p.ULx; n1.BLy = p.ULy + halfY;
  n1.BRx = p.ULx + halfX; n1.BRy = p.ULy + halfY;
  n2.ULx = n1.URx; n2.ULy = n1.URy;
  n2.URx = p.URx; n2.URy = p.URy;
  n2.BLx = n1.BRx; n2.BLy = n1.BRy;
  n2.BRx = p.URx; n2.BRy = p.ULy + halfY;
  n3.ULx = n1.BLx; n3.ULy = n1.BLy;
  n3.URx = n1.BRx; n3.URy = n1.BRy;
  n3.BLx = p.BLx; n3.BLy = p.BLy;
  n3.BRx = n1.BRx; n3.BRy = p.BLy;
  n4.ULx = n3.URx; n4.ULy = n3.URy;
  n4.URx = n2.BRx; n4.URy = n2.BRy;
  n4.BLx = n3.BRx; n4.BLy = n3.BRy;
  n4.BRx = p.BRx; n4.BRy = p.BRy;
  for (const KP& kp : p.keys) {
    if (kp.x < n1.URx) {
      if (kp.y < n1.BRy) n1.keys.push_back(kp); else n3.keys.push_back(kp);
    } else if (kp.y < n1.BRy) n2.keys.push_back(kp);
    else n4.keys.push_back(kp);
  }
  if (n1.keys.size() == 1) n1.noMore = true;
  if (n2.keys.size() == 1) n2.noMore = true;
  if (n3.keys.size() == 1) n3.noMore = true;
  if (n4.keys.size() == 1) n4.noMore = true;
}

std::vector<KP> distribute_octree(const std::vector<KP>& in, int minX, int maxX, int minY, int maxY, int N) {
  std::vector<KP> result;
  if (in.empty()) return result;
  // :484 (nIni==0 for very tall images is a division by zero in the reference; clamped here)
  int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
  if (nIni < 1) nIni = 1;
  const float hX = static_cast<float>(maxX - minX) / nIni;
  std::list<Node> lNodes;
  std::vector<Node*> ini(nIni);
  long seq = 0;
  for (int i = 0; i < nIni; ++i) {
    Node ni;
    ni.ULx = (int)(hX * static_cast<float>(i)); ni.ULy = 0;
    ni.URx = (int)(hX * static_cast<float>(i + 1)); ni.URy = 0;
    ni.BLx = ni.ULx; ni.BLy = maxY - minY;
    ni.BRx = ni.URx; ni.BRy = maxY - minY;
    ni.seq = seq++;
    lNodes.push_back(ni);
    ini[i] = &lNodes.back();
  }
  for (const KP& kp : in) {
    int r = (int)(kp.x / hX);
    if (r >= nIni) r = nIni - 1;  // cannot happen for FAST output (x <= W'-4); guards UB
    ini[r]->keys.push_back(kp);
  }
  for (auto lit = lNodes.begin(); lit != lNodes.end();) {
    if (lit->keys.size() == 1) { lit->noMore = true; ++lit; }
    else if (lit->keys.empty()) lit = lNodes.erase(lit);
    else ++lit;
  }
  bool bFinish = false;
  std::vector<std::pair<int, Node*>> vSizeAndPtr;
  auto push_child = [&](Node& c, int* nToExpand) {
    if (c.keys.empty()) return;
    c.seq = seq++;
    lNodes.push_front(c);
    if (c.keys.size() > 1) {
      if (nToExpand) ++*nToExpand;
      vSizeAndPtr.push_back(std::make_pair((int)c.keys.size(), &lNodes.front()));
      lNodes.front().lit = lNodes.begin();
    }
  };
  while (!bFinish) {
    const int prevSize = (int)lNodes.size();
    auto lit = lNodes.begin();
    int nToExpand = 0;
    vSizeAndPtr.clear();
    while (lit != lNodes.end()) {
      if (lit->noMore) { ++lit; continue; }
      Node n1, n2, n3, n4;
      divide_node(*lit, n1, n2, n3, n4);
      push_child(n1, &nToExpand); push_child(n2, &nToExpand);
      push_child(n3, &nToExpand); push_child(n4, &nToExpand);
      lit = lNodes.erase(lit);
    }
    if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) {
      bFinish = true;
    } else if (((int)lNodes.size() + nToExpand * 3) > N) {
      while (!bFinish) {
        const int prevSize2 = (int)lNodes.size();
        std::vector<std::pair<int, Node*>> prev = vSizeAndPtr;
        vSizeAndPtr.clear();
        // :625 std::sort on (size, pointer): ties defined as creation sequence (header item 1)
        std::stable_sort(prev.begin(), prev.end(),
                         [](const std::pair<int, Node*>& a, const std::pair<int, Node*>& b) {
                           if (a.first != b.first) return a.first < b.first;
                           return a.second->seq < b.second->seq;
                         });
        for (int j = (int)prev.size() - 1; j >= 0; --j) {
          Node n1, n2, n3, n4;
          divide_node(*prev[j].second, n1, n2, n3, n4);
          push_child(n1, nullptr); push_child(n2, nullptr);
          push_child(n3, nullptr); push_child(n4, nullptr);
          lNodes.erase(prev[j].second->lit);
          if ((int)lNodes.size() >= N) break;
        }
        if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize2) bFinish = true;
      }
    }
  }
  result.reserve(lNodes.size());
  for (auto& nd : lNodes) {  // :682-701
    const KP* best = &nd.keys[0];
    float maxResponse = best->response;
    for (size_t k = 1; k < nd.keys.size(); ++k)
      if (nd.keys[k].response > maxResponse) { best = &nd.keys[k]; maxResponse = nd.keys[k].response; }
    result.push_back(*best);
  }
  return result;
}

}  // namespace

// ===========================================================================================
struct orc_extractor {
  int nfeatures, nlevels, iniThFAST, minThFAST;
  double scaleFactor;  // orb_extractor.h:79 -- a double holding the float argument
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> featuresPerLevel, umax;
  std::vector<Plane> padded;   // (w+38)x(h+38)
  std::vector<int> lw, lh;     // level sizes
  std::vector<Plane> blurred;  // stage output
  std::vector<std::vector<KP>> cand, levelKps;

  const uint8_t* roi(int l) const { return padded[l].row(EDGE_THRESHOLD) + EDGE_THRESHOLD; }
  uint8_t* roi(int l) { return padded[l].row(EDGE_THRESHOLD) + EDGE_THRESHOLD; }
};

extern "C" {

void orc_resize_linear(const uint8_t* s, int sw, int sh, int ss, uint8_t* d, int dw, int dh, int ds) {
  resize_linear(s, sw, sh, ss, d, dw, dh, ds);
}
void orc_border_reflect101(uint8_t* buf, int w, int h, int stride, int b) { border_reflect101(buf, w, h, stride, b); }
void orc_gaussian7x7(const uint8_t* s, int w, int h, int ss, uint8_t* d, int ds) { gaussian7x7(s, w, h, ss, d, ds); }
int orc_fast9(const uint8_t* img, int w, int h, int stride, int th, int nms, orc_keypoint* out, int cap) {
  std::vector<KP> v;
  fast9(img, w, h, stride, th, nms != 0, v);
  for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
// cv::cvtColor(RGB2GRAY / BGR2GRAY / RGBA2GRAY / BGRA2GRAY) on 8-bit images (call sites
// src/core/tracker.cpp:110-127): OpenCV 4.x fixed point, 15 fractional bits,
// gray = (R*9798 + G*19235 + B*3735 + 16384) >> 15  (pinned against cv2 4.13; OpenCV 3.x used 14 bits)
void orc_cvt_gray(const uint8_t* src, size_t n_px, int channels, int rgb_order, uint8_t* dst) {
  for (size_t i = 0; i < n_px; ++i) {
    const uint8_t* p = src + i * (size_t)channels;
    const int r = rgb_order ? p[0] : p[2], g = p[1], b = rgb_order ? p[2] : p[0];
    dst[i] = (uint8_t)((r * 9798 + g * 19235 + b * 3735 + 16384) >> 15);
  }
}
int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

// ORBextractor::ORBextractor (orb_extractor.cpp:351-411)
orc_extractor* orc_extractor_create(int nfeatures, float scaleFactorF, int nlevels, int iniTh, int minTh) {
  orc_extractor* e = new orc_extractor();
  e->nfeatures = nfeatures; e->nlevels = nlevels; e->iniThFAST = iniTh; e->minThFAST = minTh;
  e->scaleFactor = scaleFactorF;
  e->scale.resize(nlevels); e->sigma2.resize(nlevels);
  e->scale[0] = 1.0f; e->sigma2[0] = 1.0f;
  for (int i = 1; i < nlevels; i++) {
    e->scale[i] = (float)(e->scale[i - 1] * e->scaleFactor);  // float*double -> float (:362)
    e->sigma2[i] = e->scale[i] * e->scale[i];
  }
  e->invScale.resize(nlevels); e->invSigma2.resize(nlevels);
  for (int i = 0; i < nlevels; i++) { e->invScale[i] = 1.0f / e->scale[i]; e->invSigma2[i] = 1.0f / e->sigma2[i]; }
  e->featuresPerLevel.resize(nlevels);
  const float factor = (float)(1.0f / e->scaleFactor);  // :375
  float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
  int sum = 0;
  for (int level = 0; level < nlevels - 1; level++) {
    e->featuresPerLevel[level] = cvRoundF(nDesired);
    sum += e->featuresPerLevel[level];
    nDesired *= factor;
  }
  e->featuresPerLevel[nlevels - 1] = std::max(nfeatures - sum, 0);
  // umax (:393-410)
  e->umax.assign(HALF_PATCH_SIZE + 1, 0);
  int v, v0;
  const int vmax = cvFloorD(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
  const int vmin = cvCeilD(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
  const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
  for (v = 0; v <= vmax; ++v) e->umax[v] = cvRoundD(std::sqrt(hp2 - v * v));
  for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
    while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
    e->umax[v] = v0;
    ++v0;
  }
  e->padded.resize(nlevels); e->blurred.resize(nlevels);
  e->lw.assign(nlevels, 0); e->lh.assign(nlevels, 0);
  e->cand.resize(nlevels); e->levelKps.resize(nlevels);
  return e;
}
void orc_extractor_destroy(orc_extractor* e) { delete e; }
int orc_levels(const orc_extractor* e) { return e->nlevels; }
void orc_scale_factors(const orc_extractor* e, float* s, float* is, float* s2, float* is2) {
  for (int i = 0; i < e->nlevels; ++i) {
    if (s) s[i] = e->scale[i];
    if (is) is[i] = e->invScale[i];
    if (s2) s2[i] = e->sigma2[i];
    if (is2) is2[i] = e->invSigma2[i];
  }
}
void orc_features_per_level(const orc_extractor* e, int* out) { for (int i = 0; i < e->nlevels; ++i) out[i] = e->featuresPerLevel[i]; }
void orc_umax(const orc_extractor* e, int* out) { for (int i = 0; i < 16; ++i) out[i] = e->umax[i]; }

const uint8_t* orc_pyramid_level(const orc_extractor* e, int l, int* w, int* h, int* stride) {
  *w = e->lw[l]; *h = e->lh[l]; *stride = e->padded[l].stride;
  return e->roi(l);
}
const uint8_t* orc_pyramid_padded(const orc_extractor* e, int l, int* w, int* h, int* stride) {
  *w = e->padded[l].w; *h = e->padded[l].h; *stride = e->padded[l].stride;
  return e->padded[l].data.data();
}
int orc_stage_candidates(const orc_extractor* e, int l, orc_keypoint* out, int cap) {
  const auto& v = e->cand[l];
  for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}
int orc_stage_level_keypoints(const orc_extractor* e, int l, orc_keypoint* out, int cap) {
  const auto& v = e->levelKps[l];
  for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}
const uint8_t* orc_stage_blurred(const orc_extractor* e, int l, int* w, int* h, int* stride) {
  *w = e->blurred[l].w; *h = e->blurred[l].h; *stride = e->blurred[l].stride;
  return e->blurred[l].data.data();
}

int orc_distribute_octree(const orc_keypoint* cand, int n, int minX, int maxX, int minY, int maxY, int N,
                          orc_keypoint* out, int cap) {
  std::vector<KP> in(cand, cand + n);
  std::vector<KP> r = distribute_octree(in, minX, maxX, minY, maxY, N);
  for (int i = 0; i < (int)r.size() && i < cap; ++i) out[i] = r[i];
  return (int)r.size();
}

// ORBextractor::Compute (orb_extractor.cpp:985-1049)
int orc_extract(orc_extractor* e, const uint8_t* img, int w, int h, int stride, orc_keypoint* kps, uint8_t* desc,
                int cap) {
  if (!img || w <= 0 || h <= 0) return 0;  // :990-991
  const int nlevels = e->nlevels;
  // ---- ComputePyramid (:1051-1076)
  for (int level = 0; level < nlevels; ++level) {
    const float scale = e->invScale[level];
    const int sw = cvRoundF((float)w * scale), sh = cvRoundF((float)h * scale);
    e->lw[level] = sw; e->lh[level] = sh;
    e->padded[level].alloc(sw + EDGE_THRESHOLD * 2, sh + EDGE_THRESHOLD * 2);
    uint8_t* roi = e->roi(level);
    const int ps = e->padded[level].stride;
    if (level != 0) {
      resize_linear(e->roi(level - 1), e->lw[level - 1], e->lh[level - 1], e->padded[level - 1].stride, roi, sw, sh, ps);
    } else {
      for (int y = 0; y < h; ++y) std::memcpy(roi + (size_t)y * ps, img + (size_t)y * stride, (size_t)w);
    }
    border_reflect101(e->padded[level].data.data(), sw, sh, ps, EDGE_THRESHOLD);
  }
  // ---- ComputeKeyPointsOctTree (:706-794)
  const float W = 30;
  std::vector<std::vector<KP>> all(nlevels);
  for (int level = 0; level < nlevels; ++level) {
    const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
    const int maxBorderX = e->lw[level] - EDGE_THRESHOLD + 3;
    const int maxBorderY = e->lh[level] - EDGE_THRESHOLD + 3;
    std::vector<KP>& toDistribute = e->cand[level];
    toDistribute.clear();
    e->levelKps[level].clear();
    const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols < 1 || nRows < 1) continue;  // reference divides by zero here; level too small
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    const uint8_t* roi = e->roi(level);
    const int ps = e->padded[level].stride;
    std::vector<KP> cell;
    for (int i = 0; i < nRows; i++) {
      const float iniY = (float)(minBorderY + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBorderY - 3) continue;
      if (maxY > maxBorderY) maxY = (float)maxBorderY;
      for (int j = 0; j < nCols; j++) {
        const float iniX = (float)(minBorderX + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBorderX - 6) continue;
        if (maxX > maxBorderX) maxX = (float)maxBorderX;
        const uint8_t* sub = roi + (ptrdiff_t)(int)iniY * ps + (int)iniX;
        const int cw = (int)maxX - (int)iniX, ch = (int)maxY - (int)iniY;
        fast9(sub, cw, ch, ps, e->iniThFAST, true, cell);
        if (cell.empty()) fast9(sub, cw, ch, ps, e->minThFAST, true, cell);
        for (KP& k : cell) {
          k.x += j * wCell;
          k.y += i * hCell;
          toDistribute.push_back(k);
        }
      }
    }
    std::vector<KP>& keypoints = all[level];
    keypoints = distribute_octree(toDistribute, minBorderX, maxBorderX, minBorderY, maxBorderY,
                                  e->featuresPerLevel[level]);
    const int scaledPatchSize = (int)(PATCH_SIZE * e->scale[level]);
    for (KP& k : keypoints) {
      k.x += minBorderX;
      k.y += minBorderY;
      k.octave = level;
      k.size = (float)scaledPatchSize;
    }
  }
  for (int level = 0; level < nlevels; ++level)  // computeOrientation (:413-420)
    for (KP& k : all[level]) k.angle = ic_angle(e->roi(level), e->padded[level].stride, k.x, k.y, e->umax);

  int nkeypoints = 0;
  for (int level = 0; level < nlevels; ++level) nkeypoints += (int)all[level].size();
  const bool fits = nkeypoints <= cap;
  int offset = 0;
  for (int level = 0; level < nlevels; ++level) {
    std::vector<KP>& keypoints = all[level];
    e->levelKps[level] = keypoints;
    if (keypoints.empty()) { e->blurred[level].alloc(0, 0); continue; }
    // clone of the un-padded level + GaussianBlur (:1029-1030)
    Plane& work = e->blurred[level];
    work.alloc(e->lw[level], e->lh[level]);
    gaussian7x7(e->roi(level), e->lw[level], e->lh[level], e->padded[level].stride, work.data.data(), work.stride);
    for (size_t i = 0; i < keypoints.size(); ++i) {
      if (fits) orb_descriptor(keypoints[i], work.data.data(), work.stride, work.h, desc + (size_t)(offset + i) * 32);
    }
    if (level != 0) {
      const float scale = e->scale[level];
      for (KP& k : keypoints) { k.x *= scale; k.y *= scale; }
    }
    if (fits) for (size_t i = 0; i < keypoints.size(); ++i) kps[offset + i] = keypoints[i];
    offset += (int)keypoints.size();
  }
  return fits ? nkeypoints : -nkeypoints;
}

// Frame::ComputeStereoMatches (frame.cpp:406-577)
int orc_stereo_match(const orc_extractor* L, const orc_extractor* R, int nl, const orc_keypoint* kl,
                     const uint8_t* dl, int nr, const orc_keypoint* kr, const uint8_t* dr, float bf, float baseline,
                     float* uRight, float* depth) {
  for (int i = 0; i < nl; ++i) { uRight[i] = -1.0f; depth[i] = -1.0f; }
  const int TH_HIGH = 100, TH_LOW = 50;  // orb_matcher.cpp:5-6
  const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
  const int nRows = L->lh[0];
  std::vector<std::vector<size_t>> vRowIndices(nRows);
  for (int iR = 0; iR < nr; iR++) {
    const float kpY = kr[iR].y;
    const float r = 2.0f * L->scale[kr[iR].octave];
    const int maxr = (int)std::ceil(kpY + r);
    const int minr = (int)std::floor(kpY - r);
    for (int yi = minr; yi <= maxr; ++yi)
      if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);  // guard: reference indexes unchecked
  }
  const float minZ = baseline;  // header item 2
  const float minD = 0;
  const float maxD = bf / minZ;
  std::vector<std::pair<int, int>> vDistIdx;
  for (int iL = 0; iL < nl; ++iL) {
    const KP& kpL = kl[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const size_t row = (size_t)vL;
    if (row >= (size_t)nRows) continue;  // guard
    const std::vector<size_t>& cands = vRowIndices[row];
    if (cands.empty()) continue;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = TH_HIGH;
    size_t bestIdxR = 0;
    const uint8_t* dL = dl + (size_t)iL * 32;
    for (size_t iC = 0; iC < cands.size(); ++iC) {
      const size_t iR = cands[iC];
      const KP& kpR = kr[iR];
      if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
      const float uR = kpR.x;
      if (uR >= minU && uR <= maxU) {
        const int dist = descriptor_distance(dL, dr + iR * 32);
        if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
      }
    }
    if (bestDist < thOrbDist) {
      const float uR0 = kr[bestIdxR].x;
      const float scaleFactor = L->invScale[kpL.octave];
      const float scaleduL = std::round(kpL.x * scaleFactor);
      const float scaledvL = std::round(kpL.y * scaleFactor);
      const float scaleduR0 = std::round(uR0 * scaleFactor);
      const int w = 5;
      const uint8_t* pl = L->roi(kpL.octave); const int sl = L->padded[kpL.octave].stride;
      const uint8_t* pr = R->roi(kpL.octave); const int sr = R->padded[kpL.octave].stride;
      const int rcols = R->lw[kpL.octave];
      const int yl0 = (int)(scaledvL - w), xl0 = (int)(scaleduL - w);
      float IL[11][11];
      const float cL = (float)pl[(ptrdiff_t)(yl0 + w) * sl + xl0 + w];
      for (int y = 0; y < 11; ++y)
        for (int x = 0; x < 11; ++x) IL[y][x] = (float)pl[(ptrdiff_t)(yl0 + y) * sl + xl0 + x] - cL;
      int bestDistS = INT_MAX;
      int bestincR = 0;
      const int Lw = 5;
      float vDists[11];
      const float iniu = scaleduR0 + Lw - w;
      const float endu = scaleduR0 + Lw + w + 1;
      if (iniu < 0 || endu >= rcols) continue;
      for (int incR = -Lw; incR <= Lw; ++incR) {
        const int xr0 = (int)(scaleduR0 + incR - w);
        const float cR = (float)pr[(ptrdiff_t)(yl0 + w) * sr + xr0 + w];
        double acc = 0;  // cv::norm(NORM_L1) accumulates in double; values are exact integers
        for (int y = 0; y < 11; ++y)
          for (int x = 0; x < 11; ++x)
            acc += std::fabs(IL[y][x] - ((float)pr[(ptrdiff_t)(yl0 + y) * sr + xr0 + x] - cR));
        const float dist = (float)acc;
        if (dist < bestDistS) { bestDistS = (int)dist; bestincR = incR; }
        vDists[Lw + incR] = dist;
      }
      if (bestincR == -Lw || bestincR == Lw) continue;
      const float dist1 = vDists[Lw + bestincR - 1], dist2 = vDists[Lw + bestincR], dist3 = vDists[Lw + bestincR + 1];
      const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
      if (deltaR < -1 || deltaR > 1) continue;
      float bestuR = L->scale[kpL.octave] * (scaleduR0 + (float)bestincR + deltaR);
      float disparity = (uL - bestuR);
      if (disparity >= minD && disparity < maxD) {
        if (disparity <= 0) { disparity = 0.01f; bestuR = uL - 0.01f; }
        depth[iL] = bf / disparity;
        uRight[iL] = bestuR;
        vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
      }
    }
  }
  if (vDistIdx.empty()) return 0;  // header item 3
  std::sort(vDistIdx.begin(), vDistIdx.end());
  const float median = (float)vDistIdx[vDistIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  int kept = (int)vDistIdx.size();
  for (int i = (int)vDistIdx.size() - 1; i >= 0; --i) {
    if (vDistIdx[i].first < thDist) break;
    uRight[vDistIdx[i].second] = -1.0f;
    depth[vDistIdx[i].second] = -1.0f;
    --kept;
  }
  return kept;
}

}  // extern "C"

// ===========================================================================================
// Frame grid + matchers
struct orc_frame {
  int n = 0, nlevels = 0;
  std::vector<KP> kps;
  std::vector<uint8_t> desc;
  std::vector<float> uR;
  std::vector<float> scale;
  float minX, maxX, minY, maxY, gw, gh;
  static const int COLS = 64, ROWS = 48;  // frame.h:104-105
  std::vector<size_t> grid[COLS][ROWS];
};

namespace {
// Frame::GetFeaturesInArea (frame.cpp:348-403)
std::vector<size_t> features_in_area(const orc_frame* f, float x, float y, float r, int minLevel, int maxLevel) {
  std::vector<size_t> v;
  const int nMinCellX = std::max(0, (int)std::floor((x - f->minX - r) / f->gw));
  const int nMaxCellX = std::min(orc_frame::COLS - 1, (int)std::ceil((x - f->minX + r) / f->gw));
  if (nMaxCellX < 0 || nMinCellX >= orc_frame::COLS) return v;
  const int nMinCellY = std::max(0, (int)std::floor((y - f->minY - r) / f->gh));
  const int nMaxCellY = std::min(orc_frame::ROWS - 1, (int)std::ceil((y - f->minY + r) / f->gh));
  if (nMaxCellY < 0 || nMinCellY >= orc_frame::ROWS) return v;
  const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
    for (int iy = nMinCellY; iy <= nMaxCellY; ++iy) {
      const std::vector<size_t>& cell = f->grid[ix][iy];
      for (size_t j = 0; j < cell.size(); ++j) {
        const KP& kp = f->kps[cell[j]];
        if (bCheckLevels) {
          if (kp.octave < minLevel) continue;
          if (maxLevel >= 0 && kp.octave > maxLevel) continue;
        }
        const float dx = kp.x - x, dy = kp.y - y;
        if (std::fabs(dx) < r && std::fabs(dy) < r) v.push_back(cell[j]);
      }
    }
  return v;
}

// OrbMatcher::ComputeThreeMaxima (orb_matcher.cpp:1584-1625)
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}
const int HISTO_LENGTH = 30;
}  // namespace

namespace {
struct FV { int n; const uint32_t* ids; const int* start; const uint32_t* idx; };
// the two-iterator walk over the common vocabulary nodes (orb_matcher.cpp:152-236, :527-609, :670-768);
// lower_bound on a sorted map == advance until >=
template <class Fn>
void for_common_nodes(const FV& A, const FV& B, Fn fn) {
  int a = 0, b = 0;
  while (a < A.n && b < B.n) {
    if (A.ids[a] == B.ids[b]) { fn(a, b); ++a; ++b; }
    else if (A.ids[a] < B.ids[b]) ++a;
    else ++b;
  }
}
}  // namespace

extern "C" {

orc_frame* orc_frame_create(int n, const orc_keypoint* kps, const uint8_t* desc, const float* uR, float minX,
                            float maxX, float minY, float maxY, int nlevels, const float* scale) {
  orc_frame* f = new orc_frame();
  f->n = n; f->nlevels = nlevels;
  f->kps.assign(kps, kps + n);
  f->desc.assign(desc, desc + (size_t)n * 32);
  if (uR) f->uR.assign(uR, uR + n); else f->uR.assign(n, -1.0f);
  f->scale.assign(scale, scale + nlevels);
  f->minX = minX; f->maxX = maxX; f->minY = minY; f->maxY = maxY;
  f->gw = static_cast<float>(maxX - minX) / orc_frame::COLS;  // frame.cpp:223-224
  f->gh = static_cast<float>(maxY - minY) / orc_frame::ROWS;
  for (int i = 0; i < n; ++i) {  // AssignFeaturesToGrid + PosInGrid (frame.cpp:234-248, 339-346)
    const int px = (int)std::round((kps[i].x - minX) / f->gw);
    const int py = (int)std::round((kps[i].y - minY) / f->gh);
    if (px >= 0 && px < orc_frame::COLS && py >= 0 && py < orc_frame::ROWS) f->grid[px][py].push_back(i);
  }
  return f;
}
void orc_frame_destroy(orc_frame* f) { delete f; }
int orc_features_in_area(const orc_frame* f, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap) {
  std::vector<size_t> v = features_in_area(f, x, y, r, minLevel, maxLevel);
  for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = (int)v[i];
  return (int)v.size();
}

// OrbMatcher::SearchForInitialization (orb_matcher.cpp:264-382)
int orc_search_for_initialization(const orc_frame* F1, const orc_frame* F2, float* prevMatched, int* vnMatches12,
                                  int windowSize, float nnratio, int checkOri) {
  const int TH_LOW = 50;
  int nmatches = 0;
  for (int i = 0; i < F1->n; ++i) vnMatches12[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  std::vector<int> vMatchedDistance(F2->n, INT_MAX);
  std::vector<int> vnMatches21(F2->n, -1);
  for (int i1 = 0; i1 < F1->n; i1++) {
    const KP& kp1 = F1->kps[i1];
    const int level1 = kp1.octave;
    if (level1 > 0) continue;
    std::vector<size_t> vIndices2 =
        features_in_area(F2, prevMatched[2 * i1], prevMatched[2 * i1 + 1], (float)windowSize, level1, level1);
    if (vIndices2.empty()) continue;
    const uint8_t* d1 = F1->desc.data() + (size_t)i1 * 32;
    int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
    for (size_t i2 : vIndices2) {
      const int dist = descriptor_distance(d1, F2->desc.data() + i2 * 32);
      if (vMatchedDistance[i2] <= dist) continue;
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = (int)i2; }
      else if (dist < bestDist2) { bestDist2 = dist; }
    }
    if (bestDist <= TH_LOW) {
      if (bestDist < (float)bestDist2 * nnratio) {
        if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
        vnMatches12[i1] = bestIdx2;
        vnMatches21[bestIdx2] = i1;
        vMatchedDistance[bestIdx2] = bestDist;
        nmatches++;
        if (checkOri) {
          float rot = F1->kps[i1].angle - F2->kps[bestIdx2].angle;
          if (rot < 0.0) rot += 360.0f;
          int bin = (int)std::round(rot * factor);
          if (bin == HISTO_LENGTH) bin = 0;
          rotHist[bin].push_back(i1);
        }
      }
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i])
        if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
    }
  }
  for (int i1 = 0; i1 < F1->n; i1++)
    if (vnMatches12[i1] >= 0) {
      prevMatched[2 * i1] = F2->kps[vnMatches12[i1]].x;
      prevMatched[2 * i1 + 1] = F2->kps[vnMatches12[i1]].y;
    }
  return nmatches;
}

// OrbMatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (orb_matcher.cpp:13-111)
int orc_search_by_projection_mappoints(const orc_frame* F, int nMP, const uint8_t* valid, const float* projX,
                                       const float* projY, const float* projXR, const int* predLevel,
                                       const float* viewCos, const uint8_t* mpDesc, const uint8_t* hasObs,
                                       const uint8_t* occupiedIn, int th, float nnratio, int* assigned) {
  const int TH_HIGH = 100;
  int nmatches = 0;
  std::vector<uint8_t> occupied(occupiedIn, occupiedIn + F->n);
  for (int i = 0; i < F->n; ++i) assigned[i] = -1;
  const bool bFactor = (th != 1);
  for (int iMP = 0; iMP < nMP; iMP++) {
    if (!valid[iMP]) continue;
    const int nPredictedLevel = predLevel[iMP];
    float r = (viewCos[iMP] > 0.998) ? 2.5f : 4.0f;  // RadiusByViewingCos (:105-111)
    if (bFactor) r *= th;
    const std::vector<size_t> vIndices = features_in_area(F, projX[iMP], projY[iMP], r * F->scale[nPredictedLevel],
                                                          nPredictedLevel - 1, nPredictedLevel);
    if (vIndices.empty()) continue;
    const uint8_t* d0 = mpDesc + (size_t)iMP * 32;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (size_t idx : vIndices) {
      if (occupied[idx]) continue;
      if (F->uR[idx] > 0) {
        const float er = std::fabs(projXR[iMP] - F->uR[idx]);
        if (er > r * F->scale[nPredictedLevel]) continue;
      }
      const int dist = descriptor_distance(d0, F->desc.data() + idx * 32);
      if (dist < bestDist) {
        bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F->kps[idx].octave; bestIdx = (int)idx;
      } else if (dist < bestDist2) {
        bestLevel2 = F->kps[idx].octave; bestDist2 = dist;
      }
    }
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
      assigned[bestIdx] = iMP;  // F.SetMapPoint(bestIdx, pMP)
      occupied[bestIdx] = hasObs[iMP];
      nmatches++;
    }
  }
  return nmatches;
}

// OrbMatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) (orb_matcher.cpp:1312-1453)
int orc_search_by_projection_lastframe(const orc_frame* C, int nLast, const uint8_t* valid, const float* us,
                                       const float* vs, const float* invzcs, const int* lastOctave,
                                       const float* lastAngle, const uint8_t* mpDesc, const uint8_t* hasObs, float bf,
                                       int bForward, int bBackward, const uint8_t* occupiedIn, float th, int checkOri,
                                       int* assigned) {
  const int TH_HIGH = 100;
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  std::vector<uint8_t> occupied(occupiedIn, occupiedIn + C->n);
  for (int i = 0; i < C->n; ++i) assigned[i] = -1;
  for (int i = 0; i < nLast; ++i) {
    if (!valid[i]) continue;
    const float invzc = invzcs[i];
    if (invzc < 0) continue;
    const float u = us[i], v = vs[i];
    if (u < C->minX || u > C->maxX) continue;
    if (v < C->minY || v > C->maxY) continue;
    const int nLastOctave = lastOctave[i];
    const float radius = th * C->scale[nLastOctave];
    std::vector<size_t> vIndices2;
    if (bForward) vIndices2 = features_in_area(C, u, v, radius, nLastOctave, -1);
    else if (bBackward) vIndices2 = features_in_area(C, u, v, radius, 0, nLastOctave);
    else vIndices2 = features_in_area(C, u, v, radius, nLastOctave - 1, nLastOctave + 1);
    if (vIndices2.empty()) continue;
    const uint8_t* dMP = mpDesc + (size_t)i * 32;
    int bestDist = 256, bestIdx2 = -1;
    for (size_t i2 : vIndices2) {
      if (occupied[i2]) continue;
      if (C->uR[i2] > 0) {
        const float ur = u - bf * invzc;
        const float er = std::fabs(ur - C->uR[i2]);
        if (er > radius) continue;
      }
      const int dist = descriptor_distance(dMP, C->desc.data() + i2 * 32);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = (int)i2; }
    }
    if (bestDist <= TH_HIGH) {
      assigned[bestIdx2] = i;  // CurrentFrame.SetMapPoint(bestIdx2, pMP)
      occupied[bestIdx2] = hasObs[i];
      ++nmatches;
      if (checkOri) {
        float rot = lastAngle[i] - C->kps[bestIdx2].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(bestIdx2);
      }
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++)
      if (i != ind1 && i != ind2 && i != ind3)
        for (int idx : rotHist[i]) { assigned[idx] = -1; --nmatches; }
  }
  return nmatches;
}

// OrbMatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (orb_matcher.cpp:133-262).  The DBoW2
// FeatureVectors (std::map<NodeId, vector<unsigned>>) arrive flattened (ids ascending + offsets + indices).
int orc_search_by_bow(const orc_frame* F, int nKF, const uint8_t* kfDesc, const float* kfAngle, const uint8_t* kfValid,
                      int kfNodes, const uint32_t* kfIds, const int* kfStart, const uint32_t* kfIdx, int fNodes,
                      const uint32_t* fIds, const int* fStart, const uint32_t* fIdx, float nnratio, int checkOri,
                      int* matchedKf) {
  (void)nKF;
  const int TH_LOW = 50;
  int nmatches = 0;
  for (int i = 0; i < F->n; ++i) matchedKf[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  int a = 0, b = 0;
  while (a < kfNodes && b < fNodes) {
    if (kfIds[a] == fIds[b]) {
      for (int k = kfStart[a]; k < kfStart[a + 1]; ++k) {
        const unsigned realIdxKF = kfIdx[k];
        if (!kfValid[realIdxKF]) continue;
        const uint8_t* dKF = kfDesc + (size_t)realIdxKF * 32;
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int j = fStart[b]; j < fStart[b + 1]; ++j) {
          const unsigned realIdxF = fIdx[j];
          if (matchedKf[realIdxF] >= 0) continue;
          const int dist = descriptor_distance(dKF, F->desc.data() + (size_t)realIdxF * 32);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = (int)realIdxF; }
          else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 <= TH_LOW) {
          if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
            matchedKf[bestIdxF] = (int)realIdxKF;
            if (checkOri) {
              float rot = kfAngle[realIdxKF] - F->kps[bestIdxF].angle;
              if (rot < 0.0) rot += 360.0f;
              int bin = (int)std::round(rot * factor);
              if (bin == HISTO_LENGTH) bin = 0;
              rotHist[bin].push_back(bestIdxF);
            }
            nmatches++;
          }
        }
      }
      ++a; ++b;
    } else if (kfIds[a] < fIds[b]) {
      ++a;  // lower_bound(Fit->first) on a sorted map == advance until >=
    } else {
      ++b;
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rotHist[i]) { matchedKf[idx] = -1; nmatches--; }
    }
  }
  return nmatches;
}


// ---- the remaining OrbMatcher searches.  As in include/orbfe.h the cv::Mat geometry that precedes
// GetFeaturesInArea is the caller's: valid[i] says the map point passed those gates, (u, v) is its projection. ----


// OrbMatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (orb_matcher.cpp:384-497)
int orc_search_by_projection_sim3(const orc_frame* KF, int nMP, const uint8_t* valid, const float* us, const float* vs,
                                  const int* predLevel, const uint8_t* mpDesc, const uint8_t* matchedIn, int th, int* matched) {
  const int TH_LOW = 50;
  std::vector<uint8_t> vpMatched(matchedIn, matchedIn + KF->n);
  for (int i = 0; i < KF->n; ++i) matched[i] = -1;
  int nmatches = 0;
  for (int iMP = 0; iMP < nMP; iMP++) {
    if (!valid[iMP]) continue;                                 // :411-449
    const int nPredictedLevel = predLevel[iMP];
    const float radius = th * KF->scale[nPredictedLevel];      // :454
    const std::vector<size_t> vIndices = features_in_area(KF, us[iMP], vs[iMP], radius, -1, -1);
    if (vIndices.empty()) continue;
    const uint8_t* dMP = mpDesc + (size_t)iMP * 32;
    int bestDist = 256, bestIdx = -1;
    for (size_t idx : vIndices) {
      if (vpMatched[idx]) continue;                            // :469
      const int kpLevel = KF->kps[idx].octave;
      if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
      const int dist = descriptor_distance(dMP, KF->desc.data() + idx * 32);
      if (dist < bestDist) { bestDist = dist; bestIdx = (int)idx; }
    }
    if (bestDist <= TH_LOW) { vpMatched[bestIdx] = 1; matched[bestIdx] = iMP; nmatches++; }
  }
  return nmatches;
}

// OrbMatcher::SearchByProjection(Frame& Cur, KeyFrame*, sAlreadyFound, th, ORBdist) (orb_matcher.cpp:1455-1582)
int orc_search_by_projection_keyframe(const orc_frame* C, int nKF, const uint8_t* valid, const float* us, const float* vs,
                                      const int* predLevel, const float* kfAngle, const uint8_t* mpDesc,
                                      const uint8_t* occupied, float th, int ORBdist, int checkOri, int* assigned) {
  int nmatches = 0;
  std::vector<uint8_t> hasMP(occupied, occupied + C->n);
  for (int i = 0; i < C->n; ++i) assigned[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  for (int i = 0; i < nKF; i++) {
    if (!valid[i]) continue;
    const float u = us[i], v = vs[i];
    if (u < C->minX || u > C->maxX) continue;                  // :1490-1495
    if (v < C->minY || v > C->maxY) continue;
    const int nPredictedLevel = predLevel[i];
    const float radius = th * C->scale[nPredictedLevel];       // :1511
    const std::vector<size_t> vIndices2 = features_in_area(C, u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1);
    if (vIndices2.empty()) continue;
    const uint8_t* dMP = mpDesc + (size_t)i * 32;
    int bestDist = 256, bestIdx2 = -1;
    for (size_t i2 : vIndices2) {
      if (hasMP[i2]) continue;                                 // :1526
      const int dist = descriptor_distance(dMP, C->desc.data() + i2 * 32);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = (int)i2; }
    }
    if (bestDist <= ORBdist) {
      hasMP[bestIdx2] = 1; assigned[bestIdx2] = i; ++nmatches;
      if (checkOri) {
        float rot = kfAngle[i] - C->kps[bestIdx2].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(bestIdx2);
      }
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++)
      if (i != ind1 && i != ind2 && i != ind3)
        for (int idx : rotHist[i]) { assigned[idx] = -1; --nmatches; }
  }
  return nmatches;
}

// the scan of OrbMatcher::Fuse(KeyFrame*, vpMapPoints, th) (orb_matcher.cpp:866-931; ur != NULL) and of
// Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint) (:1025-1061; ur == NULL): best keypoint per map point
int orc_fuse(const orc_frame* KF, int nMP, const uint8_t* valid, const float* us, const float* vs, const float* urs,
             const int* predLevel, const uint8_t* mpDesc, float th, int* bestIdxOut) {
  const int TH_LOW = 50;
  int nFused = 0;
  for (int i = 0; i < nMP; i++) {
    bestIdxOut[i] = -1;
    if (!valid[i]) continue;
    const float u = us[i], v = vs[i];
    const int nPredictedLevel = predLevel[i];
    const float radius = th * KF->scale[nPredictedLevel];
    const std::vector<size_t> vIndices = features_in_area(KF, u, v, radius, -1, -1);
    if (vIndices.empty()) continue;
    const uint8_t* dMP = mpDesc + (size_t)i * 32;
    int bestDist = urs ? 256 : INT_MAX, bestIdx = -1;
    for (size_t idx : vIndices) {
      const KP& kp = KF->kps[idx];
      const int kpLevel = kp.octave;
      if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
      if (urs) {
        const float invSigma2 = 1.0f / (kpLevel == 0 ? 1.0f : KF->scale[kpLevel] * KF->scale[kpLevel]);  // orb_extractor.cpp:363,372
        if (KF->uR[idx] >= 0) {  // reprojection error in stereo (:893-906)
          const float ex = u - kp.x, ey = v - kp.y, er = urs[i] - KF->uR[idx];
          const float e2 = ex * ex + ey * ey + er * er;
          if (e2 * invSigma2 > 7.8) continue;
        } else {
          const float ex = u - kp.x, ey = v - kp.y;
          const float e2 = ex * ex + ey * ey;
          if (e2 * invSigma2 > 5.99) continue;
        }
      }
      const int dist = descriptor_distance(dMP, KF->desc.data() + idx * 32);
      if (dist < bestDist) { bestDist = dist; bestIdx = (int)idx; }
    }
    if (bestDist <= TH_LOW) { bestIdxOut[i] = bestIdx; nFused++; }
  }
  return nFused;
}

// OrbMatcher::SearchBySim3 (orb_matcher.cpp:1081-1310)
int orc_search_by_sim3(const orc_frame* KF1, const orc_frame* KF2, const uint8_t* valid1, const float* u1, const float* v1,
                       const int* lvl1, const uint8_t* desc1, const uint8_t* valid2, const float* u2, const float* v2,
                       const int* lvl2, const uint8_t* desc2, float th, int* match12) {
  const int TH_HIGH = 100;
  const int N1 = KF1->n, N2 = KF2->n;
  std::vector<int> vnMatch1(N1, -1), vnMatch2(N2, -1);
  auto side = [&](const orc_frame* dst, int N, const uint8_t* valid, const float* us, const float* vs, const int* lvl,
                  const uint8_t* mpDesc, std::vector<int>& vnMatch) {
    for (int i = 0; i < N; i++) {
      if (!valid[i]) continue;
      const int nPredictedLevel = lvl[i];
      const float radius = th * dst->scale[nPredictedLevel];
      const std::vector<size_t> vIndices = features_in_area(dst, us[i], vs[i], radius, -1, -1);
      if (vIndices.empty()) continue;
      const uint8_t* dMP = mpDesc + (size_t)i * 32;
      int bestDist = INT_MAX, bestIdx = -1;
      for (size_t idx : vIndices) {
        const KP& kp = dst->kps[idx];
        if (kp.octave < nPredictedLevel - 1 || kp.octave > nPredictedLevel) continue;
        const int dist = descriptor_distance(dMP, dst->desc.data() + idx * 32);
        if (dist < bestDist) { bestDist = dist; bestIdx = (int)idx; }
      }
      if (bestDist <= TH_HIGH) vnMatch[i] = bestIdx;
    }
  };
  side(KF2, N1, valid1, u1, v1, lvl1, desc1, vnMatch1);  // :1132-1209
  side(KF1, N2, valid2, u2, v2, lvl2, desc2, vnMatch2);  // :1212-1289
  int nFound = 0;
  for (int i1 = 0; i1 < N1; i1++) {                      // :1291-1307
    match12[i1] = -1;
    const int idx2 = vnMatch1[i1];
    if (idx2 >= 0 && vnMatch2[idx2] == i1) { match12[i1] = idx2; nFound++; }
  }
  return nFound;
}

// OrbMatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (orb_matcher.cpp:499-632)
int orc_search_by_bow_keyframes(const orc_frame* KF2, int n1, const uint8_t* desc1, const float* angle1, const uint8_t* valid1,
                                const uint8_t* valid2, int nodes1, const uint32_t* ids1, const int* start1, const uint32_t* idx1v,
                                int nodes2, const uint32_t* ids2, const int* start2, const uint32_t* idx2v, float nnratio,
                                int checkOri, int* matches12) {
  const int TH_LOW = 50;
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  std::vector<uint8_t> vbMatched2(KF2->n, 0);
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  int nmatches = 0;
  const FV A{nodes1, ids1, start1, idx1v}, B{nodes2, ids2, start2, idx2v};
  for_common_nodes(A, B, [&](int a, int b) {
    for (int k1 = start1[a]; k1 < start1[a + 1]; ++k1) {
      const size_t idx1 = idx1v[k1];
      if (!valid1[idx1]) continue;                                    // :535-539
      const uint8_t* d1 = desc1 + idx1 * 32;
      int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
      for (int k2 = start2[b]; k2 < start2[b + 1]; ++k2) {
        const size_t idx2 = idx2v[k2];
        if (vbMatched2[idx2] || !valid2[idx2]) continue;              // :553-557
        const int dist = descriptor_distance(d1, KF2->desc.data() + idx2 * 32);
        if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = (int)idx2; }
        else if (dist < bestDist2) bestDist2 = dist;
      }
      if (bestDist1 < TH_LOW) {                                       // strict here (:575), '<=' in the Frame overload
        if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
          matches12[idx1] = bestIdx2;
          vbMatched2[bestIdx2] = 1;
          if (checkOri) {
            float rot = angle1[idx1] - KF2->kps[bestIdx2].angle;
            if (rot < 0.0) rot += 360.0f;
            int bin = (int)std::round(rot * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            rotHist[bin].push_back((int)idx1);
          }
          nmatches++;
        }
      }
    }
  });
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rotHist[i]) { matches12[idx] = -1; nmatches--; }
    }
  }
  return nmatches;
}

// OrbMatcher::SearchForTriangulation (orb_matcher.cpp:634-802) with CheckDistEpipolarLine (:114-131).  The
// reference never sets vbMatched2[bestIdx2] (the assignment of upstream ORB-SLAM2 is absent at :737-741), so a
// KeyFrame-2 feature may be paired with several KeyFrame-1 features; restated as is.
int orc_search_for_triangulation(const orc_frame* KF2, int n1, const orc_keypoint* kps1, const uint8_t* desc1,
                                 const uint8_t* valid1, const uint8_t* stereo1, const uint8_t* valid2, int nodes1,
                                 const uint32_t* ids1, const int* start1, const uint32_t* idx1v, int nodes2,
                                 const uint32_t* ids2, const int* start2, const uint32_t* idx2v, const float* F12, float ex,
                                 float ey, int onlyStereo, int checkOri, int* matches12) {
  const int TH_LOW = 50;
  int nmatches = 0;
  std::vector<uint8_t> vbMatched2(KF2->n, 0);
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  const FV A{nodes1, ids1, start1, idx1v}, B{nodes2, ids2, start2, idx2v};
  for_common_nodes(A, B, [&](int a, int b) {
    for (int k1 = start1[a]; k1 < start1[a + 1]; ++k1) {
      const size_t idx1 = idx1v[k1];
      if (!valid1[idx1]) continue;                 // already a MapPoint (:681)
      const bool bStereo1 = stereo1[idx1] != 0;
      if (onlyStereo && !bStereo1) continue;
      const orc_keypoint& kp1 = kps1[idx1];
      const uint8_t* d1 = desc1 + idx1 * 32;
      int bestDist = TH_LOW, bestIdx2 = -1;
      for (int k2 = start2[b]; k2 < start2[b + 1]; ++k2) {
        const size_t idx2 = idx2v[k2];
        if (vbMatched2[idx2] || !valid2[idx2]) continue;
        const bool bStereo2 = KF2->uR[idx2] >= 0;
        if (onlyStereo && !bStereo2) continue;
        const int dist = descriptor_distance(d1, KF2->desc.data() + idx2 * 32);
        if (dist > TH_LOW || dist > bestDist) continue;
        const KP& kp2 = KF2->kps[idx2];
        if (!bStereo1 && !bStereo2) {
          const float distex = ex - kp2.x, distey = ey - kp2.y;
          if (distex * distex + distey * distey < 100 * KF2->scale[kp2.octave]) continue;
        }
        // CheckDistEpipolarLine (:114-131): l = x1' F12
        const float la = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
        const float lb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
        const float lc = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
        const float num = la * kp2.x + lb * kp2.y + lc;
        const float den = la * la + lb * lb;
        if (den == 0) continue;
        const float dsqr = num * num / den;
        const float sigma2 = kp2.octave == 0 ? 1.0f : KF2->scale[kp2.octave] * KF2->scale[kp2.octave];  // orb_extractor.cpp:363
        if (dsqr < 3.84 * sigma2) { bestIdx2 = (int)idx2; bestDist = dist; }
      }
      if (bestIdx2 >= 0) {
        matches12[idx1] = bestIdx2;
        nmatches++;
        if (checkOri) {
          float rot = kp1.angle - KF2->kps[bestIdx2].angle;
          if (rot < 0.0) rot += 360.0f;
          int bin = (int)std::round(rot * factor);
          if (bin == HISTO_LENGTH) bin = 0;
          rotHist[bin].push_back((int)idx1);
        }
      }
    }
  });
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rotHist[i]) { matches12[idx] = -1; nmatches--; }
    }
  }
  return nmatches;
}

}  // extern "C"

// ===========================================================================================
// N3: DBoW2 TemplatedVocabulary<FORB>::transform (third_party/DBoW2/DBoW2/TemplatedVocabulary.h) restated with the
// reference's own containers (BowVector = std::map<WordId, double>, FeatureVector = std::map<NodeId, vector<unsigned>>).
#include <map>
struct orc_vocabulary {
  struct Node {                       // TemplatedVocabulary.h:295-322
    double weight = 0;
    std::vector<unsigned> children;
    uint8_t descriptor[32] = {0};
    unsigned word_id = 0;
    bool isLeaf() const { return children.empty(); }
  };
  int k = 0, L = 0, scoring = 0, weighting = 0;
  std::vector<Node> nodes;
  int nwords = 0;
};

namespace {
// transform(feature, word_id, weight, nid, levelsup) (TemplatedVocabulary.h:1198-1248)
void voc_transform_one(const orc_vocabulary* V, const uint8_t* feature, unsigned& word_id, double& weight, unsigned* nid, int levelsup) {
  const int nid_level = V->L - levelsup;
  if (nid_level <= 0 && nid != nullptr) *nid = 0;  // root
  unsigned final_id = 0;
  int current_level = 0;
  do {
    ++current_level;
    const std::vector<unsigned>& nodes = V->nodes[final_id].children;
    final_id = nodes[0];
    double best_d = descriptor_distance(feature, V->nodes[final_id].descriptor);  // FORB::distance (FORB.cpp:81-101)
    for (size_t c = 1; c < nodes.size(); ++c) {
      const unsigned id = nodes[c];
      const double d = descriptor_distance(feature, V->nodes[id].descriptor);
      if (d < best_d) { best_d = d; final_id = id; }
    }
    if (nid != nullptr && current_level == nid_level) *nid = final_id;
  } while (!V->nodes[final_id].isLeaf());
  word_id = V->nodes[final_id].word_id;
  weight = V->nodes[final_id].weight;
}
}  // namespace

extern "C" {

// the tree loadFromTextFile builds (TemplatedVocabulary.h:1372-1417)
orc_vocabulary* orc_vocabulary_create(int k, int L, int scoring, int weighting, int n_nodes, const int* parent,
                                      const uint8_t* is_leaf, const uint8_t* desc, const double* weight) {
  orc_vocabulary* V = new orc_vocabulary();
  V->k = k; V->L = L; V->scoring = scoring; V->weighting = weighting;
  V->nodes.resize(n_nodes);
  for (int nid = 1; nid < n_nodes; ++nid) {
    V->nodes[parent[nid]].children.push_back((unsigned)nid);
    std::memcpy(V->nodes[nid].descriptor, desc + (size_t)nid * 32, 32);
    V->nodes[nid].weight = weight[nid];
    if (is_leaf[nid]) V->nodes[nid].word_id = (unsigned)V->nwords++;
  }
  return V;
}
void orc_vocabulary_destroy(orc_vocabulary* V) { delete V; }

// transform(features, BowVector&, FeatureVector&, levelsup) (TemplatedVocabulary.h:1124-1190) with BowVector::addWeight /
// addIfNotExist / normalize (BowVector.cpp:34-87) and FeatureVector::addFeature (FeatureVector.cpp:32-47).
// Where the reference leaves `nid` unwritten (leaf above nid_level) it is indeterminate there; defined as 0 here.
int orc_bow_transform(const orc_vocabulary* V, int n, const uint8_t* desc, int levelsup, unsigned* word_out, unsigned* node_out,
                      unsigned* bow_words, double* bow_values, int* n_bow, unsigned* fv_nodes, int* fv_start, unsigned* fv_idx,
                      int* n_fv) {
  std::map<unsigned, double> v;
  std::map<unsigned, std::vector<unsigned>> fv;
  *n_bow = 0; *n_fv = 0; fv_start[0] = 0;
  if (V->nwords == 0) return 0;  // empty()
  // mustNormalize (ScoringObject.h:73-90)
  const bool must = V->scoring != 5;
  const bool l2 = V->scoring == 1;
  for (int i = 0; i < n; ++i) {
    unsigned id = 0, nid = 0;
    double w = 0;
    voc_transform_one(V, desc + (size_t)i * 32, id, w, &nid, levelsup);
    if (word_out) word_out[i] = id;
    if (node_out) node_out[i] = nid;
    if (w > 0) {  // not stopped
      if (V->weighting == 0 || V->weighting == 1) {  // TF_IDF, TF: addWeight
        auto vit = v.lower_bound(id);
        if (vit != v.end() && !(v.key_comp()(id, vit->first))) vit->second += w;
        else v.insert(vit, std::make_pair(id, w));
      } else {  // IDF, BINARY: addIfNotExist
        auto vit = v.lower_bound(id);
        if (vit == v.end() || v.key_comp()(id, vit->first)) v.insert(vit, std::make_pair(id, w));
      }
      fv[nid].push_back((unsigned)i);
    }
  }
  if ((V->weighting == 0 || V->weighting == 1) && !v.empty() && !must) {
    const double nd = (double)v.size();
    for (auto& e : v) e.second /= nd;
  }
  if (must) {  // BowVector::normalize
    double norm = 0.0;
    if (!l2) { for (auto& e : v) norm += std::fabs(e.second); }
    else { for (auto& e : v) norm += e.second * e.second; norm = std::sqrt(norm); }
    if (norm > 0.0) for (auto& e : v) e.second /= norm;
  }
  int u = 0;
  for (auto& e : v) { bow_words[u] = e.first; bow_values[u] = e.second; ++u; }
  *n_bow = u;
  int f = 0, p = 0;
  for (auto& e : fv) {
    fv_nodes[f] = e.first;
    fv_start[f] = p;
    for (unsigned idx : e.second) fv_idx[p++] = idx;
    ++f;
  }
  fv_start[f] = p;
  *n_fv = f;
  return u;
}

}  // extern "C"

// ===========================================================================================
// N2: the Frame tail -- Frame::UndistortKeyPoints (frame.cpp:614-641) and Frame::IsInFrustum (frame.cpp:277-337)
// with MapPoint::PredictScale (map_point.cpp:382-396).  The cv::Mat arithmetic is restated from OpenCV and pinned
// against cv2 (tests/golden/cv2_frame_tail.npz): gemm 3x3 * 3x1 + 3x1 in float (products and sums in float, in k
// order), cv::norm with a double accumulator; Mat::dot (double accumulator, not exposed through cv2) is restated
// from modules/core/src/matmul (dotProd_) and is unpinned.
extern "C" {

// cv::undistortPoints(src, dst, K, dist, cv::Mat(), K) for CV_32FC2 points (calib3d cvUndistortPointsInternal, 5 fixed
// iterations, double arithmetic); K = fx, fy, cx, cy as floats (calib_mat_ is CV_32F); dist = n_dist (4, 5, 8, 12 or 14)
// float coefficients k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4 [tauX tauY]]]] (tilt must be 0).
// Frame::UndistortKeyPoints copies the keypoints unchanged when dist[0] == 0 (frame.cpp:616-619).
void orc_undistort_points(int n, const float* xy_in, float fxf, float fyf, float cxf, float cyf, const float* dist, int n_dist,
                          float* xy_out) {
  if (n_dist < 1 || dist[0] == 0.0f) { std::memcpy(xy_out, xy_in, (size_t)n * 2 * sizeof(float)); return; }
  double k[14] = {0};
  for (int i = 0; i < n_dist && i < 14; ++i) k[i] = dist[i];
  const double fx = fxf, fy = fyf, cx = cxf, cy = cyf, ifx = 1. / fx, ify = 1. / fy;
  // RR = P * R with R = I and P = K (cvMatMul on 3x3 doubles: sum over k in order; zeros and ones are exact)
  const double RR[3][3] = {{fx, 0, cx}, {0, fy, cy}, {0, 0, 1}};
  for (int i = 0; i < n; ++i) {
    double x = xy_in[2 * i], y = xy_in[2 * i + 1];
    const double u = x, v = y;
    x = (x - cx) * ifx;
    y = (y - cy) * ify;
    const double x0 = x, y0 = y;  // the tilt compensation is the identity for tauX = tauY = 0
    for (int j = 0; j < 5; ++j) {
      const double r2 = x * x + y * y;
      const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
      if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
      const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
      const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
      x = (x0 - deltaX) * icdist;
      y = (y0 - deltaY) * icdist;
    }
    const double xx = RR[0][0] * x + RR[0][1] * y + RR[0][2];
    const double yy = RR[1][0] * x + RR[1][1] * y + RR[1][2];
    const double ww = 1. / (RR[2][0] * x + RR[2][1] * y + RR[2][2]);
    xy_out[2 * i] = (float)(xx * ww);
    xy_out[2 * i + 1] = (float)(yy * ww);
  }
}

// Frame::IsInFrustum for n map points (the loop of Tracker::SearchLocalPoints, core/tracker.cpp:1196-1211).
// world/normal: n x 3; min_dist / max_dist = Get{Min,Max}DistanceInvariance() (0.8f * min_dist_, 1.2f * max_dist_), max_dist_raw =
// MapPoint::max_dist_ itself; Rcw row-major; outputs = the track_* fields (frame.cpp:328-334).  Returns the number in view.
int orc_is_in_frustum(int n, const float* world, const float* normal, const float* min_dist, const float* max_dist,
                      const float* max_dist_raw,
                      const float* Rcw, const float* tcw, const float* Ow, float fx, float fy, float cx, float cy, float bf,
                      float min_x, float max_x, float min_y, float max_y, float log_scale_factor, int n_levels,
                      float viewing_cos_limit, uint8_t* in_view, float* proj_x, float* proj_y, float* proj_xr, int* level,
                      float* view_cos) {
  int count = 0;
  for (int i = 0; i < n; ++i) {
    in_view[i] = 0; proj_x[i] = 0; proj_y[i] = 0; proj_xr[i] = 0; level[i] = 0; view_cos[i] = 0;
    const float* P = world + 3 * (size_t)i;
    float Pc[3];
    for (int r = 0; r < 3; ++r) {  // Rcw_*P + tcw_ (cv::gemm small-matrix path: float products, float sums)
      const float t0 = Rcw[3 * r] * P[0] + Rcw[3 * r + 1] * P[1] + Rcw[3 * r + 2] * P[2];
      Pc[r] = (float)((double)t0 * 1.0 + (double)tcw[r] * 1.0);
    }
    const float PcX = Pc[0], PcY = Pc[1], PcZ = Pc[2];
    if (PcZ < 0.0f) continue;
    const float invz = 1.0f / PcZ;
    const float u = fx * PcX * invz + cx;
    const float v = fy * PcY * invz + cy;
    if (u < min_x || u > max_x) continue;
    if (v < min_y || v > max_y) continue;
    const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
    double s = 0;  // cv::norm(PO): normL2Sqr with a double accumulator, then sqrt
    for (int c = 0; c < 3; ++c) { const double e = PO[c]; s += e * e; }
    const float dist = (float)std::sqrt(s);
    if (dist < min_dist[i] || dist > max_dist[i]) continue;
    const float* Pn = normal + 3 * (size_t)i;
    double dot = 0;  // Mat::dot: dotProd_ accumulates (double)a*b
    for (int c = 0; c < 3; ++c) dot += (double)PO[c] * Pn[c];
    const float viewCos = (float)(dot / dist);
    if (viewCos < viewing_cos_limit) continue;
    // MapPoint::PredictScale (map_point.cpp:382-396)
    const float ratio = max_dist_raw[i] / dist;  // PredictScale divides the RAW max_dist_, not 1.2f * max_dist_ (map_point.cpp:386 vs :363)
    int nScale = (int)std::ceil(std::log(ratio) / log_scale_factor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= n_levels) nScale = n_levels - 1;
    in_view[i] = 1;
    proj_x[i] = u;
    proj_xr[i] = u - bf * invz;
    proj_y[i] = v;
    level[i] = nScale;
    view_cos[i] = viewCos;
    ++count;
  }
  return count;
}

}  // extern "C"
