// k_frame.cuh -- N2, the Frame tail either side of the extractor / matchers:
//   k_undistort_points  Frame::UndistortKeyPoints (frame.cpp:614-641) = cv::undistortPoints(mat, mat, K, dist, Mat(), K):
//                       5 fixed-point iterations per keypoint in double (calib3d cvUndistortPointsInternal), one thread
//                       per keypoint, every operation individually rounded (no FMA) so that it equals the CPU library.
//   k_is_in_frustum     Frame::IsInFrustum (frame.cpp:277-337) + MapPoint::PredictScale (map_point.cpp:382-396) for every
//                       local map point at once (the loop of Tracker::SearchLocalPoints, core/tracker.cpp:1196-1211); the
//                       outputs are exactly the arrays orbfe_search_by_projection_mappoints consumes.
// OpenCV arithmetic pinned against cv2 (tests/golden/cv2_frame_tail.npz): gemm 3x3*3x1+3x1 = float products and float sums
// in k order; cv::norm = double accumulator + sqrt; Mat::dot = double accumulator.  std::log(float) = glibc logf, restated
// below (sysdeps/ieee754/flt-32/e_logf.c, glibc >= 2.27; checked against libm on 22M inputs, tests/test_oracle_primitives.py).
#pragma once
#include "orbfe_common.cuh"
// (kernels are `static`: this header is included by orbfe_frame.cu and by orbfe_match.cu, which chains the frustum test into
// the projection search without a host round trip)

__device__ __forceinline__ double orbfe_dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double orbfe_da(double a, double b) { return __dadd_rn(a, b); }

__device__ __forceinline__ float orbfe_glibc_logf(float x) {
  // {invc, logc} for the 16 sub-intervals of [0x1.66p-1, 0x1.66p0) (glibc __logf_data)
  const double T[16][2] = {
      {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
      {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
      {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
      {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
      {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
      {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
      {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
      {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
  const double Ln2 = 0x1.62e42fefa39efp-1;
  const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
  unsigned ix = __float_as_uint(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2u == 0u) return -__int_as_float(0x7f800000);                                       // log(+-0) = -inf
    if (ix == 0x7f800000u) return x;                                                              // log(inf) = inf
    if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return __int_as_float(0x7fc00000);          // negative / NaN
    ix = __float_as_uint(__fmul_rn(x, 8388608.0f));                                               // subnormal: normalise
    ix -= 23u << 23;
  }
  const unsigned tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) & 15u);
  const int k = (int)tmp >> 23;
  const unsigned iz = ix - (tmp & (0x1ffu << 23));
  const double z = (double)__uint_as_float(iz);
  const double r = orbfe_da(orbfe_dm(z, T[i][0]), -1.0);
  const double y0 = orbfe_da(T[i][1], orbfe_dm((double)k, Ln2));
  const double r2 = orbfe_dm(r, r);
  double y = orbfe_da(orbfe_dm(A1, r), A2);
  y = orbfe_da(orbfe_dm(A0, r2), y);
  y = orbfe_da(orbfe_dm(y, r2), orbfe_da(y0, r));
  return (float)y;
}

struct UndistortArgs {
  double fx, fy, cx, cy, ifx, ify;
  double k[14];
};

static __global__ void __launch_bounds__(256)
k_undistort_points(const UndistortArgs U, const int n, const float* __restrict__ kpIn, float* __restrict__ kpOut,
                   const int strideFloats) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float* in = kpIn + (size_t)i * strideFloats;
  float* out = kpOut + (size_t)i * strideFloats;
  for (int c = 2; c < strideFloats; ++c) out[c] = in[c];  // the other cv::KeyPoint fields are copied (frame.cpp:634-639)
  const double u = in[0], v = in[1];
  double x = orbfe_dm(orbfe_da(u, -U.cx), U.ifx), y = orbfe_dm(orbfe_da(v, -U.cy), U.ify);
  const double x0 = x, y0 = y;
  const double* k = U.k;
  for (int j = 0; j < 5; ++j) {
    const double r2 = orbfe_da(orbfe_dm(x, x), orbfe_dm(y, y));
    const double num = orbfe_da(1.0, orbfe_dm(orbfe_da(orbfe_dm(orbfe_da(orbfe_dm(k[7], r2), k[6]), r2), k[5]), r2));
    const double den = orbfe_da(1.0, orbfe_dm(orbfe_da(orbfe_dm(orbfe_da(orbfe_dm(k[4], r2), k[1]), r2), k[0]), r2));
    const double icdist = __ddiv_rn(num, den);
    if (icdist < 0) {
      x = orbfe_dm(orbfe_da(u, -U.cx), U.ifx);
      y = orbfe_dm(orbfe_da(v, -U.cy), U.ify);
      break;
    }
    // deltaX = 2*k[2]*x*y + k[3]*(r2 + 2*x*x) + k[8]*r2 + k[9]*r2*r2  (left to right)
    const double dX = orbfe_da(orbfe_da(orbfe_da(orbfe_dm(orbfe_dm(orbfe_dm(2.0, k[2]), x), y),
                                                 orbfe_dm(k[3], orbfe_da(r2, orbfe_dm(orbfe_dm(2.0, x), x)))),
                                        orbfe_dm(k[8], r2)),
                               orbfe_dm(orbfe_dm(k[9], r2), r2));
    // deltaY = k[2]*(r2 + 2*y*y) + 2*k[3]*x*y + k[10]*r2 + k[11]*r2*r2
    const double dY = orbfe_da(orbfe_da(orbfe_da(orbfe_dm(k[2], orbfe_da(r2, orbfe_dm(orbfe_dm(2.0, y), y))),
                                                 orbfe_dm(orbfe_dm(orbfe_dm(2.0, k[3]), x), y)),
                                        orbfe_dm(k[10], r2)),
                               orbfe_dm(orbfe_dm(k[11], r2), r2));
    x = orbfe_dm(orbfe_da(x0, -dX), icdist);
    y = orbfe_dm(orbfe_da(y0, -dY), icdist);
  }
  // RR = K: xx = fx*x + 0*y + cx, yy = 0*x + fy*y + cy, ww = 1/(0*x + 0*y + 1)
  const double xx = orbfe_da(orbfe_da(orbfe_dm(U.fx, x), orbfe_dm(0.0, y)), U.cx);
  const double yy = orbfe_da(orbfe_da(orbfe_dm(0.0, x), orbfe_dm(U.fy, y)), U.cy);
  const double ww = __ddiv_rn(1.0, orbfe_da(orbfe_da(orbfe_dm(0.0, x), orbfe_dm(0.0, y)), 1.0));
  out[0] = (float)orbfe_dm(xx, ww);
  out[1] = (float)orbfe_dm(yy, ww);
}

struct FrustumArgs {
  float R[9], t[3], Ow[3];
  float fx, fy, cx, cy, bf, minX, maxX, minY, maxY, logScaleFactor, viewingCosLimit;
  int nLevels;
};

static __global__ void __launch_bounds__(256)
k_is_in_frustum(const FrustumArgs A, const int n, const float* __restrict__ world, const float* __restrict__ normal,
                const float* __restrict__ minDist, const float* __restrict__ maxDist, const float* __restrict__ maxDistRaw,
                uint8_t* __restrict__ inView,
                float* __restrict__ projX, float* __restrict__ projY, float* __restrict__ projXR, int* __restrict__ level,
                float* __restrict__ viewCosOut, int* __restrict__ count) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  bool ok = i < n;
  float u = 0.f, v = 0.f, ur = 0.f, viewCos = 0.f;
  int nScale = 0;
  if (ok) {
    const float Px = world[3 * (size_t)i], Py = world[3 * (size_t)i + 1], Pz = world[3 * (size_t)i + 2];
    float Pc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)  // Rcw_*P + tcw_
      Pc[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(A.R[3 * r], Px), __fmul_rn(A.R[3 * r + 1], Py)), __fmul_rn(A.R[3 * r + 2], Pz)), A.t[r]);
    if (Pc[2] < 0.0f) ok = false;
    const float invz = __fdiv_rn(1.0f, Pc[2]);
    u = __fadd_rn(__fmul_rn(__fmul_rn(A.fx, Pc[0]), invz), A.cx);
    v = __fadd_rn(__fmul_rn(__fmul_rn(A.fy, Pc[1]), invz), A.cy);
    if (u < A.minX || u > A.maxX || v < A.minY || v > A.maxY) ok = false;
    const float POx = __fsub_rn(Px, A.Ow[0]), POy = __fsub_rn(Py, A.Ow[1]), POz = __fsub_rn(Pz, A.Ow[2]);
    const double s = orbfe_da(orbfe_da(orbfe_dm(POx, POx), orbfe_dm(POy, POy)), orbfe_dm(POz, POz));
    const float dist = (float)__dsqrt_rn(s);
    if (ok && (dist < minDist[i] || dist > maxDist[i])) ok = false;
    if (ok) {
      const float nx = normal[3 * (size_t)i], ny = normal[3 * (size_t)i + 1], nz = normal[3 * (size_t)i + 2];
      const double dot = orbfe_da(orbfe_da(orbfe_dm(POx, nx), orbfe_dm(POy, ny)), orbfe_dm(POz, nz));
      viewCos = (float)__ddiv_rn(dot, (double)dist);
      if (viewCos < A.viewingCosLimit) ok = false;
    }
    if (ok) {
      const float ratio = __fdiv_rn(maxDistRaw[i], dist);  // max_dist_ itself, not the 1.2x invariance bound (map_point.cpp:386)
      nScale = (int)ceilf(__fdiv_rn(orbfe_glibc_logf(ratio), A.logScaleFactor));
      if (nScale < 0) nScale = 0;
      else if (nScale >= A.nLevels) nScale = A.nLevels - 1;
      ur = __fsub_rn(u, __fmul_rn(A.bf, invz));
    }
  }
  if (i < n) {
    inView[i] = ok ? 1 : 0;
    projX[i] = ok ? u : 0.f;
    projY[i] = ok ? v : 0.f;
    projXR[i] = ok ? ur : 0.f;
    level[i] = ok ? nScale : 0;
    viewCosOut[i] = ok ? viewCos : 0.f;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(count, __popc(bal));
}

// parity tap: the logf restatement over an array (tests sweep it against libm)
static __global__ void __launch_bounds__(256)
k_debug_logf(const float* __restrict__ x, const int n, float* __restrict__ y) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) y[i] = orbfe_glibc_logf(x[i]);
}

// Tracker::SearchLocalPoints glue (core/tracker.cpp:1213-1226 -> orb_matcher.cpp:29-52): the frustum outputs become the window
// queries of SearchByProjection(Frame&, vpMapPoints, th) on the device: radius = RadiusByViewingCos(viewCos) [* th] *
// scale_factors[level], levels [level-1, level], stereo-right consistency against track_projected_x_right.
static __global__ void __launch_bounds__(256)
k_frustum_to_queries(const int n, const uint8_t* __restrict__ inView, const float* __restrict__ projX, const float* __restrict__ projY,
                     const float* __restrict__ projXR, const int* __restrict__ level, const float* __restrict__ viewCos,
                     const float* __restrict__ scale, const int th, uint8_t* __restrict__ qValid, float* __restrict__ qx,
                     float* __restrict__ qy, float* __restrict__ qr, float* __restrict__ qxr, int* __restrict__ qMinL,
                     int* __restrict__ qMaxL) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const bool ok = inView[i] != 0;
  const int lvl = level[i];
  float r = ((double)viewCos[i] > 0.998) ? 2.5f : 4.0f;  // RadiusByViewingCos (orb_matcher.cpp:105-111)
  if (th != 1) r = __fmul_rn(r, (float)th);               // :43-44
  qValid[i] = ok ? 1 : 0;
  qx[i] = projX[i];
  qy[i] = projY[i];
  qr[i] = ok ? __fmul_rn(r, scale[lvl]) : 0.f;            // :46
  qxr[i] = projXR[i];
  qMinL[i] = lvl - 1;
  qMaxL[i] = lvl;
}
