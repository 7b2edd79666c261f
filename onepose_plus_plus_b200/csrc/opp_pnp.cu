// opp_pnp.cu — batched RANSAC-PnP on the device: the consumer of the matcher's output
// (reference: src/utils/metric_utils.py:121-204 `ransac_PnP` = cv2.solvePnPRansac(EPnP, 10000
// iterations, reprojectionError) per frame on the CPU after a D2H sync; called per frame from
// `compute_query_pose_errors` :207-292 and demo.py:132).
//
// One CTA per query image.  The matches of image b are the contiguous run of `m_bids == b` in the
// match lists (the matcher emits them in ascending (b, i) order).  Per image:
//   1. hypotheses: every thread draws 4 distinct matches (counter-based hash RNG), solves P3P on the
//      first three (Grunert's quartic in the depth ratio, closed-form Ferrari roots + Newton
//      polish, pose from the two triangle frames) and keeps the root that best reprojects the 4th;
//   2. scoring: reprojection error of every match under the hypothesis, inlier = err < thr and in
//      front of the camera; block-wide argmax of the inlier count (ties -> lowest hypothesis id, so
//      the result does not depend on scheduling);
//   3. refinement (local optimisation): Gauss-Newton / LM on the 6-DoF pose over the current
//      inliers (normal equations accumulated by the whole CTA, 6x6 Cholesky by one thread),
//      inlier set re-evaluated between rounds.  The optimum of the reprojection error over the
//      inliers is what cv2's iterative refinement converges to as well, which is what the parity
//      test compares against.
// All geometry runs in fp64 (a few MFLOP per image); inputs / outputs are fp32.
#include <cstdint>
#include <cstdio>

#include "../../include/opp_b200.h"
#include "opp_common.cuh"

namespace opp {

struct Pose {
  double R[9];
  double t[3];
};

__device__ __forceinline__ uint32_t pnp_hash(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

__host__ __device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
__host__ __device__ __forceinline__ double dot3(const double* a, const double* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
__host__ __device__ __forceinline__ bool normalize3(double* a) {
  const double n = sqrt(dot3(a, a));
  if (!(n > 1e-300)) return false;
  a[0] /= n;
  a[1] /= n;
  a[2] /= n;
  return true;
}

// largest real root of x^3 + a2 x^2 + a1 x + a0
__host__ __device__ double cubic_largest_real(double a2, double a1, double a0) {
  const double p = a1 - a2 * a2 / 3.0;
  const double q = 2.0 * a2 * a2 * a2 / 27.0 - a2 * a1 / 3.0 + a0;
  const double disc = q * q / 4.0 + p * p * p / 27.0;
  double y;
  if (disc > 0.0) {
    const double sq = sqrt(disc);
    y = cbrt(-q / 2.0 + sq) + cbrt(-q / 2.0 - sq);
  } else if (p < 0.0) {
    const double r = sqrt(-p / 3.0);
    double c = -q / 2.0 / (r * r * r);
    c = fmin(1.0, fmax(-1.0, c));
    y = 2.0 * r * cos(acos(c) / 3.0);
  } else {
    y = 0.0;
  }
  return y - a2 / 3.0;
}

// real roots of A4 x^4 + ... + A0 (Ferrari), each polished with Newton steps; returns the count
__host__ __device__ int solve_quartic(double A4, double A3, double A2, double A1, double A0, double* roots) {
  if (!(fabs(A4) > 1e-14)) return 0;
  const double b = A3 / A4, c = A2 / A4, d = A1 / A4, e = A0 / A4;
  const double p = c - 3.0 * b * b / 8.0;
  const double q = d - b * c / 2.0 + b * b * b / 8.0;
  const double r = e - b * d / 4.0 + b * b * c / 16.0 - 3.0 * b * b * b * b / 256.0;
  int n = 0;
  double y[4];
  const double m = cubic_largest_real(p, p * p / 4.0 - r, -q * q / 8.0);
  if (m > 1e-14) {
    const double s = sqrt(2.0 * m);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double sg = k == 0 ? 1.0 : -1.0;
      const double B = sg * s, C = p / 2.0 + m - sg * q / (2.0 * s);
      const double D = B * B - 4.0 * C;
      if (D >= 0.0) {
        const double sd = sqrt(D);
        y[n++] = (-B + sd) / 2.0;
        y[n++] = (-B - sd) / 2.0;
      } else if (D > -1e-9 * fmax(1.0, B * B)) {
        y[n++] = -B / 2.0;
      }
    }
  } else {
    const double D = p * p / 4.0 - r;   // biquadratic y^4 + p y^2 + r
    if (D >= 0.0) {
      const double sd = sqrt(D);
      const double z0 = -p / 2.0 + sd, z1 = -p / 2.0 - sd;
      if (z0 >= 0.0) {
        y[n++] = sqrt(z0);
        y[n++] = -sqrt(z0);
      }
      if (z1 >= 0.0 && n <= 2) {
        y[n++] = sqrt(z1);
        y[n++] = -sqrt(z1);
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    double x = y[i] - b / 4.0;
#pragma unroll
    for (int it = 0; it < 3; ++it) {
      const double f = (((x + b) * x + c) * x + d) * x + e;
      const double fp = ((4.0 * x + 3.0 * b) * x + 2.0 * c) * x + d;
      if (fabs(fp) > 1e-300) x -= f / fp;
    }
    roots[i] = x;
  }
  return n;
}

// orthonormal frame of a triangle: columns e1 = (Q1-Q0)^, e3 = (e1 x (Q2-Q0))^, e2 = e3 x e1
__host__ __device__ bool tri_frame(const double* Q0, const double* Q1, const double* Q2, double* F) {
  double e1[3] = {Q1[0] - Q0[0], Q1[1] - Q0[1], Q1[2] - Q0[2]};
  double w[3] = {Q2[0] - Q0[0], Q2[1] - Q0[1], Q2[2] - Q0[2]};
  double e3[3], e2[3];
  if (!normalize3(e1)) return false;
  cross3(e1, w, e3);
  if (!normalize3(e3)) return false;
  cross3(e3, e1, e2);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    F[k * 3 + 0] = e1[k];
    F[k * 3 + 1] = e2[k];
    F[k * 3 + 2] = e3[k];
  }
  return true;
}

// P3P (Grunert 1841 as restated by Haralick et al. 1994): world points P[3], unit bearings f[3].
// Calls `visit(pose)` for every admissible solution.
template <class V>
__host__ __device__ void p3p_grunert(const double (*P)[3], const double (*f)[3], V&& visit) {
  double d12[3], d02[3], d01[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    d12[k] = P[1][k] - P[2][k];
    d02[k] = P[0][k] - P[2][k];
    d01[k] = P[0][k] - P[1][k];
  }
  const double a2 = dot3(d12, d12), b2 = dot3(d02, d02), c2 = dot3(d01, d01);
  if (!(a2 > 1e-20 && b2 > 1e-20 && c2 > 1e-20)) return;
  const double ca = dot3(f[1], f[2]), cb = dot3(f[0], f[2]), cg = dot3(f[0], f[1]);
  const double q = (a2 - c2) / b2, ac = (a2 + c2) / b2;
  const double A4 = (q - 1.0) * (q - 1.0) - 4.0 * c2 / b2 * ca * ca;
  const double A3 = 4.0 * (q * (1.0 - q) * cb - (1.0 - ac) * ca * cg + 2.0 * c2 / b2 * ca * ca * cb);
  const double A2 = 2.0 * (q * q - 1.0 + 2.0 * q * q * cb * cb + 2.0 * ((b2 - c2) / b2) * ca * ca -
                           4.0 * ac * ca * cb * cg + 2.0 * ((b2 - a2) / b2) * cg * cg);
  const double A1 = 4.0 * (-q * (1.0 + q) * cb + 2.0 * a2 / b2 * cg * cg * cb - (1.0 - ac) * ca * cg);
  const double A0 = (1.0 + q) * (1.0 + q) - 4.0 * a2 / b2 * cg * cg;
  double roots[4];
  const int n = solve_quartic(A4, A3, A2, A1, A0, roots);
  double Fw[9];
  if (!tri_frame(P[0], P[1], P[2], Fw)) return;
  for (int i = 0; i < n; ++i) {
    const double v = roots[i];
    if (!(v > 0.0) || !isfinite(v)) continue;
    const double den = 2.0 * (cg - v * ca);
    if (!(fabs(den) > 1e-12)) continue;
    const double u = ((q - 1.0) * v * v - 2.0 * q * cb * v + 1.0 + q) / den;
    if (!(u > 0.0)) continue;
    const double s1sq = b2 / (1.0 + v * v - 2.0 * v * cb);
    if (!(s1sq > 0.0)) continue;
    const double s1 = sqrt(s1sq);
    double X[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      X[0][k] = s1 * f[0][k];
      X[1][k] = u * s1 * f[1][k];
      X[2][k] = v * s1 * f[2][k];
    }
    double Fc[9];
    if (!tri_frame(X[0], X[1], X[2], Fc)) continue;
    Pose ps;
    // R = Fc * Fw^T
#pragma unroll
    for (int r_ = 0; r_ < 3; ++r_)
#pragma unroll
      for (int c_ = 0; c_ < 3; ++c_)
        ps.R[r_ * 3 + c_] = Fc[r_ * 3 + 0] * Fw[c_ * 3 + 0] + Fc[r_ * 3 + 1] * Fw[c_ * 3 + 1] +
                            Fc[r_ * 3 + 2] * Fw[c_ * 3 + 2];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      ps.t[k] = X[0][k] - (ps.R[k * 3] * P[0][0] + ps.R[k * 3 + 1] * P[0][1] + ps.R[k * 3 + 2] * P[0][2]);
    visit(ps);
  }
}

struct Cam {
  double fx, fy, cx, cy, skew;
};

// squared reprojection error (pixels^2) of world point p against pixel (u, v); +inf behind the camera
__device__ __forceinline__ double reproj_err2(const Pose& ps, const Cam& cam, const double* p, double u,
                                              double v) {
  const double x = ps.R[0] * p[0] + ps.R[1] * p[1] + ps.R[2] * p[2] + ps.t[0];
  const double y = ps.R[3] * p[0] + ps.R[4] * p[1] + ps.R[5] * p[2] + ps.t[1];
  const double z = ps.R[6] * p[0] + ps.R[7] * p[1] + ps.R[8] * p[2] + ps.t[2];
  if (!(z > 1e-9)) return INFINITY;
  const double xn = x / z, yn = y / z;
  const double du = cam.fx * xn + cam.skew * yn + cam.cx - u;
  const double dv = cam.fy * yn + cam.cy - v;
  return du * du + dv * dv;
}

constexpr int kPnpThreads = 256;

__device__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < kPnpThreads / 32; ++w) s += red[w];
  return s;
}

__global__ void __launch_bounds__(kPnpThreads)
pnp_ransac_kernel(const float* __restrict__ pts3d, const float* __restrict__ pts2d,
                  const long long* __restrict__ m_bids, int M, const float* __restrict__ Kmat,
                  float scale, float thr, int n_hyp, unsigned seed, int refine_rounds,
                  float* __restrict__ pose_out, int* __restrict__ n_inl_out,
                  unsigned char* __restrict__ inl_mask, int* __restrict__ status_out) {
  pdl_sync();
  __shared__ int seg[2];
  __shared__ unsigned long long best_key[kPnpThreads / 32];
  __shared__ Pose best_pose;
  __shared__ double red[kPnpThreads / 32];
  __shared__ double Hs[27];
  __shared__ int flag;
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  // [lo, hi) = run of matches with m_bids == b (lower bounds of b and b + 1)
  if (tid < 2) {
    const long long key = b + tid;
    int lo = 0, hi = M;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (m_bids[mid] < key) lo = mid + 1; else hi = mid;
    }
    seg[tid] = lo;
  }
  __syncthreads();
  const int lo = seg[0], n = seg[1] - seg[0];
  const float* K = Kmat + b * 9;
  Cam cam{(double)K[0], (double)K[4], (double)K[2], (double)K[5], (double)K[1]};
  const double thr2 = (double)thr * (double)thr;
  float* pose = pose_out + b * 12;
  auto fail = [&]() {
    if (tid < 12) pose[tid] = (tid == 0 || tid == 5 || tid == 10) ? 1.f : 0.f;
    if (tid == 0) {
      n_inl_out[b] = 0;
      status_out[b] = 0;
    }
    for (int i = tid; i < n; i += kPnpThreads) inl_mask[lo + i] = 0;
  };
  if (n < 4) {
    fail();
    return;
  }
  auto load_pt = [&](int i, double* p, double& u, double& v) {
    const float* q = pts3d + (long long)(lo + i) * 3;
    p[0] = (double)q[0] * scale;
    p[1] = (double)q[1] * scale;
    p[2] = (double)q[2] * scale;
    u = pts2d[(long long)(lo + i) * 2];
    v = pts2d[(long long)(lo + i) * 2 + 1];
  };

  // ---------------------------------------------------------------- 1+2: hypotheses and scoring
  int my_count = -1;
  unsigned my_h = 0xffffffffu;
  Pose my_pose;
  for (int h = tid; h < n_hyp; h += kPnpThreads) {
    int idx[4];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int pick = -1;
      for (int attempt = 0; attempt < 8 && pick < 0; ++attempt) {
        const uint32_t r = pnp_hash(seed ^ pnp_hash((uint32_t)b * 0x9E3779B9u + (uint32_t)h) ^
                                    ((uint32_t)(j * 8 + attempt + 1) * 0x85EBCA6Bu));
        const int cand = (int)(((unsigned long long)r * (unsigned long long)n) >> 32);
        bool dup = false;
        for (int k = 0; k < j; ++k) dup |= idx[k] == cand;
        if (!dup) pick = cand;
      }
      if (pick < 0) ok = false;
      idx[j] = pick < 0 ? 0 : pick;
    }
    if (!ok) continue;
    double P[3][3], f[3][3], p4[3], u4, v4;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double u, v;
      load_pt(idx[j], P[j], u, v);
      const double yn = (v - cam.cy) / cam.fy;
      const double xn = (u - cam.cx - cam.skew * yn) / cam.fx;
      f[j][0] = xn;
      f[j][1] = yn;
      f[j][2] = 1.0;
      normalize3(f[j]);
    }
    load_pt(idx[3], p4, u4, v4);
    Pose cand;
    double cand_err = INFINITY;
    p3p_grunert(P, f, [&](const Pose& ps) {
      const double e = reproj_err2(ps, cam, p4, u4, v4);
      if (e < cand_err) {
        cand_err = e;
        cand = ps;
      }
    });
    if (!(cand_err < INFINITY)) continue;
    int count = 0;
    for (int i = 0; i < n; ++i) {
      double p[3], u, v;
      load_pt(i, p, u, v);
      count += reproj_err2(cand, cam, p, u, v) < thr2 ? 1 : 0;
    }
    if (count > my_count) {   // strided h is increasing: the first best stays (lowest id on ties)
      my_count = count;
      my_h = (unsigned)h;
      my_pose = cand;
    }
  }
  // block argmax of (count, lowest h)
  unsigned long long key = my_count < 0 ? 0ull
                                        : (((unsigned long long)(unsigned)my_count + 1ull) << 32) |
                                              (unsigned long long)(0xffffffffu - my_h);
  unsigned long long wkey = key;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, wkey, o);
    wkey = other > wkey ? other : wkey;
  }
  if ((tid & 31) == 0) best_key[tid >> 5] = wkey;
  __syncthreads();
  unsigned long long bkey = 0;
#pragma unroll
  for (int w = 0; w < kPnpThreads / 32; ++w) bkey = best_key[w] > bkey ? best_key[w] : bkey;
  if (bkey == 0ull) {   // no admissible hypothesis at all
    fail();
    return;
  }
  if (key == bkey) best_pose = my_pose;   // keys are unique (h is)
  __syncthreads();
  Pose cur = best_pose;

  // ---------------------------------------------------------------- 3: refinement on the inliers
  int n_inl = 0;
  for (int round = 0; round <= refine_rounds; ++round) {
    // inlier set under the current pose
    int cnt = 0;
    for (int i = tid; i < n; i += kPnpThreads) {
      double p[3], u, v;
      load_pt(i, p, u, v);
      const unsigned char in = reproj_err2(cur, cam, p, u, v) < thr2 ? 1 : 0;
      inl_mask[lo + i] = in;
      cnt += in;
    }
    n_inl = (int)(block_sum((double)cnt, red) + 0.5);
    if (round == refine_rounds || n_inl < 4) break;
    __syncthreads();   // inl_mask written by other threads is read below
    for (int it = 0; it < 10; ++it) {
      double acc[27];
#pragma unroll
      for (int k = 0; k < 27; ++k) acc[k] = 0.0;
      for (int i = tid; i < n; i += kPnpThreads) {
        if (!inl_mask[lo + i]) continue;
        double p[3], u, v;
        load_pt(i, p, u, v);
        const double Y0 = cur.R[0] * p[0] + cur.R[1] * p[1] + cur.R[2] * p[2];
        const double Y1 = cur.R[3] * p[0] + cur.R[4] * p[1] + cur.R[5] * p[2];
        const double Y2 = cur.R[6] * p[0] + cur.R[7] * p[1] + cur.R[8] * p[2];
        const double x = Y0 + cur.t[0], y = Y1 + cur.t[1], z = Y2 + cur.t[2];
        if (!(z > 1e-9)) continue;
        const double iz = 1.0 / z, xn = x * iz, yn = y * iz;
        const double ru = cam.fx * xn + cam.skew * yn + cam.cx - u;
        const double rv = cam.fy * yn + cam.cy - v;
        // d(u)/dXc, d(v)/dXc
        const double gu[3] = {cam.fx * iz, cam.skew * iz, -(cam.fx * xn + cam.skew * yn) * iz};
        const double gv[3] = {0.0, cam.fy * iz, -cam.fy * yn * iz};
        // Xc = exp(w) Y + t  ->  dXc/dw = -[Y]x ; J row = (g x ... ) : g^T (-[Y]x) = (Y x g)^T
        const double Yv[3] = {Y0, Y1, Y2};
        double ju[6], jv[6];
        cross3(Yv, gu, ju);
        cross3(Yv, gv, jv);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          ju[3 + k] = gu[k];
          jv[3 + k] = gv[k];
        }
        int o = 0;
#pragma unroll
        for (int r_ = 0; r_ < 6; ++r_) {
#pragma unroll
          for (int c_ = r_; c_ < 6; ++c_) acc[o++] += ju[r_] * ju[c_] + jv[r_] * jv[c_];
        }
#pragma unroll
        for (int r_ = 0; r_ < 6; ++r_) acc[21 + r_] += ju[r_] * ru + jv[r_] * rv;
      }
      for (int k = 0; k < 27; ++k) {
        const double s = block_sum(acc[k], red);
        if (tid == 0) Hs[k] = s;
      }
      __syncthreads();
      if (tid == 0) {
        // solve (H + lambda diag H) d = -g by Cholesky
        double H[6][6], g[6], L[6][6], d[6];
        int o = 0;
        for (int r_ = 0; r_ < 6; ++r_)
          for (int c_ = r_; c_ < 6; ++c_) {
            H[r_][c_] = Hs[o];
            H[c_][r_] = Hs[o];
            ++o;
          }
        for (int r_ = 0; r_ < 6; ++r_) {
          g[r_] = Hs[21 + r_];
          H[r_][r_] *= 1.0 + 1e-9;
        }
        bool okc = true;
        for (int r_ = 0; r_ < 6 && okc; ++r_)
          for (int c_ = 0; c_ <= r_; ++c_) {
            double s = H[r_][c_];
            for (int k = 0; k < c_; ++k) s -= L[r_][k] * L[c_][k];
            if (r_ == c_) {
              if (!(s > 1e-300)) {
                okc = false;
                break;
              }
              L[r_][r_] = sqrt(s);
            } else {
              L[r_][c_] = s / L[c_][c_];
            }
          }
        double step = 0.0;
        if (okc) {
          double yv[6];
          for (int r_ = 0; r_ < 6; ++r_) {
            double s = -g[r_];
            for (int k = 0; k < r_; ++k) s -= L[r_][k] * yv[k];
            yv[r_] = s / L[r_][r_];
          }
          for (int r_ = 5; r_ >= 0; --r_) {
            double s = yv[r_];
            for (int k = r_ + 1; k < 6; ++k) s -= L[k][r_] * d[k];
            d[r_] = s / L[r_][r_];
          }
          // R <- exp(w) R (Rodrigues), t <- t + dt
          const double th = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
          double E[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
          if (th > 1e-16) {
            const double kx = d[0] / th, ky = d[1] / th, kz = d[2] / th;
            const double sn = sin(th), cs = 1.0 - cos(th);
            const double Kx[9] = {0, -kz, ky, kz, 0, -kx, -ky, kx, 0};
            double K2[9];
            for (int r_ = 0; r_ < 3; ++r_)
              for (int c_ = 0; c_ < 3; ++c_)
                K2[r_ * 3 + c_] = Kx[r_ * 3] * Kx[c_] + Kx[r_ * 3 + 1] * Kx[3 + c_] + Kx[r_ * 3 + 2] * Kx[6 + c_];
            for (int k = 0; k < 9; ++k) E[k] += sn * Kx[k] + cs * K2[k];
          }
          double Rn[9];
          for (int r_ = 0; r_ < 3; ++r_)
            for (int c_ = 0; c_ < 3; ++c_)
              Rn[r_ * 3 + c_] = E[r_ * 3] * best_pose.R[c_] + E[r_ * 3 + 1] * best_pose.R[3 + c_] +
                                E[r_ * 3 + 2] * best_pose.R[6 + c_];
          for (int k = 0; k < 9; ++k) best_pose.R[k] = Rn[k];
          for (int k = 0; k < 3; ++k) best_pose.t[k] += d[3 + k];
          for (int k = 0; k < 6; ++k) step = fmax(step, fabs(d[k]));
        }
        flag = (!okc || step < 1e-12) ? 1 : 0;
      }
      __syncthreads();
      cur = best_pose;
      const int stop = flag;
      __syncthreads();
      if (stop) break;
    }
  }
  if (tid < 9) pose[(tid / 3) * 4 + tid % 3] = (float)cur.R[tid];
  if (tid < 3) pose[tid * 4 + 3] = (float)(cur.t[tid] / (double)scale);
  if (tid == 0) {
    n_inl_out[b] = n_inl;
    status_out[b] = n_inl >= 4 ? 1 : 0;
  }
}

}  // namespace opp

using namespace opp;

extern "C" int opp_pnp_ransac(const float* pts3d, const float* pts2d, const long long* m_bids, int m,
                              const float* intrinsics, int batch, float scale, float reproj_thr,
                              int hypotheses, unsigned seed, int refine_rounds, float* poses,
                              int* n_inliers, unsigned char* inlier_mask, int* status,
                              opp_stream_t stream) {
  OPP_REQUIRE(intrinsics && poses && n_inliers && status, "null pointer");
  OPP_REQUIRE(m == 0 || (pts3d && pts2d && m_bids && inlier_mask), "null match lists");
  OPP_REQUIRE(batch > 0 && hypotheses > 0 && scale > 0.f && reproj_thr > 0.f && refine_rounds >= 0,
              "bad pnp arguments");
  OPP_CHECK_CUDA(opp::launch_pdl(pnp_ransac_kernel, dim3(batch), dim3(kPnpThreads), 0, (cudaStream_t)stream, 
      pts3d, pts2d, m_bids, m, intrinsics, scale, reproj_thr, hypotheses, seed, refine_rounds, poses,
      n_inliers, inlier_mask, status));
  OPP_CHECK_CUDA(cudaGetLastError());
  return OPP_OK;
}
