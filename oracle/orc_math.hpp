// ORACLE — test infrastructure only. Never linked into, imported by, or called from the product
// path (simple-mpc_b200/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it.
//
// PARITY UNPINNED: Aligator 0.16 / Pinocchio 3.8 / ndcurves (the libraries that hold the arithmetic
// of reference src/mpc.cpp:212) are not in /root/reference and cannot be built here; the reference's
// own tests hold no floating-point golden vector for this path. This restatement is
// pinned by finite-difference checks and KKT-residual checks instead (tests/test_oracle_*.py).
//
// orc_math.hpp — small dense fp64 helpers, SO(3)/SE(3) exp/log and their Jacobians.
// Follows the published closed forms Pinocchio implements (pinocchio/spatial/explog.hpp, [UPSTREAM]),
// used by the reference through MultibodyPhaseSpace (reference src/kinodynamics.cpp:46) and
// pinocchio::difference (reference src/robot-handler.cpp:90).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include <cassert>

namespace orc
{

  // Row-major dynamic matrix.
  struct Mat
  {
    int r = 0, c = 0;
    std::vector<double> d;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), d((size_t)r_ * c_, 0.0) {}
    void resize(int r_, int c_)
    {
      r = r_;
      c = c_;
      d.assign((size_t)r_ * c_, 0.0);
    }
    void zero() { std::fill(d.begin(), d.end(), 0.0); }
    double & operator()(int i, int j) { return d[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return d[(size_t)i * c + j]; }
    double * row(int i) { return &d[(size_t)i * c]; }
    const double * row(int i) const { return &d[(size_t)i * c]; }
  };
  typedef std::vector<double> Vec;

  // ---------- 3-vectors / 3x3 (row-major double[9]) ----------
  inline void cross3(const double * a, const double * b, double * o)
  {
    double x = a[1] * b[2] - a[2] * b[1];
    double y = a[2] * b[0] - a[0] * b[2];
    double z = a[0] * b[1] - a[1] * b[0];
    o[0] = x;
    o[1] = y;
    o[2] = z;
  }
  inline double dot3(const double * a, const double * b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
  inline void matvec3(const double * R, const double * v, double * o)
  {
    double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    o[0] = x;
    o[1] = y;
    o[2] = z;
  }
  inline void matTvec3(const double * R, const double * v, double * o)
  {
    double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
    o[0] = x;
    o[1] = y;
    o[2] = z;
  }
  inline void matmul3(const double * A, const double * B, double * o)
  {
    double t[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(o, t, sizeof(t));
  }
  inline void matTmul3(const double * A, const double * B, double * o) // A^T B
  {
    double t[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        t[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
    memcpy(o, t, sizeof(t));
  }
  inline void skew3(const double * w, double * S)
  {
    S[0] = 0;
    S[1] = -w[2];
    S[2] = w[1];
    S[3] = w[2];
    S[4] = 0;
    S[5] = -w[0];
    S[6] = -w[1];
    S[7] = w[0];
    S[8] = 0;
  }
  inline void eye3(double * R)
  {
    for (int i = 0; i < 9; i++)
      R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }

  struct SE3
  {
    double R[9];
    double p[3];
    SE3()
    {
      eye3(R);
      p[0] = p[1] = p[2] = 0;
    }
  };
  inline SE3 se3_mul(const SE3 & a, const SE3 & b)
  {
    SE3 o;
    matmul3(a.R, b.R, o.R);
    double t[3];
    matvec3(a.R, b.p, t);
    for (int i = 0; i < 3; i++)
      o.p[i] = a.p[i] + t[i];
    return o;
  }
  inline SE3 se3_inv(const SE3 & a)
  {
    SE3 o;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        o.R[3 * i + j] = a.R[3 * j + i];
    double t[3];
    matTvec3(a.R, a.p, t);
    for (int i = 0; i < 3; i++)
      o.p[i] = -t[i];
    return o;
  }

  // ---------- spatial vectors [lin; ang] ----------
  // motion x motion
  inline void mxm(const double * a, const double * b, double * o)
  {
    double t1[3], t2[3], t3[3];
    cross3(a + 3, b, t1);
    cross3(a, b + 3, t2);
    cross3(a + 3, b + 3, t3);
    for (int i = 0; i < 3; i++)
    {
      o[i] = t1[i] + t2[i];
      o[3 + i] = t3[i];
    }
  }
  // motion x* force
  inline void mxf(const double * a, const double * f, double * o)
  {
    double t1[3], t2[3], t3[3];
    cross3(a + 3, f, t1);
    cross3(a + 3, f + 3, t2);
    cross3(a, f, t3);
    for (int i = 0; i < 3; i++)
    {
      o[i] = t1[i];
      o[3 + i] = t2[i] + t3[i];
    }
  }
  // Adjoint action of M on a motion expressed in M's frame: result in the parent frame.
  inline void se3_act_motion(const SE3 & M, const double * m, double * o)
  {
    double Rv[3], Rw[3], pxRw[3];
    matvec3(M.R, m, Rv);
    matvec3(M.R, m + 3, Rw);
    cross3(M.p, Rw, pxRw);
    for (int i = 0; i < 3; i++)
    {
      o[i] = Rv[i] + pxRw[i];
      o[3 + i] = Rw[i];
    }
  }

  // 6x6 row-major helpers
  inline void mat6vec(const double * A, const double * v, double * o)
  {
    double t[6];
    for (int i = 0; i < 6; i++)
    {
      double s = 0;
      for (int j = 0; j < 6; j++)
        s += A[6 * i + j] * v[j];
      t[i] = s;
    }
    memcpy(o, t, sizeof(t));
  }

  // ---------- series-safe trigonometric coefficients ----------
  // A = sin(t)/t, B = (1-cos t)/t^2, C = (t - sin t)/t^3, and for the inverse
  // D = 1/t^2 - (1+cos t)/(2 t sin t); E.. used by the SE(3) Q block. Series below t^2 < 0.25.
  struct TrigCoef
  {
    double A, B, C, D;
    double E; // (t^2 + 2 cos t - 2) / (2 t^4)      ->  1/24 - t^2/720 ...
    double F; // (2 t - 3 sin t + t cos t)/(2 t^5)  ->  1/120 - ... (Barfoot's 4th coefficient)
  };
  inline TrigCoef trig_coef(double t2)
  {
    TrigCoef k;
    if (t2 < 0.25)
    {
      // alternating Taylor series in t2, ten terms (< 1e-17 truncation at t2 = 0.25), evaluated in Horner form with
      // constant coefficients (the device code evaluates exactly the same expressions):
      // A = sum (-1)^n t2^n/(2n+1)!   B = sum (-1)^n t2^n/(2n+2)!   C = sum (-1)^n t2^n/(2n+3)!
      // E = sum (-1)^n t2^n/(2n+4)!   F = sum (-1)^n (n+1) t2^n/(2n+5)!
      //   F: (2t - 3 sin t + t cos t)/(2 t^5): the coefficient of t^(2n+5) in (-3 sin t + t cos t) is
      //   (-1)^n (2n+2)/(2n+5)!, halved
      k.A = 1.0 / 1.0 + t2 * (-1.0 / 6.0 + t2 * (1.0 / 120.0 + t2 * (-1.0 / 5040.0 + t2 * (1.0 / 362880.0 + t2 * (-1.0 / 39916800.0 + t2 * (1.0 / 6227020800.0 + t2 * (-1.0 / 1307674368000.0 + t2 * (1.0 / 355687428096000.0 + t2 * (-1.0 / 121645100408832000.0)))))))));
      k.B = 1.0 / 2.0 + t2 * (-1.0 / 24.0 + t2 * (1.0 / 720.0 + t2 * (-1.0 / 40320.0 + t2 * (1.0 / 3628800.0 + t2 * (-1.0 / 479001600.0 + t2 * (1.0 / 87178291200.0 + t2 * (-1.0 / 20922789888000.0 + t2 * (1.0 / 6402373705728000.0 + t2 * (-1.0 / 2432902008176640000.0)))))))));
      k.C = 1.0 / 6.0 + t2 * (-1.0 / 120.0 + t2 * (1.0 / 5040.0 + t2 * (-1.0 / 362880.0 + t2 * (1.0 / 39916800.0 + t2 * (-1.0 / 6227020800.0 + t2 * (1.0 / 1307674368000.0 + t2 * (-1.0 / 355687428096000.0 + t2 * (1.0 / 121645100408832000.0 + t2 * (-1.0 / 51090942171709440000.0)))))))));
      k.E = 1.0 / 24.0 + t2 * (-1.0 / 720.0 + t2 * (1.0 / 40320.0 + t2 * (-1.0 / 3628800.0 + t2 * (1.0 / 479001600.0 + t2 * (-1.0 / 87178291200.0 + t2 * (1.0 / 20922789888000.0 + t2 * (-1.0 / 6402373705728000.0 + t2 * (1.0 / 2432902008176640000.0 + t2 * (-1.0 / 1124000727777607680000.0)))))))));
      k.F = 1.0 / 120.0 + t2 * (-2.0 / 5040.0 + t2 * (3.0 / 362880.0 + t2 * (-4.0 / 39916800.0 + t2 * (5.0 / 6227020800.0 + t2 * (-6.0 / 1307674368000.0 + t2 * (7.0 / 355687428096000.0 + t2 * (-8.0 / 121645100408832000.0 + t2 * (9.0 / 51090942171709440000.0 + t2 * (-10.0 / 25852016738884976640000.0)))))))));
      // D = 1/t2 - (1+cos)/(2 t sin) = (1 - A/(2B)) / t2 ; series: 1/12 + t2/720 + t2^2/30240 + t2^3/1209600 + ...
      k.D = 1.0 / 12 + t2 * (1.0 / 720 + t2 * (1.0 / 30240 + t2 * (1.0 / 1209600 + t2 * (1.0 / 47900160 + t2 * (691.0 / 1307674368000.0 + t2 * (1.0 / 74724249600.0 + t2 * (3617.0 / 10670622842880000.0)))))));
    }
    else
    {
      double t = std::sqrt(t2);
      double s = std::sin(t), c = std::cos(t);
      k.A = s / t;
      k.B = (1 - c) / t2;
      k.C = (t - s) / (t2 * t);
      k.D = 1.0 / t2 - (1 + c) / (2 * t * s);
      k.E = (t2 + 2 * c - 2) / (2 * t2 * t2);
      k.F = (2 * t - 3 * s + t * c) / (2 * t2 * t2 * t);
    }
    return k;
  }

  // R = exp3(w)
  inline void exp3(const double * w, double * R)
  {
    double t2 = dot3(w, w);
    TrigCoef k = trig_coef(t2);
    double W[9], W2[9];
    skew3(w, W);
    matmul3(W, W, W2);
    for (int i = 0; i < 9; i++)
      R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + k.A * W[i] + k.B * W2[i];
  }
  // w = log3(R) (theta < pi assumed: MPC states stay far from the cut locus)
  inline void log3(const double * R, double * w)
  {
    double tr = R[0] + R[4] + R[8];
    double ax[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]}; // 2 sin(t) * axis
    double s2 = 0.5 * std::sqrt(dot3(ax, ax));              // sin t
    double c = 0.5 * (tr - 1.0);
    double t = std::atan2(s2, c);
    double f; // t / (2 sin t); no cancellation, only 0/0 at t = 0
    if (t < 1e-8)
      f = 0.5 * (1 + t * t / 6);
    else
      f = 0.5 * t / s2;
    for (int i = 0; i < 3; i++)
      w[i] = f * ax[i];
  }
  // Right Jacobian of exp3 and its inverse (Pinocchio Jexp3 / Jlog3):
  //   Jr = I - B W + C W^2 ;  Jr^-1 = I + W/2 + D W^2
  inline void Jexp3(const double * w, double * J)
  {
    TrigCoef k = trig_coef(dot3(w, w));
    double W[9], W2[9];
    skew3(w, W);
    matmul3(W, W, W2);
    for (int i = 0; i < 9; i++)
      J[i] = ((i % 4 == 0) ? 1.0 : 0.0) - k.B * W[i] + k.C * W2[i];
  }
  inline void Jlog3(const double * w, double * J)
  {
    TrigCoef k = trig_coef(dot3(w, w));
    double W[9], W2[9];
    skew3(w, W);
    matmul3(W, W, W2);
    for (int i = 0; i < 9; i++)
      J[i] = ((i % 4 == 0) ? 1.0 : 0.0) + 0.5 * W[i] + k.D * W2[i];
  }

  // M = exp6(nu), nu = [v; w]:  R = exp3(w), p = Jl(w) v with Jl = I + B W + C W^2
  inline SE3 exp6(const double * nu)
  {
    SE3 M;
    const double * v = nu;
    const double * w = nu + 3;
    TrigCoef k = trig_coef(dot3(w, w));
    double W[9], W2[9];
    skew3(w, W);
    matmul3(W, W, W2);
    double V[9];
    for (int i = 0; i < 9; i++)
    {
      double I = (i % 4 == 0) ? 1.0 : 0.0;
      M.R[i] = I + k.A * W[i] + k.B * W2[i];
      V[i] = I + k.B * W[i] + k.C * W2[i];
    }
    matvec3(V, v, M.p);
    return M;
  }
  // nu = log6(M)
  inline void log6(const SE3 & M, double * nu)
  {
    double w[3];
    log3(M.R, w);
    TrigCoef k = trig_coef(dot3(w, w));
    double W[9], W2[9], Vi[9];
    skew3(w, W);
    matmul3(W, W, W2);
    for (int i = 0; i < 9; i++)
      Vi[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * W[i] + k.D * W2[i];
    matvec3(Vi, M.p, nu);
    nu[3] = w[0];
    nu[4] = w[1];
    nu[5] = w[2];
  }
  // Left-Jacobian coupling block Q(v,w) of SE(3) (Barfoot, State Estimation for Robotics, eq. 7.86)
  inline void se3_Q(const double * v, const double * w, double * Q)
  {
    TrigCoef k = trig_coef(dot3(w, w));
    double W[9], V[9], WV[9], VW[9], WVW[9], WWV[9], VWW[9], WVWW[9], WWVW[9];
    skew3(w, W);
    skew3(v, V);
    matmul3(W, V, WV);
    matmul3(V, W, VW);
    matmul3(WV, W, WVW);
    matmul3(W, WV, WWV);
    matmul3(VW, W, VWW);
    matmul3(WVW, W, WVWW);
    matmul3(W, WVW, WWVW);
    for (int i = 0; i < 9; i++)
      Q[i] = 0.5 * V[i] + k.C * (WV[i] + VW[i] + WVW[i]) + k.E * (WWV[i] + VWW[i] - 3 * WVW[i]) + k.F * (WVWW[i] + WWVW[i]);
  }
  // Right Jacobian of exp6 (Pinocchio Jexp6): d/d(nu) of M(nu) in the local tangent at M(nu).
  //   Jr(nu) = Jl(-nu) = [[Jl3(-w), Q(-v,-w)],[0, Jl3(-w)]],  Jl3(-w) = Jr3(w)
  inline void Jexp6(const double * nu, double * J /*6x6*/)
  {
    double mv[3] = {-nu[0], -nu[1], -nu[2]}, mw[3] = {-nu[3], -nu[4], -nu[5]};
    double Jr[9], Q[9];
    Jexp3(nu + 3, Jr);
    se3_Q(mv, mw, Q);
    for (int i = 0; i < 36; i++)
      J[i] = 0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
      {
        J[6 * i + j] = Jr[3 * i + j];
        J[6 * (i + 3) + (j + 3)] = Jr[3 * i + j];
        J[6 * i + (j + 3)] = Q[3 * i + j];
      }
  }
  // Inverse of the above (Pinocchio Jlog6): [[Jr^-1, -Jr^-1 Q Jr^-1],[0, Jr^-1]]
  inline void Jlog6(const double * nu, double * J /*6x6*/)
  {
    double mv[3] = {-nu[0], -nu[1], -nu[2]}, mw[3] = {-nu[3], -nu[4], -nu[5]};
    double Ji[9], Q[9], T[9], B[9];
    Jlog3(nu + 3, Ji);
    se3_Q(mv, mw, Q);
    matmul3(Ji, Q, T);
    matmul3(T, Ji, B);
    for (int i = 0; i < 36; i++)
      J[i] = 0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
      {
        J[6 * i + j] = Ji[3 * i + j];
        J[6 * (i + 3) + (j + 3)] = Ji[3 * i + j];
        J[6 * i + (j + 3)] = -B[3 * i + j];
      }
  }
  // 6x6 action matrix of M^-1 (motion from parent frame to M's frame): [[R^T, -R^T [p]x],[0, R^T]]
  inline void se3_actinv_matrix(const SE3 & M, double * X)
  {
    double P[9], RtP[9];
    skew3(M.p, P);
    matTmul3(M.R, P, RtP);
    for (int i = 0; i < 36; i++)
      X[i] = 0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
      {
        X[6 * i + j] = M.R[3 * j + i];
        X[6 * (i + 3) + (j + 3)] = M.R[3 * j + i];
        X[6 * i + (j + 3)] = -RtP[3 * i + j];
      }
  }

  // quaternion (x,y,z,w) <-> rotation
  inline void quat_to_R(const double * q, double * R)
  {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double n = 1.0 / std::sqrt(x * x + y * y + z * z + w * w);
    x *= n;
    y *= n;
    z *= n;
    w *= n;
    R[0] = 1 - 2 * (y * y + z * z);
    R[1] = 2 * (x * y - z * w);
    R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);
    R[4] = 1 - 2 * (x * x + z * z);
    R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);
    R[7] = 2 * (y * z + x * w);
    R[8] = 1 - 2 * (x * x + y * y);
  }
  inline void R_to_quat(const double * R, double * q)
  {
    double tr = R[0] + R[4] + R[8];
    double x, y, z, w;
    if (tr > 0)
    {
      double s = std::sqrt(tr + 1.0) * 2;
      w = 0.25 * s;
      x = (R[7] - R[5]) / s;
      y = (R[2] - R[6]) / s;
      z = (R[3] - R[1]) / s;
    }
    else if (R[0] > R[4] && R[0] > R[8])
    {
      double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
      w = (R[7] - R[5]) / s;
      x = 0.25 * s;
      y = (R[1] + R[3]) / s;
      z = (R[2] + R[6]) / s;
    }
    else if (R[4] > R[8])
    {
      double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
      w = (R[2] - R[6]) / s;
      x = (R[1] + R[3]) / s;
      y = 0.25 * s;
      z = (R[5] + R[7]) / s;
    }
    else
    {
      double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
      w = (R[3] - R[1]) / s;
      x = (R[2] + R[6]) / s;
      y = (R[5] + R[7]) / s;
      z = 0.25 * s;
    }
    q[0] = x;
    q[1] = y;
    q[2] = z;
    q[3] = w;
  }

} // namespace orc
