// dev_math.cuh — scalar fp64 helpers used inside the kernels: 3-vectors, row-major 3x3, spatial
// cross products, compact world-frame inertias, SO(3)/SE(3) exp/log and their right Jacobians.
// Stands in for the pinocchio spatial / explog headers the reference reaches through
// MultibodyPhaseSpace (reference src/kinodynamics.cpp:46) and pinocchio::difference
// (reference src/robot-handler.cpp:90).
#pragma once
#include "par.cuh"

namespace smpc
{

  SMPC_DEV void v3cross(const double * a, const double * b, double * o)
  {
    const double x = a[1] * b[2] - a[2] * b[1];
    const double y = a[2] * b[0] - a[0] * b[2];
    const double z = a[0] * b[1] - a[1] * b[0];
    o[0] = x;
    o[1] = y;
    o[2] = z;
  }
  SMPC_DEV double v3dot(const double * a, const double * b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
  // o = R v ; o = R^T v  (R row-major)
  SMPC_DEV void m3v(const double * R, const double * v, double * o)
  {
    const double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    const double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    const double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    o[0] = x;
    o[1] = y;
    o[2] = z;
  }
  SMPC_DEV void m3tv(const double * R, const double * v, double * o)
  {
    const double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    const double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    const double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
    o[0] = x;
    o[1] = y;
    o[2] = z;
  }
  SMPC_DEV void m3m(const double * A, const double * B, double * o)
  {
    double t[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++)
        t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
#pragma unroll
    for (int i = 0; i < 9; i++)
      o[i] = t[i];
  }
  SMPC_DEV void m3tm(const double * A, const double * B, double * o) // A^T B
  {
    double t[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++)
        t[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
#pragma unroll
    for (int i = 0; i < 9; i++)
      o[i] = t[i];
  }
  SMPC_DEV void m3mt(const double * A, const double * B, double * o) // A B^T
  {
    double t[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++)
        t[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
#pragma unroll
    for (int i = 0; i < 9; i++)
      o[i] = t[i];
  }
  SMPC_DEV void skew(const double * w, double * S)
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

  // spatial vectors [lin; ang]
  SMPC_DEV void mxm(const double * a, const double * b, double * o) // motion x motion
  {
    double t1[3], t2[3], t3[3];
    v3cross(a + 3, b, t1);
    v3cross(a, b + 3, t2);
    v3cross(a + 3, b + 3, t3);
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
      o[i] = t1[i] + t2[i];
      o[3 + i] = t3[i];
    }
  }
  SMPC_DEV void mxf(const double * a, const double * f, double * o) // motion x* force
  {
    double t1[3], t2[3], t3[3];
    v3cross(a + 3, f, t1);
    v3cross(a + 3, f + 3, t2);
    v3cross(a, f, t3);
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
      o[i] = t1[i];
      o[3 + i] = t2[i] + t3[i];
    }
  }

  // Compact spatial inertia about the WORLD origin: Y = {m, mc[3], Io[6] = xx xy xz yy yz zz}.
  //   Y (vl, va) = ( m vl + va x mc ,  Io va + mc x vl )
  SMPC_DEV void sym3v(const double * S, const double * v, double * o)
  {
    const double x = S[0] * v[0] + S[1] * v[1] + S[2] * v[2];
    const double y = S[1] * v[0] + S[3] * v[1] + S[4] * v[2];
    const double z = S[2] * v[0] + S[4] * v[1] + S[5] * v[2];
    o[0] = x;
    o[1] = y;
    o[2] = z;
  }
  SMPC_DEV void inertia_apply(const double * Y /*10*/, const double * v, double * o)
  {
    double t1[3], t2[3], t3[3];
    v3cross(v + 3, Y + 1, t1);
    sym3v(Y + 4, v + 3, t2);
    v3cross(Y + 1, v, t3);
    const double m = Y[0];
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
      o[i] = m * v[i] + t1[i];
      o[3 + i] = t2[i] + t3[i];
    }
  }
  // Time derivative of a composite inertia: dY = {dmc[3], dIo[6]} (mass rate is zero).
  //   dY (wl, wa) = ( wa x dmc ,  dIo wa + dmc x wl )
  SMPC_DEV void dinertia_apply(const double * dmc, const double * dIo, const double * w, double * o)
  {
    double t1[3], t2[3], t3[3];
    v3cross(w + 3, dmc, t1);
    sym3v(dIo, w + 3, t2);
    v3cross(dmc, w, t3);
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
      o[i] = t1[i];
      o[3 + i] = t2[i] + t3[i];
    }
  }

  // ---- series-safe trigonometric coefficients (same closed forms / series as the oracle) ----
  struct TrigCoef
  {
    double A, B, C, D, E, F;
  };
  // closed forms for rotations beyond 0.5 rad: rare on this path (steps dt v, state errors), and sqrt / sin / cos with their
  // range-reduction slow paths are ~450 instructions per inlined copy - six copies were 13 % of the knot kernel's code, which
  // already runs short of instruction cache (ncu: "no instruction" stalls 1.3 per issue) - so ONE out-of-line copy
  static SMPC_DEV_NOINLINE void trig_coef_large(double t2, TrigCoef & k)
  {
    const double t = sqrt(t2);
    const double s = sin(t), c = cos(t);
    k.A = s / t;
    k.B = (1 - c) / t2;
    k.C = (t - s) / (t2 * t);
    k.D = 1.0 / t2 - (1 + c) / (2 * t * s);
    k.E = (t2 + 2 * c - 2) / (2 * t2 * t2);
    k.F = (2 * t - 3 * s + t * c) / (2 * t2 * t2 * t);
  }
  SMPC_DEV TrigCoef trig_coef(double t2)
  {
    TrigCoef k;
    if (t2 < 0.25)
    {
      // Ten-term Taylor sums in Horner form with constant coefficients, the same expressions as oracle/orc_math.hpp.
      // (Both used to divide every term by its factorial: 50 IEEE divisions run by one lane while the warp waits,
      // about 11 percent of the knot kernel's issue slots in profiles/r01e.)
      k.A = 1.0 / 1.0 + t2 * (-1.0 / 6.0 + t2 * (1.0 / 120.0 + t2 * (-1.0 / 5040.0 + t2 * (1.0 / 362880.0 + t2 * (-1.0 / 39916800.0 + t2 * (1.0 / 6227020800.0 + t2 * (-1.0 / 1307674368000.0 + t2 * (1.0 / 355687428096000.0 + t2 * (-1.0 / 121645100408832000.0)))))))));
      k.B = 1.0 / 2.0 + t2 * (-1.0 / 24.0 + t2 * (1.0 / 720.0 + t2 * (-1.0 / 40320.0 + t2 * (1.0 / 3628800.0 + t2 * (-1.0 / 479001600.0 + t2 * (1.0 / 87178291200.0 + t2 * (-1.0 / 20922789888000.0 + t2 * (1.0 / 6402373705728000.0 + t2 * (-1.0 / 2432902008176640000.0)))))))));
      k.C = 1.0 / 6.0 + t2 * (-1.0 / 120.0 + t2 * (1.0 / 5040.0 + t2 * (-1.0 / 362880.0 + t2 * (1.0 / 39916800.0 + t2 * (-1.0 / 6227020800.0 + t2 * (1.0 / 1307674368000.0 + t2 * (-1.0 / 355687428096000.0 + t2 * (1.0 / 121645100408832000.0 + t2 * (-1.0 / 51090942171709440000.0)))))))));
      k.E = 1.0 / 24.0 + t2 * (-1.0 / 720.0 + t2 * (1.0 / 40320.0 + t2 * (-1.0 / 3628800.0 + t2 * (1.0 / 479001600.0 + t2 * (-1.0 / 87178291200.0 + t2 * (1.0 / 20922789888000.0 + t2 * (-1.0 / 6402373705728000.0 + t2 * (1.0 / 2432902008176640000.0 + t2 * (-1.0 / 1124000727777607680000.0)))))))));
      k.F = 1.0 / 120.0 + t2 * (-2.0 / 5040.0 + t2 * (3.0 / 362880.0 + t2 * (-4.0 / 39916800.0 + t2 * (5.0 / 6227020800.0 + t2 * (-6.0 / 1307674368000.0 + t2 * (7.0 / 355687428096000.0 + t2 * (-8.0 / 121645100408832000.0 + t2 * (9.0 / 51090942171709440000.0 + t2 * (-10.0 / 25852016738884976640000.0)))))))));
      k.D = 1.0 / 12
            + t2 * (1.0 / 720 + t2 * (1.0 / 30240 + t2 * (1.0 / 1209600 + t2 * (1.0 / 47900160 + t2 * (691.0 / 1307674368000.0 + t2 * (1.0 / 74724249600.0 + t2 * (3617.0 / 10670622842880000.0)))))));
    }
    else
      trig_coef_large(t2, k);
    return k;
  }
  SMPC_DEV void eye3(double * R)
  {
#pragma unroll
    for (int i = 0; i < 9; i++)
      R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }
  // w = log3(R), theta < pi
  SMPC_DEV void log3(const double * R, double * w)
  {
    const double tr = R[0] + R[4] + R[8];
    const double ax[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    const double s2 = 0.5 * sqrt(v3dot(ax, ax));
    const double c = 0.5 * (tr - 1.0);
    const double t = atan2(s2, c);
    const double f = (t < 1e-8) ? 0.5 * (1 + t * t / 6) : 0.5 * t / s2;
#pragma unroll
    for (int i = 0; i < 3; i++)
      w[i] = f * ax[i];
  }
  // Jr = I - B W + C W^2 (right Jacobian of exp3) ; Jr^-1 = I + W/2 + D W^2
  SMPC_DEV void Jexp3(const double * w, double * J)
  {
    const TrigCoef k = trig_coef(v3dot(w, w));
    double W[9], W2[9];
    skew(w, W);
    m3m(W, W, W2);
#pragma unroll
    for (int i = 0; i < 9; i++)
      J[i] = ((i % 4 == 0) ? 1.0 : 0.0) - k.B * W[i] + k.C * W2[i];
  }
  SMPC_DEV void Jlog3(const double * w, double * J)
  {
    const TrigCoef k = trig_coef(v3dot(w, w));
    double W[9], W2[9];
    skew(w, W);
    m3m(W, W, W2);
#pragma unroll
    for (int i = 0; i < 9; i++)
      J[i] = ((i % 4 == 0) ? 1.0 : 0.0) + 0.5 * W[i] + k.D * W2[i];
  }
  // (R, p) = exp6(nu)
  SMPC_DEV void exp6(const double * nu, double * R, double * p)
  {
    const double * w = nu + 3;
    const TrigCoef k = trig_coef(v3dot(w, w));
    double W[9], W2[9], V[9];
    skew(w, W);
    m3m(W, W, W2);
#pragma unroll
    for (int i = 0; i < 9; i++)
    {
      const double I = (i % 4 == 0) ? 1.0 : 0.0;
      R[i] = I + k.A * W[i] + k.B * W2[i];
      V[i] = I + k.B * W[i] + k.C * W2[i];
    }
    m3v(V, nu, p);
  }
  SMPC_DEV void log6(const double * R, const double * p, double * nu)
  {
    double w[3];
    log3(R, w);
    const TrigCoef k = trig_coef(v3dot(w, w));
    double W[9], W2[9], Vi[9];
    skew(w, W);
    m3m(W, W, W2);
#pragma unroll
    for (int i = 0; i < 9; i++)
      Vi[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * W[i] + k.D * W2[i];
    m3v(Vi, p, nu);
    nu[3] = w[0];
    nu[4] = w[1];
    nu[5] = w[2];
  }
  // coupling block of the SE(3) left Jacobian (Barfoot eq. 7.86)
  SMPC_DEV void se3_Q(const double * v, const double * w, double * Q)
  {
    const TrigCoef k = trig_coef(v3dot(w, w));
    double W[9], V[9], WV[9], VW[9], WVW[9], WWV[9], VWW[9], WVWW[9], WWVW[9];
    skew(w, W);
    skew(v, V);
    m3m(W, V, WV);
    m3m(V, W, VW);
    m3m(WV, W, WVW);
    m3m(W, WV, WWV);
    m3m(VW, W, VWW);
    m3m(WVW, W, WVWW);
    m3m(W, WVW, WWVW);
#pragma unroll
    for (int i = 0; i < 9; i++)
      Q[i] = 0.5 * V[i] + k.C * (WV[i] + VW[i] + WVW[i]) + k.E * (WWV[i] + VWW[i] - 3 * WVW[i]) + k.F * (WVWW[i] + WWVW[i]);
  }
  // right Jacobian of exp6: [[Jr, Q(-nu)],[0, Jr]] and its inverse [[Ji, -Ji Q Ji],[0, Ji]] (6x6 row-major)
  SMPC_DEV void Jexp6(const double * nu, double * J)
  {
    const double mv[3] = {-nu[0], -nu[1], -nu[2]}, mw[3] = {-nu[3], -nu[4], -nu[5]};
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
  SMPC_DEV void Jlog6(const double * nu, double * J)
  {
    const double mv[3] = {-nu[0], -nu[1], -nu[2]}, mw[3] = {-nu[3], -nu[4], -nu[5]};
    double Ji[9], Q[9], T[9], B[9];
    Jlog3(nu + 3, Ji);
    se3_Q(mv, mw, Q);
    m3m(Ji, Q, T);
    m3m(T, Ji, B);
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
  // 6x6 action matrix of (R,p)^-1 : [[R^T, -R^T [p]x],[0, R^T]]
  SMPC_DEV void actinv6(const double * R, const double * p, double * X)
  {
    double P[9], RtP[9];
    skew(p, P);
    m3tm(R, P, RtP);
    for (int i = 0; i < 36; i++)
      X[i] = 0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
      {
        X[6 * i + j] = R[3 * j + i];
        X[6 * (i + 3) + (j + 3)] = R[3 * j + i];
        X[6 * i + (j + 3)] = -RtP[3 * i + j];
      }
  }
  SMPC_DEV void quat_to_R(const double * q, double * R)
  {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    const double n = 1.0 / sqrt(x * x + y * y + z * z + w * w);
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
  SMPC_DEV void R_to_quat(const double * R, double * q)
  {
    const double tr = R[0] + R[4] + R[8];
    double x, y, z, w;
    if (tr > 0)
    {
      const double s = sqrt(tr + 1.0) * 2;
      w = 0.25 * s;
      x = (R[7] - R[5]) / s;
      y = (R[2] - R[6]) / s;
      z = (R[3] - R[1]) / s;
    }
    else if (R[0] > R[4] && R[0] > R[8])
    {
      const double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
      w = (R[7] - R[5]) / s;
      x = 0.25 * s;
      y = (R[1] + R[3]) / s;
      z = (R[2] + R[6]) / s;
    }
    else if (R[4] > R[8])
    {
      const double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
      w = (R[2] - R[6]) / s;
      x = (R[1] + R[3]) / s;
      y = 0.25 * s;
      z = (R[5] + R[7]) / s;
    }
    else
    {
      const double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
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

  // Free-flyer + joints configuration operations (one thread):
  //   qo = q (+) dq       (dq in R^nv, free-flyer block in the local frame)
  SMPC_DEV void integrate_q(int nv, const double * q, const double * dq, double * qo)
  {
    double R[9], E[9], ep[3], Rn[9], t[3];
    quat_to_R(q + 3, R);
    exp6(dq, E, ep);
    m3m(R, E, Rn);
    m3v(R, ep, t);
    qo[0] = q[0] + t[0];
    qo[1] = q[1] + t[1];
    qo[2] = q[2] + t[2];
    double qq[4];
    R_to_quat(Rn, qq);
    const double d = q[3] * qq[0] + q[4] * qq[1] + q[5] * qq[2] + q[6] * qq[3];
    const double sg = (d < 0) ? -1.0 : 1.0;
    for (int i = 0; i < 4; i++)
      qo[3 + i] = sg * qq[i];
    for (int i = 6; i < nv; i++)
      qo[1 + i] = q[1 + i] + dq[i];
  }
  //   d = q1 (-) q0
  SMPC_DEV void difference_q(int nv, const double * q0, const double * q1, double * d)
  {
    double R0[9], R1[9], Rd[9], dp[3], pl[3];
    quat_to_R(q0 + 3, R0);
    quat_to_R(q1 + 3, R1);
    m3tm(R0, R1, Rd);
    dp[0] = q1[0] - q0[0];
    dp[1] = q1[1] - q0[1];
    dp[2] = q1[2] - q0[2];
    m3tv(R0, dp, pl);
    log6(Rd, pl, d);
    for (int i = 6; i < nv; i++)
      d[i] = q1[1 + i] - q0[1 + i];
  }

} // namespace smpc
