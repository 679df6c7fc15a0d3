// id_qp.cuh — one CTA solves one whole-body inverse-dynamics QP (the reference's KinodynamicsID::solve /
// CentroidalID, src/inverse-dynamics/kinodynamics-id.cpp:188-232; TSID acceleration/force formulation + dense QP).
//
// Everything lives in shared memory (sized at context creation from the robot and the settings, IdDims):
//   phase group 1  world-frame kinematics by tree level, body inertias, composite inertias (CRBA), bias forces (RNEA with
//                  a = 0): joint-space inertia M and nonlinear effects h
//   phase group 2  the tasks of the formulation, one <= 6-row block at a time (frame Jacobian in the LOCAL or the
//                  world-oriented frame, desired acceleration, drift): H += w J^T J, g -= w J^T b; contact blocks also
//                  write their columns of Jc^T T, their cone rows and (when motion is a hard constraint) equality rows
//   phase group 3  proximal method of multipliers with a semismooth-Newton inner loop on
//                    min 1/2 x'Hx + g'x  s.t.  A x = b,  l <= C x <= u        x = [a; f]
//                  (dense Newton matrix K = H + rho I + (A'A + C_act'C_act)/mu factorised by a one-barrier-per-column
//                  left-looking Cholesky), tau = M_a a + h_a - (Jc^T T)_a f
// Written in the SMPC_PAR phase discipline of par.cuh, so tests/host_emul compiles it for the CPU.
// Algorithm, constants and task definitions follow oracle/orc_id.hpp (which cites TSID); the comparison with the oracle is
// at solution level (both solve the same strictly convex QP to a relative KKT residual of 1e-10).
#pragma once
#include "dev_math.cuh"
#include "dev_types.h"
#include "../../include/smpc_id.h"

namespace smpc
{

  constexpr double ID_HESSIAN_REG = 1e-8;
  constexpr double ID_BIG = 1e10;
  constexpr double ID_DDQ_MAX = 1e10;
  constexpr double ID_QP_EPS = 1e-10;
  constexpr double ID_QP_RHO = 1e-10;
  constexpr double ID_QP_MU0 = 1e-2;
  constexpr double ID_QP_MU_MIN = 1e-4;
  constexpr double ID_QP_MU_WARM = 1e-3; // first penalty parameter of a warm-started solve
  constexpr double ID_QP_GRAD_REL = 1e-13;
  constexpr int ID_QP_MAX_OUTER = 30;
  constexpr int ID_QP_MAX_INNER = 30;
  constexpr int ID_QP_MAX_NEWTON = 200;
  constexpr int ID_RED_SLOTS = 10; // reduction slots (id_red_put / id_red_get)
  constexpr int ID_THREADS = 128; // CTA width the reductions are written for (the emulation runs with SMPC_EMUL_NT = 128)

  // per-instance target (what setTarget stores in the TSID tasks)
  struct IdTargetDev
  {
    double q_post[SMPC_MAX_NV], v_post[SMPC_MAX_NV], a_post[SMPC_MAX_NV];
    double base_R[9], base_p[3];
    double base_vel[6], base_acc[6];
    double f_ref[SMPC_MAX_FEET][6];
    double com_ref[3], com_vel[3];
    double foot_R[SMPC_MAX_FEET][9], foot_p[SMPC_MAX_FEET][3], foot_vel[SMPC_MAX_FEET][6];
    int active[SMPC_MAX_FEET];
    int tracked[SMPC_MAX_FEET];
  };

  // shared-memory plan (offsets in doubles)
  struct IdDims
  {
    int nj, nq, nv, na, nfeet;
    int nfmax, nx, ldn, me, mi;
    int o_q, o_v, o_M, o_h, o_JcT, o_H, o_g, o_A, o_b, o_C, o_l, o_u, o_Jt, o_bt, o_fr;
    int o_x, o_xk, o_grad, o_dx, o_xt, o_hx, o_y, o_re, o_re2, o_z, o_si, o_si2, o_sc, o_act, o_part, o_pairs;
    int ws_stride;                           // doubles per instance of the warm-start record in global memory
    int o_K;                                 // Newton matrix; the kinematics scratch below aliases it
    int o_oMi, o_Jw, o_ov, o_oa, o_Y, o_fs, o_Fc;
    int total;
  };

  inline IdDims id_make_dims(const smpc_model_desc & m, const smpc_id_desc & D)
  {
    IdDims L;
    L.nj = m.njoints, L.nq = m.nq, L.nv = m.nv, L.na = m.nv - 6, L.nfeet = m.nfeet;
    int nf = 0, ncone = 0, nmot = 0;
    for (int f = 0; f < m.nfeet; f++)
    {
      const bool quad = D.foot_type[f] == SMPC_ID_QUAD;
      nf += quad ? 12 : 3;
      ncone += quad ? 17 : 5;
      nmot += quad ? 6 : 3;
    }
    L.nfmax = nf;
    L.nx = m.nv + nf;
    L.ldn = L.nx | 1;
    L.me = 6 + (D.settings.contact_motion_equality ? nmot : 0);
    L.mi = ncone + 2 * L.na;
    L.ws_stride = 2 + L.nx + L.me + L.mi;
    int o = 0;
    auto take = [&](int n) {
      const int at = o;
      o += (n + 1) & ~1;
      return at;
    };
    L.o_q = take(L.nq), L.o_v = take(L.nv);
    L.o_M = take(L.nv * L.nv), L.o_h = take(L.nv), L.o_JcT = take(L.nv * (nf > 0 ? nf : 1));
    L.o_H = take(L.nx * L.ldn), L.o_g = take(L.nx);
    L.o_A = take(L.me * L.ldn), L.o_b = take(L.me);
    L.o_C = take(L.mi * L.ldn), L.o_l = take(L.mi), L.o_u = take(L.mi);
    L.o_Jt = take(6 * L.nv), L.o_bt = take(8), L.o_fr = take(48);
    L.o_x = take(L.nx), L.o_xk = take(L.nx), L.o_grad = take(L.nx), L.o_dx = take(L.nx), L.o_xt = take(L.nx), L.o_hx = take(L.nx);
    L.o_y = take(L.me), L.o_re = take(L.me), L.o_re2 = take(L.me);
    L.o_z = take(L.mi), L.o_si = take(L.mi), L.o_si2 = take(L.mi);
    L.o_sc = take(16);
    L.o_act = take((L.mi + 2) / 2 + 1); // int list of the active inequality rows
#ifdef __CUDACC__
    L.o_part = take(4 * ID_RED_SLOTS);  // one value per warp and reduction slot
#else
    L.o_part = take(ID_THREADS * ID_RED_SLOTS); // the emulation keeps one value per thread
#endif
    L.o_pairs = take((L.nx * (L.nx + 1) / 2 + L.nx + 1) / 2); // int table of the (i, j) pairs of the factorisation
    // the per-foot assembly scratch (frame records, Jacobians, right-hand sides) aliases everything from o_x up to here: the
    // solver's vectors, the reduction partials and the pair table are first written when the solve starts
    {
      const int need = L.nfeet * (24 + 6 * L.nv + 6);
      if (o - L.o_x < need)
        take(need - (o - L.o_x));
    }
    L.o_K = o;
    L.o_oMi = take(12 * L.nj), L.o_Jw = take(6 * L.nv), L.o_ov = take(6 * L.nj), L.o_oa = take(6 * L.nj);
    L.o_Y = take(36 * L.nj), L.o_fs = take(6 * L.nj), L.o_Fc = take(6 * L.nv);
    const int endK = L.o_K + (((L.nx + 1) * L.ldn + 1) & ~1); // + the right-hand-side row
    L.total = o > endK ? o : endK;
    return L;
  }

  // scalar slots in sm[o_sc + ...]
  enum
  {
    SC_VAL = 0,
    SC_VT,
    SC_GN,
    SC_GS,
    SC_SLOPE,
    SC_SAME,
    SC_PRES,
    SC_DRES,
    SC_OK,
    SC_BAD,
    SC_K0 // four consecutive slots: primal residual / scale, dual residual / scale
  };

  // TaskJointPosVelAccBounds::computeAccLimits (position, velocity, viability; no acceleration bound) — oracle id_acc_bounds
  SMPC_DEV void id_acc_bounds(double q, double dq, double qmin, double qmax, double dqmax, double dt, double & lb, double & ub)
  {
    const double two_dt_sq = 2.0 / (dt * dt);
    const double max_q3 = two_dt_sq * (qmax - q - dt * dq);
    const double min_q3 = two_dt_sq * (qmin - q - dt * dq);
    const double mdq_dt = -dq / dt;
    double lbp, ubp;
    if (dq <= 0.0)
    {
      ubp = max_q3;
      if (min_q3 < mdq_dt)
        lbp = min_q3;
      else if (q != qmin)
        lbp = fmax(dq * dq / (2.0 * (q - qmin)), mdq_dt);
      else
        lbp = ID_BIG;
    }
    else
    {
      lbp = min_q3;
      if (max_q3 > mdq_dt)
        ubp = max_q3;
      else if (q != qmax)
        ubp = fmin(-dq * dq / (2.0 * (qmax - q)), mdq_dt);
      else
        ubp = -ID_BIG;
    }
    const double lbv = (-dqmax - dq) / dt, ubv = (dqmax - dq) / dt;
    const double dt2 = dt * dt, two_a = 2.0 * dt2, qn = q + dt * dq;
    const double b1 = 2.0 * dt * dq + ID_DDQ_MAX * dt2;
    const double c1 = dq * dq - 2.0 * ID_DDQ_MAX * (qmax - qn);
    const double d1 = b1 * b1 - 2.0 * two_a * c1;
    double ddq1 = mdq_dt;
    if (d1 >= 0.0)
      ddq1 = (b1 >= 0.0) ? (-2.0 * c1) / (b1 + sqrt(d1)) : (-b1 + sqrt(d1)) / two_a;
    const double b2 = 2.0 * dt * dq - ID_DDQ_MAX * dt2;
    const double c2 = 2.0 * ID_DDQ_MAX * (qn - qmin) - dq * dq;
    const double d2 = b2 * b2 - 2.0 * two_a * c2;
    double ddq2 = mdq_dt;
    if (d2 >= 0.0)
      ddq2 = (b2 <= 0.0) ? (2.0 * c2) / (-b2 + sqrt(d2)) : (-b2 - sqrt(d2)) / two_a;
    const double ubvia = fmax(ddq1, mdq_dt), lbvia = fmin(ddq2, mdq_dt);
    lb = fmax(fmax(lbp, lbv), fmax(lbvia, -ID_BIG));
    ub = fmin(fmin(ubp, ubv), fmin(ubvia, ID_BIG));
    if (ub < lb)
    {
      const double mid = (dq <= 0.0) ? lb : ub;
      lb = ub = mid;
    }
  }

  // y = Y6 x for a dense row-major 6x6
  SMPC_DEV void id_m6v(const double * Y, const double * x, double * y)
  {
#pragma unroll
    for (int r = 0; r < 6; r++)
    {
      double s = 0;
#pragma unroll
      for (int c = 0; c < 6; c++)
        s += Y[6 * r + c] * x[c];
      y[r] = s;
    }
  }

  // ---- phase group 1: M and h ------------------------------------------------------------------------------------
  SMPC_DEV void id_dynamics(const DevModel & Mo, const IdDims & L, double * sm)
  {
    const smpc_model_desc & d = Mo.d;
    const int nj = L.nj, nv = L.nv;
    double * qs = sm + L.o_q;
    double * vs = sm + L.o_v;
    double * oMi = sm + L.o_oMi;
    double * Jw = sm + L.o_Jw; // 6 x nv, row-major
    double * ov = sm + L.o_ov;
    double * oa = sm + L.o_oa;
    double * Y = sm + L.o_Y;
    double * fs = sm + L.o_fs;
    double * Fc = sm + L.o_Fc;
    double * Mm = sm + L.o_M;
    double * h = sm + L.o_h;
    for (int l = 0; l < Mo.nlevels; l++)
    {
      SMPC_PAR(tid)
      for (int k = Mo.level_start[l] + tid; k < Mo.level_start[l + 1]; k += SMPC_NT)
      {
        const int i = Mo.level_joints[k];
        const int p = d.parent[i];
        const int iv = d.idx_v[i];
        double Rj[9], pj[3] = {0, 0, 0};
        const bool ff = d.jtype[i] == SMPC_JOINT_FREEFLYER;
        if (ff)
        {
          const double * qq = qs + d.idx_q[i];
          quat_to_R(qq + 3, Rj);
          pj[0] = qq[0], pj[1] = qq[1], pj[2] = qq[2];
        }
        else
        {
          const double ang = qs[d.idx_q[i]];
          const double s = sin(ang), c = cos(ang);
          double W[9], W2[9];
          skew(d.axis[i], W);
          m3m(W, W, W2);
#pragma unroll
          for (int e = 0; e < 9; e++)
            Rj[e] = ((e % 4 == 0) ? 1.0 : 0.0) + s * W[e] + (1 - c) * W2[e];
        }
        // liMi = placement * joint motion ; oMi = oMp * liMi
        double Rl[9], pl[3], t3[3];
        m3m(d.place_R[i], Rj, Rl);
        m3v(d.place_R[i], pj, t3);
        for (int e = 0; e < 3; e++)
          pl[e] = d.place_p[i][e] + t3[e];
        double * Ri = oMi + 12 * i;
        double * pi = Ri + 9;
        if (p < 0)
        {
          for (int e = 0; e < 9; e++)
            Ri[e] = Rl[e];
          for (int e = 0; e < 3; e++)
            pi[e] = pl[e];
        }
        else
        {
          const double * Rp = oMi + 12 * p;
          double Rn[9];
          m3m(Rp, Rl, Rn);
          m3v(Rp, pl, t3);
          for (int e = 0; e < 9; e++)
            Ri[e] = Rn[e];
          for (int e = 0; e < 3; e++)
            pi[e] = Rp[9 + e] + t3[e];
        }
        // world-frame Jacobian columns, joint velocity, body velocity / drift acceleration
        double vJ[6] = {0, 0, 0, 0, 0, 0};
        const int ndof = ff ? 6 : 1;
        for (int c = 0; c < ndof; c++)
        {
          double S[6] = {0, 0, 0, 0, 0, 0};
          if (ff)
            S[c] = 1;
          else
            S[3] = d.axis[i][0], S[4] = d.axis[i][1], S[5] = d.axis[i][2];
          double Rv[3], Rw[3], pxw[3];
          m3v(Ri, S, Rv);
          m3v(Ri, S + 3, Rw);
          v3cross(pi, Rw, pxw);
          for (int e = 0; e < 3; e++)
          {
            const double lin = Rv[e] + pxw[e];
            Jw[e * nv + iv + c] = lin;
            Jw[(3 + e) * nv + iv + c] = Rw[e];
            vJ[e] += lin * vs[iv + c];
            vJ[3 + e] += Rw[e] * vs[iv + c];
          }
        }
        double vi[6], ai[6], x6[6];
        for (int e = 0; e < 6; e++)
          vi[e] = (p >= 0 ? ov[6 * p + e] : 0.0) + vJ[e];
        mxm(vi, vJ, x6);
        for (int e = 0; e < 6; e++)
        {
          ai[e] = (p >= 0 ? oa[6 * p + e] : 0.0) + x6[e];
          ov[6 * i + e] = vi[e];
          oa[6 * i + e] = ai[e];
        }
      }
      SMPC_END
    }
    // world-frame body inertia [[m I, -m[c]x], [m[c]x, Ic - m[c]x[c]x]] and bias force I a + v x* (I v): every joint at once
    {
      SMPC_PAR(tid)
      for (int i = tid; i < nj; i += SMPC_NT)
      {
        const double * Ri = oMi + 12 * i;
        const double * pi = Ri + 9;
        double vi[6], ai[6];
        for (int e = 0; e < 6; e++)
        {
          vi[e] = ov[6 * i + e];
          ai[e] = oa[6 * i + e];
        }
        double cw[3], Icw[9], T[9], Cx[9], CC[9];
        m3v(Ri, d.lever[i], cw);
        for (int e = 0; e < 3; e++)
          cw[e] += pi[e];
        m3m(Ri, d.inertia[i], T);
        m3mt(T, Ri, Icw);
        skew(cw, Cx);
        m3m(Cx, Cx, CC);
        double * I6 = Y + 36 * i;
        const double ms = d.mass[i];
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++)
          {
            I6[6 * r + c] = (r == c) ? ms : 0.0;
            I6[6 * r + 3 + c] = -ms * Cx[3 * r + c];
            I6[6 * (3 + r) + c] = ms * Cx[3 * r + c];
            I6[6 * (3 + r) + 3 + c] = Icw[3 * r + c] - ms * CC[3 * r + c];
          }
        double hi[6], Ia[6], vxh[6];
        id_m6v(I6, vi, hi);
        id_m6v(I6, ai, Ia);
        mxf(vi, hi, vxh);
        for (int e = 0; e < 6; e++)
          fs[6 * i + e] = Ia[e] + vxh[e];
      }
      SMPC_END
    }
    // composite inertias and subtree bias forces, deepest level first; each (parent, entry) sums its children in index order
    for (int l = Mo.nlevels - 2; l >= 0; l--)
    {
      SMPC_PAR(tid)
      const int n0 = Mo.level_start[l], cnt = Mo.level_start[l + 1] - n0;
      for (int idx = tid; idx < cnt * 42; idx += SMPC_NT)
      {
        const int p = Mo.level_joints[n0 + idx / 42], e = idx % 42;
        double * dst = (e < 36) ? &Y[36 * p + e] : &fs[6 * p + e - 36];
        double s = *dst;
        for (int c = p + 1; c < nj; c++)
          if (d.parent[c] == p)
            s += (e < 36) ? Y[36 * c + e] : fs[6 * c + e - 36];
        *dst = s;
      }
      SMPC_END
    }
    SMPC_PAR(tid)
    for (int idx = tid; idx < 6 * nv; idx += SMPC_NT)
    {
      const int r = idx / nv, c = idx % nv;
      const double * Yi = Y + 36 * Mo.joint_of_dof[c] + 6 * r;
      double s = 0;
#pragma unroll
      for (int e = 0; e < 6; e++)
        s += Yi[e] * Jw[e * nv + c];
      Fc[r * nv + c] = s;
    }
    SMPC_END
    SMPC_PAR(tid)
    for (int idx = tid; idx < nv * nv + nv; idx += SMPC_NT)
    {
      if (idx < nv * nv)
      {
        const int j = idx / nv, c = idx % nv;
        if (j > c)
          continue;
        double s = 0;
        if ((Mo.anc_mask[Mo.joint_of_dof[c]] >> Mo.joint_of_dof[j]) & 1u)
        {
#pragma unroll
          for (int e = 0; e < 6; e++)
            s += Jw[e * nv + j] * Fc[e * nv + c];
        }
        Mm[j * nv + c] = s;
        Mm[c * nv + j] = s;
      }
      else
      {
        const int c = idx - nv * nv;
        const int i = Mo.joint_of_dof[c];
        const double g6[6] = {0, 0, -9.81, 0, 0, 0};
        double Yg[6];
        id_m6v(Y + 36 * i, g6, Yg);
        double s = 0;
#pragma unroll
        for (int e = 0; e < 6; e++)
          s += Jw[e * nv + c] * (fs[6 * i + e] - Yg[e]);
        h[c] = s;
      }
    }
    SMPC_END
  }

  // LOCAL-frame Jacobian (into Jt, 6 x nv), velocity and classical drift of the frame (Rl, pl) carried by joint jf.
  // fr[0..8] = R, fr[9..11] = p (world placement), fr[12..17] = v_local, fr[18..23] = drift_local
  SMPC_DEV void id_frame(const DevModel & Mo, const IdDims & L, double * sm, int jf, const double * Rl, const double * pl)
  {
    const int nv = L.nv;
    double * fr = sm + L.o_fr;
    const double * oMi = sm + L.o_oMi + 12 * jf;
    SMPC_PAR(tid)
    if (tid == 0)
    {
      double R[9], t[3], p[3];
      m3m(oMi, Rl, R);
      m3v(oMi, pl, t);
      for (int e = 0; e < 3; e++)
        p[e] = oMi[9 + e] + t[e];
      for (int e = 0; e < 9; e++)
        fr[e] = R[e];
      for (int e = 0; e < 3; e++)
        fr[9 + e] = p[e];
      const double * v = sm + L.o_ov + 6 * jf;
      const double * a = sm + L.o_oa + 6 * jf;
      double pd[3], pdd[3], t1[3], t2[3];
      v3cross(v + 3, p, t);
      for (int e = 0; e < 3; e++)
        pd[e] = v[e] + t[e];
      v3cross(a + 3, p, t1);
      v3cross(v + 3, pd, t2);
      for (int e = 0; e < 3; e++)
        pdd[e] = a[e] + t1[e] + t2[e];
      m3tv(R, pd, fr + 12);
      m3tv(R, v + 3, fr + 15);
      m3tv(R, pdd, fr + 18);
      m3tv(R, a + 3, fr + 21);
    }
    SMPC_END
    SMPC_PAR(tid)
    const double * Jw = sm + L.o_Jw;
    double * Jt = sm + L.o_Jt;
    for (int c = tid; c < nv; c += SMPC_NT)
    {
      double a3[3] = {0, 0, 0}, b3[3] = {0, 0, 0};
      if ((Mo.anc_mask[jf] >> Mo.joint_of_dof[c]) & 1u)
      {
        const double lin0[3] = {Jw[c], Jw[nv + c], Jw[2 * nv + c]};
        const double ang[3] = {Jw[3 * nv + c], Jw[4 * nv + c], Jw[5 * nv + c]};
        double t[3], lin[3];
        v3cross(ang, fr + 9, t);
        for (int e = 0; e < 3; e++)
          lin[e] = lin0[e] + t[e];
        m3tv(fr, lin, a3);
        m3tv(fr, ang, b3);
      }
      for (int e = 0; e < 3; e++)
      {
        Jt[e * nv + c] = a3[e];
        Jt[(3 + e) * nv + c] = b3[e];
      }
    }
    SMPC_END
  }

  // H += w Jt(rows)^T Jt(rows) on the acceleration block, g -= w Jt(rows)^T bt(rows)
  SMPC_DEV void id_add_motion_task(const IdDims & L, double * sm, double w, int r0, int r1)
  {
    const int nv = L.nv, ld = L.ldn;
    SMPC_PAR(tid)
    const double * Jt = sm + L.o_Jt;
    const double * bt = sm + L.o_bt;
    double * H = sm + L.o_H;
    double * g = sm + L.o_g;
    for (int idx = tid; idx < nv * nv + nv; idx += SMPC_NT)
    {
      if (idx < nv * nv)
      {
        const int i = idx / nv, j = idx % nv;
        double s = H[i * ld + j];
        for (int r = r0; r < r1; r++)
          s += (w * Jt[r * nv + i]) * Jt[r * nv + j];
        H[i * ld + j] = s;
      }
      else
      {
        const int i = idx - nv * nv;
        double s = g[i];
        for (int r = r0; r < r1; r++)
          s -= (w * Jt[r * nv + i]) * bt[r];
        g[i] = s;
      }
    }
    SMPC_END
  }

  // ---- phase group 2: the QP data -----------------------------------------------------------------------------------
  // returns the layout through sm[o_sc + 10..12] = n, me, mi (as doubles) — written by thread 0
  SMPC_DEV void id_build_qp(const DevModel & Mo, const smpc_id_desc & D, const IdTargetDev & T, const IdDims & L, double * sm, int & n_out, int & me_out, int & mi_out)
  {
    const smpc_model_desc & d = Mo.d;
    const smpc_id_settings & S = D.settings;
    const int nv = L.nv, na = L.na, nfeet = L.nfeet, ld = L.ldn;
    double * H = sm + L.o_H;
    double * g = sm + L.o_g;
    double * A = sm + L.o_A;
    double * bq = sm + L.o_b;
    double * Cm = sm + L.o_C;
    double * lo = sm + L.o_l;
    double * up = sm + L.o_u;
    double * Jt = sm + L.o_Jt;
    double * bt = sm + L.o_bt;
    double * fr = sm + L.o_fr;
    double * JcT = sm + L.o_JcT;
    const double * Mm = sm + L.o_M;
    const double * h = sm + L.o_h;
    const double * qs = sm + L.o_q;
    const double * vs = sm + L.o_v;
    // variable layout (uniform over the CTA: computed redundantly by every thread)
    int foff[SMPC_MAX_FEET], fdim[SMPC_MAX_FEET], nf = 0, ncone = 0, nmot = 0;
    for (int f = 0; f < nfeet; f++)
    {
      foff[f] = nv + nf;
      fdim[f] = 0;
      if (!T.active[f])
        continue;
      const bool quad = D.foot_type[f] == SMPC_ID_QUAD;
      fdim[f] = quad ? 12 : 3;
      nf += fdim[f];
      ncone += quad ? 17 : 5;
      nmot += quad ? 6 : 3;
    }
    const int n = nv + nf;
    const int me = 6 + (S.contact_motion_equality ? nmot : 0);
    const int mi = ncone + 2 * na;
    const int ldf = nf > 0 ? nf : 1;
    n_out = n, me_out = me, mi_out = mi;
    const double weight = Mo.total_mass * 9.81;
    const double fmaxn = S.contact_weight_ratio_max * weight, fminn = S.contact_weight_ratio_min * weight;

    SMPC_PAR(tid)
    for (int idx = tid; idx < n * ld; idx += SMPC_NT)
      H[idx] = (idx / ld == idx % ld) ? ID_HESSIAN_REG : 0.0;
    for (int idx = tid; idx < n; idx += SMPC_NT)
      g[idx] = 0.0;
    for (int idx = tid; idx < me * ld; idx += SMPC_NT)
      A[idx] = 0.0;
    for (int idx = tid; idx < mi * ld; idx += SMPC_NT)
      Cm[idx] = 0.0;
    SMPC_END

    // posture task: diagonal
    if (S.w_posture > 0.)
    {
      SMPC_PAR(tid)
      const double kp = S.kp_posture, kd = 2.0 * sqrt(kp);
      for (int j = tid; j < na; j += SMPC_NT)
      {
        const double ades = -kp * (qs[7 + j] - T.q_post[j]) - kd * (vs[6 + j] - T.v_post[j]) + T.a_post[j];
        H[(6 + j) * ld + 6 + j] += S.w_posture;
        g[6 + j] -= S.w_posture * ades;
      }
      SMPC_END
    }
    // base task (LOCAL frame; velocity reference = acceleration target, reference quirk)
    if (S.w_base > 0.)
    {
      id_frame(Mo, L, sm, d.base_joint, d.base_R, d.base_p);
      SMPC_PAR(tid)
      if (tid == 0)
      {
        const double kp = S.kp_base, kd = 2.0 * sqrt(kp);
        double Rr[9], dp[3], pr[3], e6[6];
        m3tm(fr, T.base_R, Rr);
        for (int e = 0; e < 3; e++)
          dp[e] = T.base_p[e] - fr[9 + e];
        m3tv(fr, dp, pr);
        log6(Rr, pr, e6);
        for (int r = 0; r < 6; r++)
          bt[r] = kp * e6[r] + kd * (T.base_acc[r] - fr[12 + r]) - fr[18 + r];
      }
      SMPC_END
      id_add_motion_task(L, sm, S.w_base, S.centroidal ? 3 : 0, 6);
    }
    // centre-of-mass task
    if (S.centroidal && S.w_com > 0.)
    {
      SMPC_PAR(tid)
      const double * Y = sm + L.o_Y;
      const double * Jw = sm + L.o_Jw;
      const double Mtot = Y[0];
      for (int c = tid; c < nv; c += SMPC_NT)
      {
        const int i = Mo.joint_of_dof[c];
        const double * Yi = Y + 36 * i;
        const double msub = Yi[0];
        const double cs[3] = {Yi[6 * 5 + 1] / msub, Yi[6 * 3 + 2] / msub, Yi[6 * 4 + 0] / msub};
        const double ang[3] = {Jw[3 * nv + c], Jw[4 * nv + c], Jw[5 * nv + c]};
        double t[3];
        v3cross(ang, cs, t);
        for (int e = 0; e < 3; e++)
          Jt[e * nv + c] = (msub / Mtot) * (Jw[e * nv + c] + t[e]);
      }
      if (tid == 0)
      { // com / drift from the composite inertia and the subtree bias force of the root; -kd * vcom is added below
        const double kp = S.kp_com, kd = 2.0 * sqrt(kp);
        const double * fs0 = sm + L.o_fs;
        for (int r = 0; r < 3; r++)
        {
          const double com = Y[(r == 0) ? 6 * 5 + 1 : (r == 1) ? 6 * 3 + 2 : 6 * 4 + 0] / Mtot;
          bt[r] = -kp * (com - T.com_ref[r]) + kd * T.com_vel[r] - fs0[r] / Mtot;
        }
        bt[6] = kd;
      }
      SMPC_END
      SMPC_PAR(tid)
      if (tid < 3)
      { // vcom_r = Jcom(r, :) v
        double s = 0;
        for (int c = 0; c < nv; c++)
          s += Jt[tid * nv + c] * vs[c];
        bt[tid] -= bt[6] * s;
      }
      SMPC_END
      id_add_motion_task(L, sm, S.w_com, 0, 3);
    }
    // ---- feet: swing-foot tracking tasks (LOCAL frame) and contacts, all feet at once -----------------------------------
    // scratch (aliases the solver's vectors / reduction partials / pair table, unused until the solve): per foot a frame record
    // [R 9, p 3, v_local 6, drift_local 6], a 6 x nv Jacobian and the right-hand side of its motion rows
    double * frs = sm + L.o_x;
    double * Jts = frs + 24 * nfeet;
    double * bts = Jts + 6 * nv * nfeet;
    const bool track = S.centroidal && S.w_feet_tracking > 0.;
    int eoff[SMPC_MAX_FEET], ioff[SMPC_MAX_FEET]; // first equality / inequality row of each active contact
    {
      int ie_ = 6, ii_ = 0;
      for (int f = 0; f < nfeet; f++)
      {
        eoff[f] = ie_, ioff[f] = ii_;
        if (!T.active[f])
          continue;
        const bool quad = D.foot_type[f] == SMPC_ID_QUAD;
        ie_ += S.contact_motion_equality ? (quad ? 6 : 3) : 0;
        ii_ += quad ? 17 : 5;
      }
    }
    const int ii = ncone;
    // phase A: frame placement, local velocity and classical drift of every foot in use (one thread per foot)
    SMPC_PAR(tid)
    for (int f = tid; f < nfeet; f += SMPC_NT)
    {
      if (!T.active[f] && !(track && T.tracked[f]))
        continue;
      const int jf = d.foot_joint[f];
      const double * oMi = sm + L.o_oMi + 12 * jf;
      double * frf = frs + 24 * f;
      double R[9], t[3], p[3];
      m3m(oMi, d.foot_R[f], R);
      m3v(oMi, d.foot_p[f], t);
      for (int e = 0; e < 3; e++)
        p[e] = oMi[9 + e] + t[e];
      for (int e = 0; e < 9; e++)
        frf[e] = R[e];
      for (int e = 0; e < 3; e++)
        frf[9 + e] = p[e];
      const double * vj = sm + L.o_ov + 6 * jf;
      const double * aj = sm + L.o_oa + 6 * jf;
      double pd[3], pdd[3], t1[3], t2[3];
      v3cross(vj + 3, p, t);
      for (int e = 0; e < 3; e++)
        pd[e] = vj[e] + t[e];
      v3cross(aj + 3, p, t1);
      v3cross(vj + 3, pd, t2);
      for (int e = 0; e < 3; e++)
        pdd[e] = aj[e] + t1[e] + t2[e];
      m3tv(R, pd, frf + 12);
      m3tv(R, vj + 3, frf + 15);
      m3tv(R, pdd, frf + 18);
      m3tv(R, aj + 3, frf + 21);
    }
    SMPC_END
    // phase B: Jacobians (LOCAL; the linear rows of an active point contact in world axes) and motion right-hand sides
    SMPC_PAR(tid)
    const double * Jw = sm + L.o_Jw;
    for (int idx = tid; idx < nfeet * nv + nfeet; idx += SMPC_NT)
    {
      if (idx < nfeet * nv)
      {
        const int f = idx / nv, c = idx % nv;
        if (!T.active[f] && !(track && T.tracked[f]))
          continue;
        const double * frf = frs + 24 * f;
        double a3[3] = {0, 0, 0}, b3[3] = {0, 0, 0};
        if ((Mo.anc_mask[d.foot_joint[f]] >> Mo.joint_of_dof[c]) & 1u)
        {
          const double lin0[3] = {Jw[c], Jw[nv + c], Jw[2 * nv + c]};
          const double ang[3] = {Jw[3 * nv + c], Jw[4 * nv + c], Jw[5 * nv + c]};
          double t[3], lin[3];
          v3cross(ang, frf + 9, t);
          for (int e = 0; e < 3; e++)
            lin[e] = lin0[e] + t[e];
          m3tv(frf, lin, a3);
          m3tv(frf, ang, b3);
          if (T.active[f] && D.foot_type[f] != SMPC_ID_QUAD)
          { // ContactPoint with useLocalFrame(false): rotate the local linear rows back to world axes, as the oracle does
            double w3[3];
            m3v(frf, a3, w3);
            a3[0] = w3[0], a3[1] = w3[1], a3[2] = w3[2];
          }
        }
        double * Jf = Jts + 6 * nv * f;
        for (int e = 0; e < 3; e++)
        {
          Jf[e * nv + c] = a3[e];
          Jf[(3 + e) * nv + c] = b3[e];
        }
      }
      else
      {
        const int f = idx - nfeet * nv;
        const double * frf = frs + 24 * f;
        double * bf = bts + 6 * f;
        if (T.active[f])
        {
          const double kd = 2.0 * sqrt(S.kp_contact);
          if (D.foot_type[f] == SMPC_ID_QUAD)
            for (int r = 0; r < 6; r++)
              bf[r] = -kd * frf[12 + r] - frf[18 + r];
          else
          {
            double vw[3], dw[3];
            m3v(frf, frf + 12, vw);
            m3v(frf, frf + 18, dw);
            for (int r = 0; r < 3; r++)
              bf[r] = -kd * vw[r] - dw[r];
          }
        }
        else if (track && T.tracked[f])
        {
          const double kp = S.kp_feet_tracking, kd = 2.0 * sqrt(kp);
          double Rr[9], dp[3], pr[3], e6[6], vr[6];
          m3tm(frf, T.foot_R[f], Rr);
          for (int e = 0; e < 3; e++)
            dp[e] = T.foot_p[f][e] - frf[9 + e];
          m3tv(frf, dp, pr);
          log6(Rr, pr, e6);
          m3tv(frf, T.foot_vel[f], vr);
          m3tv(frf, T.foot_vel[f] + 3, vr + 3);
          for (int r = 0; r < 6; r++)
            bf[r] = kp * e6[r] + kd * (vr[r] - frf[12 + r]) - frf[18 + r];
        }
      }
    }
    SMPC_END
    // phase C: everything the feet contribute — motion tasks into the acceleration block of (H, g) (tracking feet first, then
    // contacts, each in foot order: the oracle's order of accumulation), columns of Jc^T T, equality rows, force regularisation
    // blocks, cone rows
    SMPC_PAR(tid)
    const double wmot = (!S.contact_motion_equality && S.w_contact_motion > 0.) ? S.w_contact_motion : 0.0;
    for (int idx = tid; idx < nv * nv + nv; idx += SMPC_NT)
    {
      const bool isH = idx < nv * nv;
      const int i = isH ? idx / nv : idx - nv * nv, j = isH ? idx % nv : 0;
      double acc = isH ? H[i * ld + j] : g[i];
      for (int pass = 0; pass < 2; pass++)
        for (int f = 0; f < nfeet; f++)
        {
          const bool use = pass == 0 ? (track && T.tracked[f] && !T.active[f]) : (T.active[f] != 0);
          const double w = pass == 0 ? S.w_feet_tracking : wmot;
          if (!use || w == 0.0)
            continue;
          const int rows = (D.foot_type[f] == SMPC_ID_QUAD) ? 6 : 3;
          const double * Jf = Jts + 6 * nv * f;
          if (isH)
            for (int r = 0; r < rows; r++)
              acc += (w * Jf[r * nv + i]) * Jf[r * nv + j];
          else
            for (int r = 0; r < rows; r++)
              acc -= (w * Jf[r * nv + i]) * bts[6 * f + r];
        }
      if (isH)
        H[i * ld + j] = acc;
      else
        g[i] = acc;
    }
    // generalised force of each force variable: Jm^T T(:, j)
    for (int idx = tid; idx < nv * nf; idx += SMPC_NT)
    {
      const int c = idx / nf, jg = idx % nf;
      int f = 0;
      while (f + 1 < nfeet && (fdim[f] == 0 || jg >= foff[f] - nv + fdim[f]))
        f++;
      const int j = jg - (foff[f] - nv);
      const bool quad = D.foot_type[f] == SMPC_ID_QUAD;
      const int rows = quad ? 6 : 3;
      double Tc[6] = {0, 0, 0, 0, 0, 0};
      if (quad)
      {
        const int pt = j / 3, ax = j % 3;
        double e3[3] = {0, 0, 0}, mo[3];
        e3[ax] = 1;
        v3cross(D.quad_points[f][pt], e3, mo);
        Tc[ax] = 1;
        Tc[3] = mo[0], Tc[4] = mo[1], Tc[5] = mo[2];
      }
      else
        Tc[j] = 1;
      const double * Jf = Jts + 6 * nv * f;
      double s_ = 0;
      for (int r = 0; r < rows; r++)
        s_ += Jf[r * nv + c] * Tc[r];
      JcT[c * ldf + jg] = s_;
    }
    if (S.contact_motion_equality)
      for (int idx = tid; idx < nfeet * 6 * (nv + 1); idx += SMPC_NT)
      {
        const int f = idx / (6 * (nv + 1)), rem = idx % (6 * (nv + 1)), r = rem / (nv + 1), c = rem % (nv + 1);
        if (!T.active[f] || r >= ((D.foot_type[f] == SMPC_ID_QUAD) ? 6 : 3))
          continue;
        if (c < nv)
          A[(eoff[f] + r) * ld + c] = Jts[6 * nv * f + r * nv + c];
        else
          bq[eoff[f] + r] = bts[6 * f + r];
      }
    // force regularisation |W T f - W f_ref|^2: one thread per entry of each contact's block of H (+ one per entry of g)
    if (S.w_contact_force > 0.)
      for (int idx = tid; idx < nfeet * SMPC_ID_MAX_FVAR * (SMPC_ID_MAX_FVAR + 1); idx += SMPC_NT)
      {
        const int f = idx / (SMPC_ID_MAX_FVAR * (SMPC_ID_MAX_FVAR + 1)), rem = idx % (SMPC_ID_MAX_FVAR * (SMPC_ID_MAX_FVAR + 1));
        const int i = rem / (SMPC_ID_MAX_FVAR + 1), j = rem % (SMPC_ID_MAX_FVAR + 1);
        const int nfv = fdim[f];
        if (!T.active[f] || i >= nfv || (j >= nfv && j != SMPC_ID_MAX_FVAR))
          continue;
        const bool quad = D.foot_type[f] == SMPC_ID_QUAD;
        const int rows = quad ? 6 : 3, c0 = foff[f];
        const double wreg[6] = {1, 1, 1e-3, 2, 2, 2};
        double acc = (j == SMPC_ID_MAX_FVAR) ? g[c0 + i] : H[(c0 + i) * ld + c0 + j];
        for (int r = 0; r < rows; r++)
        {
          double ti = 0, tj = 0;
          if (quad)
          {
            for (int side = 0; side < 2; side++)
            {
              const int k = side == 0 ? i : j;
              if (k >= nfv)
                continue;
              const int pt = k / 3, ax = k % 3;
              double t = 0;
              if (r < 3)
                t = (r == ax) ? 1.0 : 0.0;
              else
              {
                double e3[3] = {0, 0, 0}, mo[3];
                e3[ax] = 1;
                v3cross(D.quad_points[f][pt], e3, mo);
                t = mo[r - 3];
              }
              (side == 0 ? ti : tj) = t;
            }
          }
          else
          {
            ti = (r == i) ? 1.0 : 0.0;
            tj = (r == j) ? 1.0 : 0.0;
          }
          const double ri = wreg[r] * ti;
          if (ri == 0.0)
            continue;
          if (j == SMPC_ID_MAX_FVAR)
            acc -= S.w_contact_force * ri * (wreg[r] * T.f_ref[f][r]);
          else
            acc += S.w_contact_force * ri * (wreg[r] * tj);
        }
        if (j == SMPC_ID_MAX_FVAR)
          g[c0 + i] = acc;
        else
          H[(c0 + i) * ld + c0 + j] = acc;
      }
    // force inequalities: friction pyramid per contact point, bounds on the (total) normal force: one thread per row
    for (int idx = tid; idx < nfeet * 17; idx += SMPC_NT)
    {
      const int f = idx / 17, r = idx % 17;
      if (!T.active[f])
        continue;
      const bool quad = D.foot_type[f] == SMPC_ID_QUAD;
      const int npts = quad ? 4 : 1, c0 = foff[f];
      if (r > 4 * npts)
        continue;
      const int row = ioff[f] + r;
      if (r < 4 * npts)
      {
        const double mu_ = S.friction_coefficient;
        const double pyr[4][3] = {{-1, 0, -mu_}, {1, 0, -mu_}, {0, -1, -mu_}, {0, 1, -mu_}};
        const int pt = r / 4;
        for (int ax = 0; ax < 3; ax++)
          Cm[row * ld + c0 + 3 * pt + ax] = pyr[r % 4][ax];
        lo[row] = -ID_BIG;
        up[row] = 0.0;
      }
      else
      {
        for (int pt = 0; pt < npts; pt++)
          Cm[row * ld + c0 + 3 * pt + 2] = 1.0;
        lo[row] = fminn;
        up[row] = fmaxn;
      }
    }
    SMPC_END
    // floating-base rows, joint acceleration bounds, actuation bounds
    SMPC_PAR(tid)
    for (int idx = tid; idx < 6 * n + 6; idx += SMPC_NT)
    {
      if (idx < 6 * n)
      {
        const int r = idx / n, c = idx % n;
        A[r * ld + c] = (c < nv) ? Mm[r * nv + c] : -JcT[r * ldf + c - nv];
      }
      else
        bq[idx - 6 * n] = -h[idx - 6 * n];
    }
    for (int j = tid; j < na; j += SMPC_NT)
    {
      double lb, ub;
      id_acc_bounds(qs[7 + j], vs[6 + j], D.qmin[j], D.qmax[j], D.vmax[j], D.control_dt, lb, ub);
      Cm[(ii + j) * ld + 6 + j] = 1.0;
      lo[ii + j] = lb;
      up[ii + j] = ub;
      lo[ii + na + j] = -D.effort[j] - h[6 + j];
      up[ii + na + j] = D.effort[j] - h[6 + j];
    }
    for (int idx = tid; idx < na * n; idx += SMPC_NT)
    {
      const int j = idx / n, c = idx % n;
      Cm[(ii + na + j) * ld + c] = (c < nv) ? Mm[(6 + j) * nv + c] : -JcT[(6 + j) * ldf + c - nv];
    }
    SMPC_END
    // row equilibration (the solver's own preconditioning: unit inf-norm constraint rows; x is unchanged)
    SMPC_PAR(tid)
    for (int idx = tid; idx < me + mi; idx += SMPC_NT)
    {
      double * row = (idx < me) ? A + idx * ld : Cm + (idx - me) * ld;
      double nr = 0;
      for (int c = 0; c < n; c++)
        nr = fmax(nr, fabs(row[c]));
      if (nr > 0)
      {
        const double sc = 1.0 / nr;
        for (int c = 0; c < n; c++)
          row[c] *= sc;
        if (idx < me)
          bq[idx] *= sc;
        else
        {
          lo[idx - me] *= sc;
          up[idx - me] *= sc;
        }
      }
    }
    SMPC_END
  }

  // ---- phase group 3: the QP solve -----------------------------------------------------------------------------------
  // CTA-wide reductions with ONE barrier: at the end of the producing phase every thread hands its partial to id_red_put
  // (device: shuffle reduction inside the warp, lane 0 stores the warp's value in part[4 k + warp]); after the phase's barrier
  // every thread combines the four warp values with id_red_get. Consecutive reductions use different slots k, so a fast warp
  // can never overwrite a slot another warp has not read yet. The CPU emulation keeps one value per thread and loops.
  enum
  {
    RK_VAL0 = 0, // subproblem value (two slots, alternating between consecutive evaluations)
    RK_VAL1,
    RK_GN,
    RK_GS,
    RK_SLOPE,
    RK_CHANGED,
    RK_PR,
    RK_PS,
    RK_DR,
    RK_DS
  };
  template <bool MAXRED>
  SMPC_DEV void id_red_put(double * part, int k, int tid, double v)
  {
#ifdef __CUDACC__
#pragma unroll
    for (int o = 16; o; o >>= 1)
    {
      const double t = __shfl_xor_sync(0xffffffffu, v, o);
      v = MAXRED ? fmax(v, t) : v + t;
    }
    if ((tid & 31) == 0)
      part[4 * k + (tid >> 5)] = v;
#else
    part[k * ID_THREADS + tid] = v;
#endif
  }
  template <bool MAXRED>
  SMPC_DEV double id_red_get(const double * part, int k)
  {
#ifdef __CUDACC__
    const double * p = part + 4 * k;
    return MAXRED ? fmax(fmax(p[0], p[1]), fmax(p[2], p[3])) : (p[0] + p[1]) + (p[2] + p[3]);
#else
    double s = MAXRED ? -1e300 : 0.0;
    for (int t = 0; t < ID_THREADS; t++)
      s = MAXRED ? fmax(s, part[k * ID_THREADS + t]) : s + part[k * ID_THREADS + t];
    return s;
#endif
  }

  // value of the subproblem at xv (returned to every thread); fills hx = H xv, re_out, si_out. rk: reduction slot
  SMPC_DEV double id_al_value(const IdDims & L, double * sm, int n, int me, int mi, double mu, const double * xv, double * re_out, double * si_out, int rk)
  {
    const int ld = L.ldn;
    SMPC_PAR(tid)
    const double * H = sm + L.o_H;
    const double * A = sm + L.o_A;
    const double * Cm = sm + L.o_C;
    const double * y = sm + L.o_y;
    const double * z = sm + L.o_z;
    const double * g = sm + L.o_g;
    const double * xk = sm + L.o_xk;
    double * hx = sm + L.o_hx;
    const double himu = 0.5 / mu;
    double acc = 0;
    for (int idx = tid; idx < n + me + mi; idx += SMPC_NT)
    {
      const double * row = (idx < n) ? H + idx * ld : (idx < n + me) ? A + (idx - n) * ld : Cm + (idx - n - me) * ld;
      double s0 = 0, s1 = 0;
      int j = 0;
      for (; j + 1 < n; j += 2)
      {
        s0 += row[j] * xv[j];
        s1 += row[j + 1] * xv[j + 1];
      }
      if (j < n)
        s0 += row[j] * xv[j];
      const double s = s0 + s1;
      if (idx < n)
      {
        hx[idx] = s;
        const double dxi = xv[idx] - xk[idx];
        acc += xv[idx] * (0.5 * s + g[idx]) + 0.5 * ID_QP_RHO * dxi * dxi;
      }
      else if (idx < n + me)
      {
        const int r = idx - n;
        const double t = s - sm[L.o_b + r] + mu * y[r];
        re_out[r] = t;
        acc += himu * t * t;
      }
      else
      {
        const int r = idx - n - me;
        const double t = s + mu * z[r];
        double vi = 0;
        if (t > sm[L.o_u + r])
          vi = t - sm[L.o_u + r];
        else if (t < sm[L.o_l + r])
          vi = t - sm[L.o_l + r];
        si_out[r] = vi;
        acc += himu * vi * vi;
      }
    }
    id_red_put<false>(sm + L.o_part, rk, tid, acc);
    SMPC_END
    return id_red_get<false>(sm + L.o_part, rk);
  }

  // ws (optional, global memory): the instance's previous solution [signature, x (nx), y (me), z (mi)] with the dims' maxima
  // as strides; a solve whose layout signature matches starts from it (the QP is strictly convex: the solution does not
  // depend on the starting point, the iteration count does) and every converged solve stores its own.
  SMPC_DEV void id_solve_qp(const IdDims & L, double * sm, int n, int me, int mi, smpc_id_info * info, double * ws, double sig)
  {
    const int ld = L.ldn;
    double * x = sm + L.o_x;
    double * xk = sm + L.o_xk;
    double * xt = sm + L.o_xt;
    double * dx = sm + L.o_dx;
    double * grad = sm + L.o_grad;
    double * y = sm + L.o_y;
    double * z = sm + L.o_z;
    double * re = sm + L.o_re;
    double * re2 = sm + L.o_re2;
    double * si = sm + L.o_si;
    double * si2 = sm + L.o_si2;
    double * sc = sm + L.o_sc;
    double * part = sm + L.o_part;
    int * act = reinterpret_cast<int *>(sm + L.o_act);
    const bool warm = (ws != nullptr) && (ws[0] == sig);
    SMPC_PAR(tid)
    for (int i = tid; i < n; i += SMPC_NT)
      x[i] = warm ? ws[1 + i] : 0.0;
    for (int i = tid; i < me; i += SMPC_NT)
      y[i] = warm ? ws[1 + L.nx + i] : 0.0;
    for (int i = tid; i < mi; i += SMPC_NT)
      z[i] = warm ? ws[1 + L.nx + L.me + i] : 0.0;
    if (tid == 0)
      sc[SC_BAD] = 0.0;
    SMPC_END
    // the pairs (i, j <= i) of the lower triangle plus the right-hand-side row (n, j), listed column by column
    // (column j: rows j..n, starting at cs(j) = j (n + 1) - j (j - 1) / 2) once per solve: pairs[idx] = (i << 8) | j.
    // Step k of the factorisation touches the contiguous tail idx >= cs(k + 1), dealt round-robin over the threads.
    const int ntri = n * (n + 1) / 2, npair = ntri + n;
    int * pairs = reinterpret_cast<int *>(sm + L.o_pairs);
    SMPC_PAR(tid)
    for (int idx = tid; idx < npair; idx += SMPC_NT)
    {
      const double hb = n + 1.5;
      int j = (int)(hb - sqrt(hb * hb - 2.0 * idx));
      j = j < 0 ? 0 : (j > n - 1 ? n - 1 : j);
      while (j > 0 && j * (n + 1) - j * (j - 1) / 2 > idx)
        j--;
      while (j < n - 1 && (j + 1) * (n + 1) - (j + 1) * j / 2 <= idx)
        j++;
      const int i = j + idx - (j * (n + 1) - j * (j - 1) / 2);
      pairs[idx] = (i << 8) | j;
    }
    SMPC_END
    double mu = warm ? ID_QP_MU_WARM : ID_QP_MU0;
    int newton = 0, outer = 0, status = 1;
    double pres = 0, dres = 0;
    for (outer = 0; outer < ID_QP_MAX_OUTER; outer++)
    {
      const double imu = 1.0 / mu;
      SMPC_PAR(tid)
      for (int i = tid; i < n; i += SMPC_NT)
        xk[i] = x[i];
      SMPC_END
      double val = id_al_value(L, sm, n, me, mi, mu, x, re, si, RK_VAL0);
      for (int it = 0; it < ID_QP_MAX_INNER && newton < ID_QP_MAX_NEWTON; it++)
      {
        // gradient (hx = H x is current), its inf-norm and the magnitude of its largest term; warp 3 compacts the list of
        // active inequality rows meanwhile
        SMPC_PAR(tid)
        const double * A = sm + L.o_A;
        const double * Cm = sm + L.o_C;
        const double * g = sm + L.o_g;
        const double * hx = sm + L.o_hx;
        double gn = 0, gs = 0;
        for (int i = tid; i < n; i += SMPC_NT)
        {
          double s = g[i] + ID_QP_RHO * (x[i] - xk[i]) + hx[i];
          double sa = fabs(g[i]) + fabs(hx[i]);
          for (int r = 0; r < me; r++)
          {
            const double t = A[r * ld + i] * re[r] * imu;
            s += t;
            sa += fabs(t);
          }
          for (int r = 0; r < mi; r++)
          {
            const double t = Cm[r * ld + i] * si[r] * imu; // si = 0 on inactive rows
            s += t;
            sa += fabs(t);
          }
          grad[i] = s;
          gn = fmax(gn, fabs(s));
          gs = fmax(gs, sa);
        }
        id_red_put<true>(part, RK_GN, tid, gn);
        id_red_put<true>(part, RK_GS, tid, gs);
#ifdef __CUDACC__
        if ((tid >> 5) == 3)
        {
          const int lane = tid & 31;
          int base = 0;
          for (int r0 = 0; r0 < mi; r0 += 32)
          {
            const int r = r0 + lane;
            const bool on = (r < mi) && (si[r] != 0.0);
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on)
              act[1 + base + __popc(m & ((1u << lane) - 1u))] = r;
            base += __popc(m);
          }
          if (lane == 0)
            act[0] = base;
        }
#else
        if (tid == 0)
        {
          int na_ = 0;
          for (int r = 0; r < mi; r++)
            if (si[r] != 0.0)
              act[1 + na_++] = r;
          act[0] = na_;
        }
#endif
        SMPC_END
        if (id_red_get<true>(part, RK_GN) <= ID_QP_GRAD_REL * id_red_get<true>(part, RK_GS))
          break;
        newton++;
        // Newton matrix (lower triangle) with the right-hand side -grad as an extra row n, on the pairs this thread owns
        SMPC_PAR(tid)
        const double * A = sm + L.o_A;
        const double * Cm = sm + L.o_C;
        const double * H = sm + L.o_H;
        double * K = sm + L.o_K;
        const int nact = act[0];
        for (int idx = tid; idx < npair; idx += SMPC_NT)
        {
          const int i = pairs[idx] >> 8, j = pairs[idx] & 255;
          if (i == n)
          {
            K[n * ld + j] = -grad[j];
            continue;
          }
          double t0 = 0, t1 = 0;
          for (int r = 0; r < me; r++)
            t0 += A[r * ld + i] * A[r * ld + j];
          for (int a = 0; a < nact; a++)
          {
            const double * cr = Cm + act[1 + a] * ld;
            t1 += cr[i] * cr[j];
          }
          K[i * ld + j] = H[i * ld + j] + (i == j ? ID_QP_RHO : 0.0) + (t0 + t1) * imu;
        }
        SMPC_END
        // right-looking LDL^T, one barrier per column, every thread updating its own pairs of the trailing block:
        //   K(i,j) -= K(i,k) K(j,k) / K(k,k)   (j > k; nobody writes column k during step k)
        // the extra row carries the forward substitution along: K(n,j) ends as z_j of U z = -grad (U unit lower, K = U D U^T)
        for (int k = 0; k < n; k++)
        {
          SMPC_PAR(tid)
          double * K = sm + L.o_K;
          const double dk = K[k * ld + k];
          const double inv = 1.0 / dk;
          if (tid == 0)
          {
            xt[k] = inv; // xt is free during the factorisation
            if (!(dk > 0))
              sc[SC_BAD] = 1.0;
          }
          for (int idx = (k + 1) * (n + 1) - (k + 1) * k / 2 + tid; idx < npair; idx += SMPC_NT)
          {
            const int i = pairs[idx] >> 8, j = pairs[idx] & 255;
            K[i * ld + j] -= K[i * ld + k] * (K[j * ld + k] * inv);
          }
          SMPC_END
        }
        if (sc[SC_BAD] != 0.0)
          break;
        // dx = U^-T D^-1 z : column-oriented back substitution. On the device one warp keeps y in registers (lane l owns rows
        // l and l + 32) and broadcasts x_k by shuffle, so the n dependent steps need no barrier.
        SMPC_PAR(tid)
        const double * K = sm + L.o_K;
#ifdef __CUDACC__
        if (tid < 32)
        {
          const int i0 = tid, i1 = tid + 32;
          const double d0 = (i0 < n) ? xt[i0] : 0.0, d1 = (i1 < n) ? xt[i1] : 0.0;
          double y0 = (i0 < n) ? K[n * ld + i0] * d0 : 0.0, y1 = (i1 < n) ? K[n * ld + i1] * d1 : 0.0;
          for (int k = n - 1; k > 0; k--)
          {
            const double xa = __shfl_sync(0xffffffffu, y0, k & 31), xb = __shfl_sync(0xffffffffu, y1, k & 31);
            const double xk_ = (k < 32) ? xa : xb;
            if (i0 < k)
              y0 -= K[k * ld + i0] * d0 * xk_;
            if (i1 < k)
              y1 -= K[k * ld + i1] * d1 * xk_;
          }
          if (i0 < n)
            dx[i0] = y0;
          if (i1 < n)
            dx[i1] = y1;
        }
#else
        if (tid == 0)
        {
          for (int i = 0; i < n; i++)
            dx[i] = K[n * ld + i] * xt[i];
          for (int k = n - 1; k > 0; k--)
            for (int i = 0; i < k; i++)
              dx[i] -= K[k * ld + i] * xt[i] * dx[k];
        }
#endif
        SMPC_END
        SMPC_PAR(tid)
        double s = 0;
        for (int i = tid; i < n; i += SMPC_NT)
          s += grad[i] * dx[i];
        id_red_put<false>(part, RK_SLOPE, tid, s);
        SMPC_END
        double alpha = 1.0, vt = val;
        const double slope = id_red_get<false>(part, RK_SLOPE);
        for (int ls = 0; ls < 30; ls++)
        {
          SMPC_PAR(tid)
          for (int i = tid; i < n; i += SMPC_NT)
            xt[i] = x[i] + alpha * dx[i];
          SMPC_END
          vt = id_al_value(L, sm, n, me, mi, mu, xt, re2, si2, (ls & 1) ? RK_VAL0 : RK_VAL1);
          if (vt <= val + 1e-4 * alpha * slope + 1e-13 * fabs(val))
            break;
          alpha *= 0.5;
        }
        // accept; a full step that keeps the active set lands on the exact minimiser of the piecewise quadratic
        SMPC_PAR(tid)
        double changed = (alpha == 1.0) ? 0.0 : 1.0;
        for (int r = tid; r < mi; r += SMPC_NT)
        {
          if (((si[r] > 0) != (si2[r] > 0)) || ((si[r] < 0) != (si2[r] < 0)))
            changed = 1.0;
          si[r] = si2[r];
        }
        for (int i = tid; i < n; i += SMPC_NT)
          x[i] = xt[i];
        for (int r = tid; r < me; r += SMPC_NT)
          re[r] = re2[r];
        id_red_put<true>(part, RK_CHANGED, tid, changed);
        SMPC_END
        val = vt;
        if (id_red_get<true>(part, RK_CHANGED) == 0.0)
          break;
      }
      if (sc[SC_BAD] != 0.0)
      {
        status = 2;
        break;
      }
      // multiplier update; relative KKT residuals (hx = H x from the last value call)
      SMPC_PAR(tid)
      for (int r = tid; r < me; r += SMPC_NT)
        y[r] = re[r] * imu;
      for (int r = tid; r < mi; r += SMPC_NT)
        z[r] = si[r] * imu;
      SMPC_END
      SMPC_PAR(tid)
      const double * A = sm + L.o_A;
      const double * Cm = sm + L.o_C;
      double pr = 0, ps = 1.0;
      for (int idx = tid; idx < me + mi; idx += SMPC_NT)
      {
        const double * row = (idx < me) ? A + idx * ld : Cm + (idx - me) * ld;
        double s = 0;
        for (int j = 0; j < n; j++)
          s += row[j] * x[j];
        if (idx < me)
        {
          pr = fmax(pr, fabs(s - sm[L.o_b + idx]));
          ps = fmax(ps, fmax(fabs(s), fabs(sm[L.o_b + idx])));
        }
        else
        {
          const int r = idx - me;
          pr = fmax(pr, fmax(s - sm[L.o_u + r], sm[L.o_l + r] - s));
          ps = fmax(ps, fabs(s));
        }
      }
      id_red_put<true>(part, RK_PR, tid, pr);
      id_red_put<true>(part, RK_PS, tid, ps);
      SMPC_END
      SMPC_PAR(tid)
      const double * A = sm + L.o_A;
      const double * Cm = sm + L.o_C;
      const double * g = sm + L.o_g;
      const double * hx = sm + L.o_hx;
      double dr = 0, ds = 1.0;
      for (int i = tid; i < n; i += SMPC_NT)
      {
        double ay = 0, cz = 0;
        for (int r = 0; r < me; r++)
          ay += A[r * ld + i] * y[r];
        for (int r = 0; r < mi; r++)
          cz += Cm[r * ld + i] * z[r];
        dr = fmax(dr, fabs(g[i] + hx[i] + ay + cz));
        ds = fmax(ds, fmax(fmax(fabs(hx[i]), fabs(g[i])), fmax(fabs(ay), fabs(cz))));
      }
      id_red_put<true>(part, RK_DR, tid, dr);
      id_red_put<true>(part, RK_DS, tid, ds);
      SMPC_END
      pres = id_red_get<true>(part, RK_PR) / id_red_get<true>(part, RK_PS);
      dres = id_red_get<true>(part, RK_DR) / id_red_get<true>(part, RK_DS);
      if (pres <= ID_QP_EPS && dres <= ID_QP_EPS)
      {
        status = 0;
        outer++;
        break;
      }
      mu = fmax(0.1 * mu, ID_QP_MU_MIN);
    }
    SMPC_PAR(tid)
    if (tid == 0)
    {
      info->status = status;
      info->outer_iters = outer;
      info->newton_iters = newton;
      info->nvar = n;
      info->neq = me;
      info->nin = mi;
      info->prim_res = pres;
      info->dual_res = dres;
    }
    if (ws != nullptr)
    { // ws[0] was read by every thread before the first barrier of this routine
      if (tid == 0)
        ws[0] = (status == 0) ? sig : -1.0;
      for (int i = tid; i < n; i += SMPC_NT)
        ws[1 + i] = x[i];
      for (int i = tid; i < me; i += SMPC_NT)
        ws[1 + L.nx + i] = y[i];
      for (int i = tid; i < mi; i += SMPC_NT)
        ws[1 + L.nx + L.me + i] = z[i];
    }
    SMPC_END
  }

  // One instance: q, v (global) -> tau (na), a (nv), f (nfeet x 12), info.
  SMPC_DEV void id_solve_cta(
    const DevModel & Mo, const smpc_id_desc & D, const IdTargetDev & T, const IdDims & L, const double * q_g, const double * v_g,
    double * sm, double * tau_g, double * a_g, double * f_g, smpc_id_info * info, double * ws = nullptr)
  {
    const int nv = L.nv, na = L.na;
    SMPC_PAR(tid)
    for (int i = tid; i < L.nq; i += SMPC_NT)
      sm[L.o_q + i] = q_g[i];
    for (int i = tid; i < nv; i += SMPC_NT)
      sm[L.o_v + i] = v_g[i];
    SMPC_END
    id_dynamics(Mo, L, sm);
    int n, me, mi;
    id_build_qp(Mo, D, T, L, sm, n, me, mi);
    int mask = 0;
    for (int f = 0; f < L.nfeet; f++)
      mask |= (T.active[f] ? 1 : 0) << f;
    id_solve_qp(L, sm, n, me, mi, info, ws, (double)(n | (me << 8) | (mi << 16) | (mask << 24)));
    SMPC_PAR(tid)
    const double * x = sm + L.o_x;
    const double * Mm = sm + L.o_M;
    const double * h = sm + L.o_h;
    const double * JcT = sm + L.o_JcT;
    const int nf = n - nv, ldf = nf > 0 ? nf : 1;
    for (int j = tid; j < na; j += SMPC_NT)
    {
      double s = h[6 + j];
      for (int c = 0; c < nv; c++)
        s += Mm[(6 + j) * nv + c] * x[c];
      for (int c = 0; c < nf; c++)
        s -= JcT[(6 + j) * ldf + c] * x[nv + c];
      tau_g[j] = s;
    }
    for (int i = tid; i < nv; i += SMPC_NT)
      a_g[i] = x[i];
    for (int i = tid; i < L.nfeet * SMPC_ID_MAX_FVAR; i += SMPC_NT)
    {
      const int f = i / SMPC_ID_MAX_FVAR, j = i % SMPC_ID_MAX_FVAR;
      int off = nv;
      for (int k = 0; k < f; k++)
        off += T.active[k] ? (D.foot_type[k] == SMPC_ID_QUAD ? 12 : 3) : 0;
      const int dim = T.active[f] ? (D.foot_type[f] == SMPC_ID_QUAD ? 12 : 3) : 0;
      f_g[i] = (j < dim) ? x[off + j] : 0.0;
    }
    SMPC_END
  }

} // namespace smpc
