// ORACLE — test infrastructure only (see orc_math.hpp header). PARITY UNPINNED.
//
// orc_id.hpp — the whole-body inverse-dynamics QP of the reference's KinodynamicsID / CentroidalID
// (reference src/inverse-dynamics/kinodynamics-id.cpp:6-237, centroidal-id.cpp:6-145), restated from TSID 1.9.0's
// acceleration/force formulation and task definitions [UPSTREAM, not in /root/reference; pixi.lock:174]:
//   tsid::InverseDynamicsFormulationAccForce::computeProblemData   variables x = [a (nv); f (3 per point contact, 12 per quad)]
//     level 0: floating-base rows of  M a + h = S^T tau + Jc^T T f        (6 equalities)
//              contact force cones / normal-force bounds, joint acceleration bounds, actuation bounds
//              contact motion rows when Settings::contact_motion_equality (kinodynamics-id.cpp:149-151)
//     level 1: sum_k w_k |A_k x - b_k|^2 over the motion / force-regularisation tasks (+ 1e-8 |x|^2, tsid solvers/fwd.hpp
//              DEFAULT_HESSIAN_REGULARIZATION)
//   tsid::tasks::TaskSE3Equality (local frame or world-oriented, mask), TaskJointPosture, TaskComEquality,
//   TaskJointPosVelAccBounds (Del Prete, "Joint position and velocity bounds in discrete-time acceleration/torque control
//   of robot manipulators", RA-L 2018), TaskActuationBounds, contacts::ContactPoint / Contact6d.
// The QP itself is strictly convex, so its solution does not depend on the solver: instead of restating ProxQP's iterates
// (proxsuite 0.7.2, un-vendored) it is solved here by a proximal augmented-Lagrangian / semismooth-Newton method to a relative KKT
// residual of 1e-10, far below the 1e-5 ProxQP stops at in TSID's wrapper; the CUDA path is compared at solution level.
// Reference quirks kept: the base task's velocity reference receives the ACCELERATION target and its acceleration
// reference stays zero (kinodynamics-id.cpp:224-225 calls setDerivative twice); contact references are reset to the
// measured foot pose at every solve (:199-216), so the proportional term of the contact tasks is identically zero.
#pragma once
#include "orc_rbd.hpp"
#include "../include/smpc_id.h"
#include <cmath>
#include <stdexcept>

namespace orc
{

  constexpr double ID_HESSIAN_REG = 1e-8; // tsid DEFAULT_HESSIAN_REGULARIZATION [UPSTREAM]
  constexpr double ID_BIG = 1e10;         // tsid's "no bound" value [UPSTREAM]
  constexpr double ID_DDQ_MAX = 1e10;     // TaskJointPosVelAccBounds acceleration bound left at its default (kinodynamics-id.cpp:88-89)
  constexpr double ID_QP_EPS = 1e-10;
  constexpr double ID_QP_RHO = 1e-10; // primal proximal weight (below the Hessian regularisation, so the weakest directions still move)
  constexpr double ID_QP_MU0 = 1e-2;  // initial penalty parameter (1/mu multiplies the constraint violations)
  constexpr double ID_QP_MU_MIN = 1e-4;
  constexpr double ID_QP_MU_WARM = 1e-3; // first penalty parameter of a warm-started solve
  constexpr double ID_QP_GRAD_REL = 1e-13; // inner stop: gradient at the rounding floor of its largest term
  constexpr int ID_QP_MAX_OUTER = 30;
  constexpr int ID_QP_MAX_INNER = 30;
  constexpr int ID_QP_MAX_NEWTON = 200;

  struct IdTarget
  {
    Vec q_post, v_post, a_post; // posture sample (kinodynamics-id.cpp:131-134)
    SE3 base_ref;               // :137
    double base_vel[6], base_acc[6]; // :141-142
    int active[SMPC_MAX_FEET];
    double f_ref[SMPC_MAX_FEET][6];
    double com_ref[3], com_vel[3]; // centroidal-id.cpp:100-103
    int tracked[SMPC_MAX_FEET];
    SE3 foot_ref[SMPC_MAX_FEET];
    double foot_vel[SMPC_MAX_FEET][6];
  };

  struct IdQP
  {
    int n = 0, me = 0, mi = 0;
    Mat H, A, C;
    Vec g, b, l, u;
  };

  struct IdSolution
  {
    Vec x, y, z, tau, a;
    double f[SMPC_MAX_FEET][SMPC_ID_MAX_FVAR];
    smpc_id_info info;
    double sig = -1.0; // layout signature of the stored (x, y, z): a matching solve starts from them (warm start)
  };

  // Joint-space inertia and nonlinear effects (gravity included) from the drift pass (same construction as full_dynamics).
  inline void id_mass_bias(const Model & m, const Kin & k0, const double * gravity, Mat & M, Vec & h)
  {
    const int nv = m.nv;
    const double g6[6] = {gravity[0], gravity[1], gravity[2], 0, 0, 0};
    M.resize(nv, nv);
    h.assign(nv, 0);
    for (int c = 0; c < nv; c++)
    {
      const int i = m.joint_of_dof[c];
      double Jc6[6], F[6], Yg[6];
      for (int r = 0; r < 6; r++)
        Jc6[r] = k0.J(r, c);
      mat6vec(&k0.Y[36 * i], Jc6, F);
      for (int j = 0; j < nv; j++)
        if (m.anc[i][m.joint_of_dof[j]])
        {
          double s = 0;
          for (int r = 0; r < 6; r++)
            s += k0.J(r, j) * F[r];
          M(j, c) = s;
          M(c, j) = s;
        }
      mat6vec(&k0.Y[36 * i], g6, Yg);
      double s = 0;
      for (int r = 0; r < 6; r++)
        s += Jc6[r] * (k0.fs[6 * i + r] - Yg[r]);
      h[c] = s;
    }
  }

  // LOCAL-frame Jacobian (6 x nv), velocity and classical acceleration drift (a = 0) of the frame F carried by joint jf.
  inline void id_frame_local(const Model & m, const Kin & k0, int jf, const SE3 & F, Mat & J, double * vloc, double * drift)
  {
    const int nv = m.nv;
    J.resize(6, nv);
    for (int c = 0; c < nv; c++)
    {
      if (!m.anc[jf][m.joint_of_dof[c]])
        continue;
      double col[6], lin[3], t[3];
      for (int r = 0; r < 6; r++)
        col[r] = k0.J(r, c);
      cross3(col + 3, F.p, t);
      for (int r = 0; r < 3; r++)
        lin[r] = col[r] + t[r];
      double a[3], b[3];
      matTvec3(F.R, lin, a);
      matTvec3(F.R, col + 3, b);
      for (int r = 0; r < 3; r++)
      {
        J(r, c) = a[r];
        J(3 + r, c) = b[r];
      }
    }
    double pd[3], pdd[3];
    point_va(&k0.ov[6 * jf], &k0.oa[6 * jf], F.p, pd, pdd);
    matTvec3(F.R, pd, vloc);
    matTvec3(F.R, &k0.ov[6 * jf + 3], vloc + 3);
    matTvec3(F.R, pdd, drift);
    matTvec3(F.R, &k0.oa[6 * jf + 3], drift + 3);
  }

  // TaskJointPosVelAccBounds::computeAccLimits with (position, velocity, viability) imposed and no acceleration bound
  // (kinodynamics-id.cpp:82-90). The viability roots are evaluated in their cancellation-free form.
  inline void id_acc_bounds(double q, double dq, double qmin, double qmax, double dqmax, double dt, double & lb, double & ub)
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
        lbp = std::max(dq * dq / (2.0 * (q - qmin)), mdq_dt);
      else
        lbp = ID_BIG;
    }
    else
    {
      lbp = min_q3;
      if (max_q3 > mdq_dt)
        ubp = max_q3;
      else if (q != qmax)
        ubp = std::min(-dq * dq / (2.0 * (qmax - q)), mdq_dt);
      else
        ubp = -ID_BIG;
    }
    const double lbv = (-dqmax - dq) / dt, ubv = (dqmax - dq) / dt;
    // viability: largest / smallest acceleration from which the position bound can still be respected braking at ddq_max
    const double dt2 = dt * dt, two_a = 2.0 * dt2, qn = q + dt * dq;
    const double b1 = 2.0 * dt * dq + ID_DDQ_MAX * dt2;
    const double c1 = dq * dq - 2.0 * ID_DDQ_MAX * (qmax - qn);
    const double d1 = b1 * b1 - 2.0 * two_a * c1;
    double ddq1 = mdq_dt;
    if (d1 >= 0.0)
      ddq1 = (b1 >= 0.0) ? (-2.0 * c1) / (b1 + std::sqrt(d1)) : (-b1 + std::sqrt(d1)) / two_a;
    const double b2 = 2.0 * dt * dq - ID_DDQ_MAX * dt2;
    const double c2 = 2.0 * ID_DDQ_MAX * (qn - qmin) - dq * dq;
    const double d2 = b2 * b2 - 2.0 * two_a * c2;
    double ddq2 = mdq_dt;
    if (d2 >= 0.0)
      ddq2 = (b2 <= 0.0) ? (2.0 * c2) / (-b2 + std::sqrt(d2)) : (-b2 - std::sqrt(d2)) / two_a;
    const double ubvia = std::max(ddq1, mdq_dt), lbvia = std::min(ddq2, mdq_dt);
    lb = std::max(std::max(lbp, lbv), std::max(lbvia, -ID_BIG));
    ub = std::min(std::min(ubp, ubv), std::min(ubvia, ID_BIG));
    if (ub < lb)
    { // conflicting bounds: keep the position-derived side that is violated (TSID resolves conflicts in favour of position)
      const double mid = (dq <= 0.0) ? lb : ub;
      lb = ub = mid;
    }
  }

  struct IdWork
  {
    Mat M;
    Vec h;
    Mat JcT; // nv x nf : Jc^T T (generalised force of each force variable)
    int foff[SMPC_MAX_FEET], fdim[SMPC_MAX_FEET];
    int nf;
  };

  // tangent basis of tsid's friction pyramid for the normal (0,0,1): t1 = x, t2 = y
  inline void id_pyramid_rows(double mu, double B[4][3])
  {
    const double r[4][3] = {{-1, 0, -mu}, {1, 0, -mu}, {0, -1, -mu}, {0, 1, -mu}};
    memcpy(B, r, sizeof(r));
  }

  inline void id_add_task(IdQP & qp, double w, const double * Arow, int col0, int ncol, double b)
  { // H += w a a^T, g -= w a b for one task row living on columns [col0, col0 + ncol)
    for (int i = 0; i < ncol; i++)
    {
      const double ai = Arow[i];
      if (ai == 0.0)
        continue;
      for (int j = 0; j < ncol; j++)
        qp.H(col0 + i, col0 + j) += w * ai * Arow[j];
      qp.g[col0 + i] -= w * ai * b;
    }
  }

  inline void id_build_qp(
    const Model & m, const smpc_id_desc & D, const IdTarget & T, const double * q, const double * v, IdQP & qp, IdWork & W)
  {
    const smpc_id_settings & S = D.settings;
    const int nv = m.nv, na = nv - 6, nfeet = m.nfeet;
    const double gravity[3] = {0, 0, -9.81};
    Kin k0;
    centroidal_pass(m, q, v, nullptr, false, k0);
    id_mass_bias(m, k0, gravity, W.M, W.h);
    const double weight = m.total_mass * 9.81;
    const double fmax = S.contact_weight_ratio_max * weight, fmin = S.contact_weight_ratio_min * weight;

    // variable layout
    int nf = 0, ncone = 0, nmot = 0;
    for (int f = 0; f < nfeet; f++)
    {
      W.foff[f] = nv + nf;
      W.fdim[f] = 0;
      if (!T.active[f])
        continue;
      const bool quad = D.foot_type[f] == SMPC_ID_QUAD;
      W.fdim[f] = quad ? 12 : 3;
      nf += W.fdim[f];
      ncone += quad ? 17 : 5;
      nmot += quad ? 6 : 3;
    }
    W.nf = nf;
    const int n = nv + nf;
    qp.n = n;
    qp.me = 6 + (S.contact_motion_equality ? nmot : 0);
    qp.mi = ncone + na + na;
    qp.H.resize(n, n);
    qp.g.assign(n, 0);
    qp.A.resize(qp.me, n);
    qp.b.assign(qp.me, 0);
    qp.C.resize(qp.mi, n);
    qp.l.assign(qp.mi, 0);
    qp.u.assign(qp.mi, 0);
    W.JcT.resize(nv, nf > 0 ? nf : 1);
    for (int i = 0; i < n; i++)
      qp.H(i, i) = ID_HESSIAN_REG;

    // posture task (kinodynamics-id.cpp:60-66, tsid TaskJointPosture)
    if (S.w_posture > 0.)
    {
      const double kp = S.kp_posture, kd = 2.0 * std::sqrt(kp);
      for (int j = 0; j < na; j++)
      {
        const double ades = -kp * (q[7 + j] - T.q_post[j]) - kd * (v[6 + j] - T.v_post[j]) + T.a_post[j];
        const double one = 1.0;
        id_add_task(qp, S.w_posture, &one, 6 + j, 1, ades);
      }
    }
    Mat J;
    double vloc[6], drift[6];
    // base task (kinodynamics-id.cpp:70-76, 218-226; centroidal-id.cpp:11-21: orientation only)
    if (S.w_base > 0.)
    {
      id_frame_local(m, k0, m.d.base_joint, k0.oMbase, J, vloc, drift);
      const double kp = S.kp_base, kd = 2.0 * std::sqrt(kp);
      double e[6];
      log6(se3_mul(se3_inv(k0.oMbase), T.base_ref), e);
      for (int r = (S.centroidal ? 3 : 0); r < 6; r++)
      {
        // v_ref (world-oriented) = R * base_acc target -> local: base_acc ; a_ref = 0 (reference quirk, header)
        const double ades = kp * e[r] + kd * (T.base_acc[r] - vloc[r]);
        id_add_task(qp, S.w_base, J.row(r), 0, nv, ades - drift[r]);
      }
    }
    // centre-of-mass task (centroidal-id.cpp:23-28, tsid TaskComEquality)
    if (S.centroidal && S.w_com > 0.)
    {
      const double kp = S.kp_com, kd = 2.0 * std::sqrt(kp);
      for (int r = 0; r < 3; r++)
      {
        const double ades = -kp * (k0.com[r] - T.com_ref[r]) - kd * (k0.vcom[r] - T.com_vel[r]);
        const double dr = k0.hd0[r] / m.total_mass;
        id_add_task(qp, S.w_com, k0.Jcom.row(r), 0, nv, ades - dr);
      }
    }
    // swing-foot tracking tasks (centroidal-id.cpp:32-55, 106-139)
    if (S.centroidal && S.w_feet_tracking > 0.)
      for (int f = 0; f < nfeet; f++)
      {
        if (!T.tracked[f])
          continue;
        id_frame_local(m, k0, m.d.foot_joint[f], k0.oMf[f], J, vloc, drift);
        const double kp = S.kp_feet_tracking, kd = 2.0 * std::sqrt(kp);
        double e[6], vr[6];
        log6(se3_mul(se3_inv(k0.oMf[f]), T.foot_ref[f]), e);
        matTvec3(k0.oMf[f].R, T.foot_vel[f], vr);
        matTvec3(k0.oMf[f].R, T.foot_vel[f] + 3, vr + 3);
        const int rows = (D.foot_type[f] == SMPC_ID_QUAD) ? 6 : 3;
        for (int r = 0; r < rows; r++)
          id_add_task(qp, S.w_feet_tracking, J.row(r), 0, nv, kp * e[r] + kd * (vr[r] - vloc[r]) - drift[r]);
      }

    // contacts (kinodynamics-id.cpp:24-55, 144-184, 199-216)
    int ie = 6, ii = 0;
    double pyr[4][3];
    id_pyramid_rows(S.friction_coefficient, pyr);
    for (int f = 0; f < nfeet; f++)
    {
      if (!T.active[f])
        continue;
      const bool quad = D.foot_type[f] == SMPC_ID_QUAD;
      const int c0 = W.foff[f], nfv = W.fdim[f];
      id_frame_local(m, k0, m.d.foot_joint[f], k0.oMf[f], J, vloc, drift);
      const double kd = 2.0 * std::sqrt(S.kp_contact);
      const int rows = quad ? 6 : 3;
      Mat Jm(rows, nv);
      double bm[6];
      if (quad)
      { // Contact6d: motion task in the LOCAL frame, reference = measured pose
        for (int r = 0; r < 6; r++)
        {
          for (int c = 0; c < nv; c++)
            Jm(r, c) = J(r, c);
          bm[r] = -kd * vloc[r] - drift[r];
        }
      }
      else
      { // ContactPoint with useLocalFrame(false): world-oriented linear rows
        const double * R = k0.oMf[f].R;
        double vw[3], dw[3];
        matvec3(R, vloc, vw);
        matvec3(R, drift, dw);
        for (int c = 0; c < nv; c++)
        {
          const double jl[3] = {J(0, c), J(1, c), J(2, c)};
          double jw[3];
          matvec3(R, jl, jw);
          for (int r = 0; r < 3; r++)
            Jm(r, c) = jw[r];
        }
        for (int r = 0; r < 3; r++)
          bm[r] = -kd * vw[r] - dw[r];
      }
      // force generator: generalised force of force variable j is Jm^T T(:, j)
      for (int j = 0; j < nfv; j++)
      {
        double Tc[6] = {0, 0, 0, 0, 0, 0};
        if (quad)
        {
          const int pt = j / 3, ax = j % 3;
          double e3[3] = {0, 0, 0};
          e3[ax] = 1;
          double mo[3];
          cross3(D.quad_points[f][pt], e3, mo);
          Tc[ax] = 1;
          for (int r = 0; r < 3; r++)
            Tc[3 + r] = mo[r];
        }
        else
          Tc[j] = 1;
        for (int c = 0; c < nv; c++)
        {
          double s = 0;
          for (int r = 0; r < rows; r++)
            s += Jm(r, c) * Tc[r];
          W.JcT(c, c0 - nv + j) = s;
        }
      }
      // motion rows: cost or equality
      if (S.contact_motion_equality)
      {
        for (int r = 0; r < rows; r++, ie++)
        {
          for (int c = 0; c < nv; c++)
            qp.A(ie, c) = Jm(r, c);
          qp.b[ie] = bm[r];
        }
      }
      else if (S.w_contact_motion > 0.)
        for (int r = 0; r < rows; r++)
          id_add_task(qp, S.w_contact_motion, Jm.row(r), 0, nv, bm[r]);
      // force regularisation (tsid ContactPoint / Contact6d m_weightForceRegTask [UPSTREAM])
      if (S.w_contact_force > 0.)
      {
        const double wreg[6] = {1, 1, 1e-3, 2, 2, 2};
        for (int r = 0; r < rows; r++)
        {
          double row[SMPC_ID_MAX_FVAR];
          for (int j = 0; j < nfv; j++)
          {
            double t = 0;
            if (quad)
            {
              const int pt = j / 3, ax = j % 3;
              if (r < 3)
                t = (r == ax) ? 1.0 : 0.0;
              else
              {
                double e3[3] = {0, 0, 0}, mo[3];
                e3[ax] = 1;
                cross3(D.quad_points[f][pt], e3, mo);
                t = mo[r - 3];
              }
            }
            else
              t = (r == j) ? 1.0 : 0.0;
            row[j] = wreg[r] * t;
          }
          id_add_task(qp, S.w_contact_force, row, c0, nfv, wreg[r] * T.f_ref[f][r]);
        }
      }
      // force inequalities: friction pyramid per contact point, bounds on the (total) normal force
      const int npts = quad ? 4 : 1;
      for (int pt = 0; pt < npts; pt++)
        for (int r = 0; r < 4; r++, ii++)
        {
          for (int ax = 0; ax < 3; ax++)
            qp.C(ii, c0 + 3 * pt + ax) = pyr[r][ax];
          qp.l[ii] = -ID_BIG;
          qp.u[ii] = 0.0;
        }
      for (int pt = 0; pt < npts; pt++)
        qp.C(ii, c0 + 3 * pt + 2) = 1.0;
      qp.l[ii] = fmin;
      qp.u[ii] = fmax;
      ii++;
    }
    // floating-base dynamics: M_u a - (Jc^T T)_u f = -h_u
    for (int r = 0; r < 6; r++)
    {
      for (int c = 0; c < nv; c++)
        qp.A(r, c) = W.M(r, c);
      for (int j = 0; j < nf; j++)
        qp.A(r, nv + j) = -W.JcT(r, j);
      qp.b[r] = -W.h[r];
    }
    // joint acceleration bounds
    for (int j = 0; j < na; j++, ii++)
    {
      double lb, ub;
      id_acc_bounds(q[7 + j], v[6 + j], D.qmin[j], D.qmax[j], D.vmax[j], D.control_dt, lb, ub);
      qp.C(ii, 6 + j) = 1.0;
      qp.l[ii] = lb;
      qp.u[ii] = ub;
    }
    // actuation bounds on tau = M_a a + h_a - (Jc^T T)_a f
    for (int j = 0; j < na; j++, ii++)
    {
      for (int c = 0; c < nv; c++)
        qp.C(ii, c) = W.M(6 + j, c);
      for (int c = 0; c < nf; c++)
        qp.C(ii, nv + c) = -W.JcT(6 + j, c);
      qp.l[ii] = -D.effort[j] - W.h[6 + j];
      qp.u[ii] = D.effort[j] - W.h[6 + j];
    }
    if (ii != qp.mi || ie != qp.me)
      throw std::runtime_error("id_build_qp: row count mismatch");
    // row equilibration (the solver's own preconditioning: unit inf-norm constraint rows; x is unchanged)
    for (int r = 0; r < qp.me; r++)
    {
      double nr = 0;
      for (int c = 0; c < n; c++)
        nr = std::max(nr, std::fabs(qp.A(r, c)));
      if (nr > 0)
      {
        const double sc = 1.0 / nr;
        for (int c = 0; c < n; c++)
          qp.A(r, c) *= sc;
        qp.b[r] *= sc;
      }
    }
    for (int r = 0; r < qp.mi; r++)
    {
      double nr = 0;
      for (int c = 0; c < n; c++)
        nr = std::max(nr, std::fabs(qp.C(r, c)));
      if (nr > 0)
      {
        const double sc = 1.0 / nr;
        for (int c = 0; c < n; c++)
          qp.C(r, c) *= sc;
        qp.l[r] *= sc;
        qp.u[r] *= sc;
      }
    }
  }

  // ---- QP: min 1/2 x'Hx + g'x  s.t.  A x = b,  l <= C x <= u -------------------------------------------------------
  // Proximal method of multipliers; each subproblem (a piecewise-quadratic, strongly convex function of x) by
  // semismooth Newton with backtracking on the subproblem value.
  inline double id_al_value(const IdQP & qp, const Vec & x, const Vec & xk, const Vec & y, const Vec & z, double mu, Vec & re, Vec & si)
  {
    const int n = qp.n;
    double val = 0;
    for (int i = 0; i < n; i++)
    {
      double s = 0;
      for (int j = 0; j < n; j++)
        s += qp.H(i, j) * x[j];
      val += x[i] * (0.5 * s + qp.g[i]) + 0.5 * ID_QP_RHO * (x[i] - xk[i]) * (x[i] - xk[i]);
    }
    for (int r = 0; r < qp.me; r++)
    {
      double s = -qp.b[r] + mu * y[r];
      for (int j = 0; j < n; j++)
        s += qp.A(r, j) * x[j];
      re[r] = s;
      val += 0.5 / mu * s * s;
    }
    for (int r = 0; r < qp.mi; r++)
    {
      double s = mu * z[r];
      for (int j = 0; j < n; j++)
        s += qp.C(r, j) * x[j];
      double vi = 0;
      if (s > qp.u[r])
        vi = s - qp.u[r];
      else if (s < qp.l[r])
        vi = s - qp.l[r];
      si[r] = vi;
      val += 0.5 / mu * vi * vi;
    }
    return val;
  }

  inline void id_solve_qp(const IdQP & qp, IdSolution & sol, double sig = -2.0, bool warm_enabled = false)
  {
    const int n = qp.n, me = qp.me, mi = qp.mi;
    Vec x(n, 0.0), y(me, 0.0), z(mi, 0.0), xk(n), re(me), si(mi), grad(n), dx(n), xt(n), re2(me), si2(mi);
    const bool warm = warm_enabled && sol.sig == sig && (int)sol.x.size() == n && (int)sol.y.size() == me && (int)sol.z.size() == mi;
    if (warm)
    {
      x = sol.x;
      y = sol.y;
      z = sol.z;
    }
    double mu = warm ? ID_QP_MU_WARM : ID_QP_MU0;
    int newton = 0, outer = 0, status = 1;
    double pres = 0, dres = 0;
    Mat K(n, n);
    for (outer = 0; outer < ID_QP_MAX_OUTER; outer++)
    {
      xk = x;
      double val = id_al_value(qp, x, xk, y, z, mu, re, si);
      for (int it = 0; it < ID_QP_MAX_INNER && newton < ID_QP_MAX_NEWTON; it++)
      {
        // gradient of the subproblem and the magnitude of its largest term (the rounding floor scales with it)
        double gn = 0, gs = 0;
        for (int i = 0; i < n; i++)
        {
          double hx = 0;
          for (int j = 0; j < n; j++)
            hx += qp.H(i, j) * x[j];
          double s = qp.g[i] + ID_QP_RHO * (x[i] - xk[i]) + hx, sa = std::fabs(qp.g[i]) + std::fabs(hx);
          for (int r = 0; r < me; r++)
          {
            const double t = qp.A(r, i) * re[r] / mu;
            s += t;
            sa += std::fabs(t);
          }
          for (int r = 0; r < mi; r++)
            if (si[r] != 0.0)
            {
              const double t = qp.C(r, i) * si[r] / mu;
              s += t;
              sa += std::fabs(t);
            }
          grad[i] = s;
          gn = std::max(gn, std::fabs(s));
          gs = std::max(gs, sa);
        }
        if (gn <= ID_QP_GRAD_REL * gs)
          break;
        newton++;
        for (int i = 0; i < n; i++)
          for (int j = 0; j <= i; j++)
          {
            double s = qp.H(i, j) + (i == j ? ID_QP_RHO : 0.0);
            double t = 0;
            for (int r = 0; r < me; r++)
              t += qp.A(r, i) * qp.A(r, j);
            for (int r = 0; r < mi; r++)
              if (si[r] != 0.0)
                t += qp.C(r, i) * qp.C(r, j);
            K(i, j) = s + t / mu;
          }
        if (!fd_chol(K, n))
          throw std::runtime_error("id_solve_qp: Newton matrix not positive definite");
        Mat rhs(n, 1);
        for (int i = 0; i < n; i++)
          rhs(i, 0) = -grad[i];
        fd_chol_solve(K, n, rhs, 1);
        double slope = 0;
        for (int i = 0; i < n; i++)
        {
          dx[i] = rhs(i, 0);
          slope += grad[i] * dx[i];
        }
        double alpha = 1.0, vt = val;
        for (int ls = 0; ls < 30; ls++)
        {
          for (int i = 0; i < n; i++)
            xt[i] = x[i] + alpha * dx[i];
          vt = id_al_value(qp, xt, xk, y, z, mu, re2, si2);
          if (vt <= val + 1e-4 * alpha * slope + 1e-13 * std::fabs(val))
            break;
          alpha *= 0.5;
        }
        // a full step that keeps the active set lands on the exact minimiser of the piecewise quadratic
        bool same = (alpha == 1.0);
        for (int r = 0; r < mi && same; r++)
          same = ((si[r] > 0) == (si2[r] > 0)) && ((si[r] < 0) == (si2[r] < 0));
        x = xt;
        re = re2;
        si = si2;
        val = vt;
        if (same)
          break;
      }
      // multiplier update and relative KKT residuals
      for (int r = 0; r < me; r++)
        y[r] = re[r] / mu;
      for (int r = 0; r < mi; r++)
        z[r] = si[r] / mu;
      pres = 0;
      dres = 0;
      double pscale = 1.0, dscale = 1.0;
      for (int r = 0; r < me; r++)
      {
        double s = 0;
        for (int j = 0; j < n; j++)
          s += qp.A(r, j) * x[j];
        pscale = std::max(pscale, std::max(std::fabs(s), std::fabs(qp.b[r])));
        pres = std::max(pres, std::fabs(s - qp.b[r]));
      }
      for (int r = 0; r < mi; r++)
      {
        double s = 0;
        for (int j = 0; j < n; j++)
          s += qp.C(r, j) * x[j];
        pscale = std::max(pscale, std::fabs(s));
        pres = std::max(pres, std::max(s - qp.u[r], qp.l[r] - s));
      }
      for (int i = 0; i < n; i++)
      {
        double hx = 0, ay = 0, cz = 0;
        for (int j = 0; j < n; j++)
          hx += qp.H(i, j) * x[j];
        for (int r = 0; r < me; r++)
          ay += qp.A(r, i) * y[r];
        for (int r = 0; r < mi; r++)
          cz += qp.C(r, i) * z[r];
        dscale = std::max(dscale, std::max(std::max(std::fabs(hx), std::fabs(qp.g[i])), std::max(std::fabs(ay), std::fabs(cz))));
        dres = std::max(dres, std::fabs(qp.g[i] + hx + ay + cz));
      }
      pres /= pscale;
      dres /= dscale;
      if (pres <= ID_QP_EPS && dres <= ID_QP_EPS)
      {
        status = 0;
        outer++;
        break;
      }
      mu = std::max(0.1 * mu, ID_QP_MU_MIN);
    }
    sol.x = x;
    sol.y = y;
    sol.z = z;
    sol.sig = (status == 0) ? sig : -1.0;
    sol.info.status = status;
    sol.info.outer_iters = outer;
    sol.info.newton_iters = newton;
    sol.info.nvar = n;
    sol.info.neq = me;
    sol.info.nin = mi;
    sol.info.prim_res = pres;
    sol.info.dual_res = dres;
  }

  // KinodynamicsID::solve for one instance.
  inline void id_solve(
    const Model & m, const smpc_id_desc & D, const IdTarget & T, const double * q, const double * v, IdSolution & sol, bool warm_start = false)
  {
    IdQP qp;
    IdWork W;
    id_build_qp(m, D, T, q, v, qp, W);
    int mask = 0;
    for (int f = 0; f < m.nfeet; f++)
      mask |= (T.active[f] ? 1 : 0) << f;
    id_solve_qp(qp, sol, (double)(qp.n | (qp.me << 8) | (qp.mi << 16) | (mask << 24)), warm_start);
    const int nv = m.nv, na = nv - 6;
    sol.a.assign(sol.x.begin(), sol.x.begin() + nv);
    sol.tau.assign(na, 0);
    for (int j = 0; j < na; j++)
    {
      double s = W.h[6 + j];
      for (int c = 0; c < nv; c++)
        s += W.M(6 + j, c) * sol.x[c];
      for (int c = 0; c < W.nf; c++)
        s -= W.JcT(6 + j, c) * sol.x[nv + c];
      sol.tau[j] = s;
    }
    memset(sol.f, 0, sizeof(sol.f));
    for (int f = 0; f < m.nfeet; f++)
      for (int j = 0; j < W.fdim[f]; j++)
        sol.f[f][j] = sol.x[W.foff[f] + j];
  }

  // KinodynamicsID::setTarget (kinodynamics-id.cpp:121-186)
  inline void id_set_target(
    const Model & m, IdTarget & T, const double * q, const double * v, const double * a, const int32_t * contact, const double * f_target)
  {
    const int nv = m.nv, na = nv - 6;
    Kin k;
    fk(m, q, k);
    T.q_post.assign(q + 7, q + 7 + na);
    T.v_post.assign(v + 6, v + 6 + na);
    T.a_post.assign(a + 6, a + 6 + na);
    T.base_ref = k.oMbase;
    for (int r = 0; r < 6; r++)
    {
      T.base_vel[r] = v[r];
      T.base_acc[r] = a[r];
    }
    for (int f = 0; f < m.nfeet; f++)
    {
      T.active[f] = contact[f] ? 1 : 0;
      for (int r = 0; r < 6; r++)
        T.f_ref[f][r] = (contact[f] && f_target) ? f_target[6 * f + r] : 0.0;
    }
  }

} // namespace orc
