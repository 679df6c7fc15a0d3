// ORACLE — test infrastructure only (see orc_math.hpp header). PARITY UNPINNED.
//
// orc_full.hpp — rigid-contact forward dynamics and its analytic derivatives, restated from the published
// algorithms behind pinocchio::constraintDynamics / computeConstraintDynamicsDerivatives (Pinocchio 3.8
// [UPSTREAM, not in /root/reference]; Carpentier, Budhiraja & Mansard, "Proximal and sparse resolution of
// constrained dynamic equations", RSS 2021; Carpentier & Mansard, RSS 2018 for the RNEA derivatives) that
// aligator's MultibodyConstraintFwdDynamics calls. Built at reference src/fulldynamics.cpp:39,50-75,139:
//   * point feet: CONTACT_3D rigid constraint between the foot frame and the world origin, reference frame LOCAL,
//     Baumgarte corrector (Kp, Kd) = settings.Kp_correction / Kd_correction (:70-71)
//   * proximal settings (accuracy 1e-9, mu 1e-10, 10 iterations) (:39)
//
//   M a - Jc^T lam = tau - b(q, v)           M: joint-space inertia, b: nonlinear effects (gravity included)
//   Jc a + gamma(q, v) = 0                   gamma: classical acceleration of the contact point at a = 0, expressed
//                                            in the foot frame, + Kd o (R^T pdot) + Kp o (R^T p)
// solved through the regularised Delassus matrix G = Jc M^-1 Jc^T + mu I with TWO proximal iterations (the second
// removes the O(mu) bias; Pinocchio stops on |dlam| <= accuracy, which the second iteration always meets for a
// full-rank Jc — a fixed count keeps the CUDA path free of data-dependent loops).
// Derivatives: implicit differentiation of the two equations above with the same factorisations,
//   dtau/dq, dtau/dv of RNEA(q, v, a) with the contact forces as external forces held fixed in the foot frame,
//   d(Jc a + gamma)/dq, /dv in closed form (world-frame spatial algebra, see contact_acc_derivs).
//   * quad feet (force_size 6): CONTACT_6D rigid constraint, reference frame LOCAL_WORLD_ALIGNED (reference :56-64):
//       c = [R 0; 0 R] (a_loc + Kd o v_loc + Kp o log6(oMc))        6 rows per foot, lam = [f; tau] in world-aligned
//     axes at the contact-frame origin. a_loc / v_loc: SPATIAL acceleration / velocity of the contact frame in its own
//     axes (no classical term for CONTACT_6D), placement error -log6(c1Mc2) = log6(oMc1) (joint 2 is the universe with
//     an identity placement), the diagonal gains act on the LOCAL components and the result is rotated back
//     (Pinocchio 3.8 constrained-dynamics.hxx [UPSTREAM]). A force held fixed in world-aligned axes does not turn
//     with the foot, which adds one term to the RNEA partials (see `lwa` below).
#pragma once
#include "orc_rbd.hpp"
#include <stdexcept>

namespace orc
{

  constexpr double FULL_PROX_MU = 1e-10; // reference src/fulldynamics.cpp:39 (second argument)
  constexpr int FULL_PROX_ITERS = 2;

  // dense Cholesky A = L L^T of the leading n x n block (row-major, lower part referenced and overwritten)
  inline bool fd_chol(Mat & A, int n)
  {
    for (int j = 0; j < n; j++)
    {
      double d = A(j, j);
      for (int k = 0; k < j; k++)
        d -= A(j, k) * A(j, k);
      if (!(d > 0))
        return false;
      d = std::sqrt(d);
      A(j, j) = d;
      for (int i = j + 1; i < n; i++)
      {
        double s = A(i, j);
        for (int k = 0; k < j; k++)
          s -= A(i, k) * A(j, k);
        A(i, j) = s / d;
      }
    }
    return true;
  }
  // X <- (L L^T)^-1 X for the m columns of X (n x m)
  inline void fd_chol_solve(const Mat & L, int n, Mat & X, int m)
  {
    for (int c = 0; c < m; c++)
    {
      for (int i = 0; i < n; i++)
      {
        double s = X(i, c);
        for (int k = 0; k < i; k++)
          s -= L(i, k) * X(k, c);
        X(i, c) = s / L(i, i);
      }
      for (int i = n - 1; i >= 0; i--)
      {
        double s = X(i, c);
        for (int k = i + 1; k < n; k++)
          s -= L(k, i) * X(k, c);
        X(i, c) = s / L(i, i);
      }
    }
  }

  struct FullDyn
  {
    int cd = 3;                // rows per contact: 3 (CONTACT_3D, LOCAL) or 6 (CONTACT_6D, LOCAL_WORLD_ALIGNED)
    int ncon = 0;              // cd rows per foot in contact
    std::vector<int> feet;     // feet in contact, insertion order (the order of `cms`, reference :116-119)
    Vec a, lam;                // joint acceleration, contact forces (foot frame, environment on robot)
    Mat Jc;                    // ncon x nv
    Vec gamma;                 // ncon
    Mat da_dq, da_dv, da_dtau; // nv x nv
    Mat dl_dq, dl_dv, dl_dtau; // ncon x nv
    Kin k;                     // kinematics pass with the final acceleration (centroidal quantities included)
  };

  // point velocity / classical acceleration of the material point p of a body with world-frame spatial (v, a)
  inline void point_va(const double * v, const double * a, const double * p, double * pd, double * pdd)
  {
    double t[3];
    cross3(v + 3, p, t);
    for (int r = 0; r < 3; r++)
      pd[r] = v[r] + t[r];
    double t1[3], t2[3];
    cross3(a + 3, p, t1);
    cross3(v + 3, pd, t2);
    for (int r = 0; r < 3; r++)
      pdd[r] = a[r] + t1[r] + t2[r];
  }

  // Contact-row value c = R^T pdd + Kd o R^T pd + Kp o R^T p of foot f for body motion (v, a), and (optionally)
  // its partial derivatives with respect to q_m and v_m for one dof m on the support chain:
  //   dv/dq_m = (v_p - v_i) x J_m                    dv/dv_m = J_m
  //   da/dq_m = (a_p - a_i) x J_m + w x (v_i - v_p)   da/dv_m = (v_im + v_p - v_i) x J_m        w = v_p x J_m
  // (i: foot joint, im: joint of dof m, p: parent of im; all world-frame, origin-referenced)
  inline void contact_row(const double * R, const double * p, const double * Kp, const double * Kd, const double * vi, const double * ai, double * c)
  {
    double pd[3], pdd[3], a1[3], a2[3], a3[3];
    point_va(vi, ai, p, pd, pdd);
    matTvec3(R, pdd, a1);
    matTvec3(R, pd, a2);
    matTvec3(R, p, a3);
    for (int r = 0; r < 3; r++)
      c[r] = a1[r] + Kd[r] * a2[r] + Kp[r] * a3[r];
  }
  inline void contact_acc_derivs(
    const double * R, const double * p, const double * Kp, const double * Kd, const double * vi, const double * ai, const double * vim,
    const double * vp, const double * ap, const double * Jm, double * dc_q, double * dc_v)
  {
    double pd[3], pdd[3];
    point_va(vi, ai, p, pd, pdd);
    double dvp[6], dap[6], dvi[6], dv_q[6], w[6], da1[6], da2[6], da_q[6], vs[6], da_v[6];
    for (int r = 0; r < 6; r++)
    {
      dvp[r] = vp[r] - vi[r];
      dap[r] = ap[r] - ai[r];
      dvi[r] = vi[r] - vp[r];
      vs[r] = vim[r] + vp[r] - vi[r];
    }
    mxm(dvp, Jm, dv_q);
    mxm(vp, Jm, w);
    mxm(dap, Jm, da1);
    mxm(w, dvi, da2);
    for (int r = 0; r < 6; r++)
      da_q[r] = da1[r] + da2[r];
    mxm(vs, Jm, da_v);
    double Jp[3], t[3];
    cross3(Jm + 3, p, t);
    for (int r = 0; r < 3; r++)
      Jp[r] = Jm[r] + t[r];
    // dpd/dq, dpdd/dq, dpdd/dv
    double dpd_q[3], dpdd_q[3], dpdd_v[3], t1[3], t2[3], t3[3], t4[3];
    cross3(dv_q + 3, p, t1);
    cross3(vi + 3, Jp, t2);
    for (int r = 0; r < 3; r++)
      dpd_q[r] = dv_q[r] + t1[r] + t2[r];
    cross3(da_q + 3, p, t1);
    cross3(ai + 3, Jp, t2);
    cross3(dv_q + 3, pd, t3);
    cross3(vi + 3, dpd_q, t4);
    for (int r = 0; r < 3; r++)
      dpdd_q[r] = da_q[r] + t1[r] + t2[r] + t3[r] + t4[r];
    cross3(da_v + 3, p, t1);
    cross3(Jm + 3, pd, t2);
    cross3(vi + 3, Jp, t3);
    for (int r = 0; r < 3; r++)
      dpdd_v[r] = da_v[r] + t1[r] + t2[r] + t3[r];
    // rotate into the foot frame; the frame itself turns with J_m.ang
    double wa[3], wv[3], u1[3], u2[3], u3[3], o1[3], o2[3], o3[3], o4[3];
    cross3(Jm + 3, pdd, wa);
    cross3(Jm + 3, pd, wv);
    for (int r = 0; r < 3; r++)
    {
      u1[r] = dpdd_q[r] - wa[r];
      u2[r] = dpd_q[r] - wv[r];
      u3[r] = Jm[r]; // Jp - J_m.ang x p
    }
    matTvec3(R, u1, o1);
    matTvec3(R, u2, o2);
    matTvec3(R, u3, o3);
    for (int r = 0; r < 3; r++)
      dc_q[r] = o1[r] + Kd[r] * o2[r] + Kp[r] * o3[r];
    matTvec3(R, dpdd_v, o1);
    matTvec3(R, Jp, o4);
    for (int r = 0; r < 3; r++)
      dc_v[r] = o1[r] + Kd[r] * o4[r];
  }

  // ---- CONTACT_6D / LOCAL_WORLD_ALIGNED ----------------------------------------------------------------------
  // world-axes motion of the contact-frame origin: [lin + ang x p ; ang]
  inline void frame_motion_lwa(const double * mo, const double * p, double * out)
  {
    double t[3];
    cross3(mo + 3, p, t);
    for (int r = 0; r < 3; r++)
    {
      out[r] = mo[r] + t[r];
      out[3 + r] = mo[3 + r];
    }
  }
  // y = [R 0; 0 R] diag(K) [R^T 0; 0 R^T] x
  inline void rkrt6(const double * R, const double * K, const double * x, double * y)
  {
    double l[6];
    matTvec3(R, x, l);
    matTvec3(R, x + 3, l + 3);
    for (int r = 0; r < 6; r++)
      l[r] *= K[r];
    matvec3(R, l, y);
    matvec3(R, l + 3, y + 3);
  }
  inline void contact_row6(const SE3 & F, const double * Kp, const double * Kd, const double * vi, const double * ai, double * c)
  {
    double V[6], A[6], e[6], ke[6], t1[6], t2[6];
    frame_motion_lwa(vi, F.p, V);
    frame_motion_lwa(ai, F.p, A);
    log6(F, e);
    rkrt6(F.R, Kd, V, t1);
    for (int r = 0; r < 6; r++)
      ke[r] = Kp[r] * e[r];
    matvec3(F.R, ke, t2);
    matvec3(F.R, ke + 3, t2 + 3);
    for (int r = 0; r < 6; r++)
      c[r] = A[r] + t1[r] + t2[r];
  }
  // partial derivatives of contact_row6 with respect to q_m and v_m (dof m on the support chain of the foot joint)
  inline void contact_acc_derivs6(
    const SE3 & F, const double * Kp, const double * Kd, const double * vi, const double * ai, const double * vim, const double * vp,
    const double * ap, const double * Jm, double * dc_q, double * dc_v)
  {
    const double * p = F.p;
    double dvp[6], dap[6], dvi[6], vs[6], dv_q[6], w[6], da1[6], da2[6], da_q[6], da_v[6];
    for (int r = 0; r < 6; r++)
    {
      dvp[r] = vp[r] - vi[r];
      dap[r] = ap[r] - ai[r];
      dvi[r] = vi[r] - vp[r];
      vs[r] = vim[r] + vp[r] - vi[r];
    }
    mxm(dvp, Jm, dv_q); // d v_i / d q_m
    mxm(vp, Jm, w);
    mxm(dap, Jm, da1);
    mxm(w, dvi, da2);
    for (int r = 0; r < 6; r++)
      da_q[r] = da1[r] + da2[r]; // d a_i / d q_m
    mxm(vs, Jm, da_v);           // d a_i / d v_m
    double Jp[6];
    frame_motion_lwa(Jm, p, Jp); // [d p / d q_m ; angular part of J_m]
    const double * wm = Jm + 3;
    double V[6], e[6];
    frame_motion_lwa(vi, p, V);
    log6(F, e);
    // dV/dq, dA/dq: the motion changes and the point moves
    double dV_q[6], dA_q[6], dA_v[6], t[3];
    frame_motion_lwa(dv_q, p, dV_q);
    cross3(vi + 3, Jp, t);
    for (int r = 0; r < 3; r++)
      dV_q[r] += t[r];
    frame_motion_lwa(da_q, p, dA_q);
    cross3(ai + 3, Jp, t);
    for (int r = 0; r < 3; r++)
      dA_q[r] += t[r];
    frame_motion_lwa(da_v, p, dA_v);
    // d/dq [R Kd R^T V] = w_m x (R Kd R^T V) + R Kd R^T (dV - w_m x V), lin and ang blocks alike
    double kv[6], wxV[6], dVr[6], t2[6], c1[6];
    rkrt6(F.R, Kd, V, kv);
    cross3(wm, V, wxV);
    cross3(wm, V + 3, wxV + 3);
    for (int r = 0; r < 6; r++)
      dVr[r] = dV_q[r] - wxV[r];
    rkrt6(F.R, Kd, dVr, t2);
    cross3(wm, kv, c1);
    cross3(wm, kv + 3, c1 + 3);
    // d/dq [R (Kp o e)] = w_m x (R Kp e) + R Kp Jlog6(e) (X_F^-1 J_m)
    double ke[6], rke[6], c2[6], Jl[36], X[36], Jloc[6], de[6], rde[6];
    for (int r = 0; r < 6; r++)
      ke[r] = Kp[r] * e[r];
    matvec3(F.R, ke, rke);
    matvec3(F.R, ke + 3, rke + 3);
    cross3(wm, rke, c2);
    cross3(wm, rke + 3, c2 + 3);
    Jlog6(e, Jl);
    se3_actinv_matrix(F, X);
    mat6vec(X, Jm, Jloc);
    mat6vec(Jl, Jloc, de);
    for (int r = 0; r < 6; r++)
      de[r] *= Kp[r];
    matvec3(F.R, de, rde);
    matvec3(F.R, de + 3, rde + 3);
    for (int r = 0; r < 6; r++)
      dc_q[r] = dA_q[r] + c1[r] + t2[r] + c2[r] + rde[r];
    double t3[6];
    rkrt6(F.R, Kd, Jp, t3); // dV/dv_m = [J_lin + J_ang x p ; J_ang]
    for (int r = 0; r < 6; r++)
      dc_v[r] = dA_v[r] + t3[r];
  }

  // Forward dynamics with rigid contacts on the feet of `mask` (cd = 3: point feet, 6: quad feet); derivatives if `derivs`.
  inline void full_dynamics(
    const Model & m, const double * q, const double * v, const double * tau, unsigned mask, const double * gravity, const double * Kp,
    const double * Kd, bool derivs, FullDyn & o, int cd = 3)
  {
    const smpc_model_desc & d = m.d;
    const int nv = m.nv, nj = m.nj;
    const double g6[6] = {gravity[0], gravity[1], gravity[2], 0, 0, 0};
    o.feet.clear();
    for (int f = 0; f < m.nfeet; f++)
      if (mask & (1u << f))
        o.feet.push_back(f);
    if (cd != 3 && cd != 6)
      throw std::runtime_error("full_dynamics: contact dimension must be 3 or 6");
    const bool lwa = (cd == 6);
    o.cd = cd;
    const int ncon = o.ncon = cd * (int)o.feet.size();

    // pass 1: drift quantities (a = 0)
    Kin k0;
    centroidal_pass(m, q, v, nullptr, false, k0);
    Mat M(nv, nv);
    Vec b(nv);
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
      b[c] = s;
    }
    o.Jc.resize(ncon, nv);
    o.gamma.assign(ncon, 0);
    for (size_t cf = 0; cf < o.feet.size(); cf++)
    {
      const int f = o.feet[cf], jf = d.foot_joint[f];
      const SE3 & F = k0.oMf[f];
      if (lwa)
      {
        contact_row6(F, Kp, Kd, &k0.ov[6 * jf], &k0.oa[6 * jf], &o.gamma[6 * cf]);
        for (int c = 0; c < nv; c++)
        {
          if (!m.anc[jf][m.joint_of_dof[c]])
            continue;
          double Jm[6], Jp[6];
          for (int r = 0; r < 6; r++)
            Jm[r] = k0.J(r, c);
          frame_motion_lwa(Jm, F.p, Jp);
          for (int r = 0; r < 6; r++)
            o.Jc(6 * (int)cf + r, c) = Jp[r];
        }
        continue;
      }
      contact_row(F.R, F.p, Kp, Kd, &k0.ov[6 * jf], &k0.oa[6 * jf], &o.gamma[3 * cf]);
      for (int c = 0; c < nv; c++)
      {
        double c3[3], l3[3];
        point_jac_col(m, k0, jf, F.p, c, c3);
        matTvec3(F.R, c3, l3);
        for (int r = 0; r < 3; r++)
          o.Jc(3 * (int)cf + r, c) = l3[r];
      }
    }
    Mat L = M;
    if (!fd_chol(L, nv))
      throw std::runtime_error("full_dynamics: joint-space inertia not positive definite");
    Mat MiJt(nv, ncon > 0 ? ncon : 1);
    for (int r = 0; r < ncon; r++)
      for (int c = 0; c < nv; c++)
        MiJt(c, r) = o.Jc(r, c);
    if (ncon)
      fd_chol_solve(L, nv, MiJt, ncon);
    Mat G(ncon > 0 ? ncon : 1, ncon > 0 ? ncon : 1);
    for (int r = 0; r < ncon; r++)
      for (int c = 0; c < ncon; c++)
      {
        double s = 0;
        for (int kk = 0; kk < nv; kk++)
          s += o.Jc(r, kk) * MiJt(kk, c);
        G(r, c) = s + (r == c ? FULL_PROX_MU : 0.0);
      }
    Mat LG = G;
    if (ncon && !fd_chol(LG, ncon))
      throw std::runtime_error("full_dynamics: Delassus matrix not positive definite");
    Mat af(nv, 1);
    for (int c = 0; c < nv; c++)
      af(c, 0) = tau[c] - b[c];
    fd_chol_solve(L, nv, af, 1);
    Vec rhs0(ncon);
    for (int r = 0; r < ncon; r++)
    {
      double s = o.gamma[r];
      for (int c = 0; c < nv; c++)
        s += o.Jc(r, c) * af(c, 0);
      rhs0[r] = -s;
    }
    o.lam.assign(ncon, 0.0);
    for (int it = 0; it < FULL_PROX_ITERS && ncon; it++)
    {
      Mat x(ncon, 1);
      for (int r = 0; r < ncon; r++)
        x(r, 0) = rhs0[r] + FULL_PROX_MU * o.lam[r];
      fd_chol_solve(LG, ncon, x, 1);
      for (int r = 0; r < ncon; r++)
        o.lam[r] = x(r, 0);
    }
    o.a.assign(nv, 0);
    for (int c = 0; c < nv; c++)
    {
      double s = af(c, 0);
      for (int r = 0; r < ncon; r++)
        s += MiJt(c, r) * o.lam[r];
      o.a[c] = s;
    }

    // pass 2: with the acceleration (also gives the centroidal quantities the costs need)
    centroidal_pass(m, q, v, o.a.data(), derivs, o.k);
    if (!derivs)
      return;
    const Kin & k = o.k;
    // subtree force with gravity and the contact forces as external forces: F_i = fs_i - Y_i g6 - sum fext
    std::vector<double> Fs(6 * nj);
    for (int i = 0; i < nj; i++)
    {
      double Yg[6];
      mat6vec(&k.Y[36 * i], g6, Yg);
      for (int r = 0; r < 6; r++)
        Fs[6 * i + r] = k.fs[6 * i + r] - Yg[r];
    }
    for (size_t cf = 0; cf < o.feet.size(); cf++)
    {
      const int f = o.feet[cf], jf = d.foot_joint[f];
      double fw[3], nw[3];
      if (lwa)
      {
        for (int r = 0; r < 3; r++)
          fw[r] = o.lam[6 * cf + r];
        cross3(k.oMf[f].p, fw, nw);
        for (int r = 0; r < 3; r++)
          nw[r] += o.lam[6 * cf + 3 + r];
      }
      else
      {
        matvec3(k.oMf[f].R, &o.lam[3 * cf], fw);
        cross3(k.oMf[f].p, fw, nw);
      }
      for (int i = jf; i >= 0; i = d.parent[i])
        for (int r = 0; r < 3; r++)
        {
          Fs[6 * i + r] -= fw[r];
          Fs[6 * i + 3 + r] -= nw[r];
        }
    }
    // RNEA partial derivatives (world-frame form of Carpentier & Mansard 2018)
    Mat dtau_dq(nv, nv), dtau_dv(nv, nv);
    const double zero6[6] = {0, 0, 0, 0, 0, 0};
    for (int c = 0; c < nv; c++)
    {
      const int ic = m.joint_of_dof[c], p = d.parent[ic];
      const double * vp = (p >= 0) ? &k.ov[6 * p] : zero6;
      double apg[6], Jm[6], dVdq[6], dAdq[6], dAdv[6], t[6], vs[6];
      for (int r = 0; r < 6; r++)
      {
        apg[r] = ((p >= 0) ? k.oa[6 * p + r] : 0.0) - g6[r];
        Jm[r] = k.J(r, c);
        vs[r] = k.ov[6 * ic + r] + vp[r];
      }
      mxm(vp, Jm, dVdq);
      mxm(apg, Jm, dAdq);
      mxm(vp, dVdq, t);
      for (int r = 0; r < 6; r++)
        dAdq[r] += t[r];
      mxm(vs, Jm, dAdv);
      for (int j = 0; j < nv; j++)
      {
        const int ij = m.joint_of_dof[j];
        int is; // joint whose subtree quantities apply
        bool transport = false;
        if (m.anc[ic][ij])
        {
          is = ic;
          transport = (ic != ij);
        }
        else if (m.anc[ij][ic])
          is = ij;
        else
          continue;
        double Tq[6], Tv[6], u1[6], u2[6], u3[6];
        mat6vec(&k.Y[36 * is], dAdq, u1);
        mat6vec(&k.dY[36 * is], dVdq, u2);
        mxf(dVdq, &k.hs[6 * is], u3);
        for (int r = 0; r < 6; r++)
          Tq[r] = u1[r] + u2[r] + u3[r];
        if (transport)
        {
          mxf(Jm, &Fs[6 * is], u1);
          for (int r = 0; r < 6; r++)
            Tq[r] += u1[r];
        }
        mat6vec(&k.Y[36 * is], dAdv, u1);
        mat6vec(&k.dY[36 * is], Jm, u2);
        mxf(Jm, &k.hs[6 * is], u3);
        for (int r = 0; r < 6; r++)
          Tv[r] = u1[r] + u2[r] + u3[r];
        double sq = 0, sv = 0;
        for (int r = 0; r < 6; r++)
        {
          sq += k.J(r, j) * Tq[r];
          sv += k.J(r, j) * Tv[r];
        }
        dtau_dq(j, c) = sq;
        dtau_dv(j, c) = sv;
      }
    }
    // `lwa`: the RNEA partials above treat an external force as attached to its body (it turns with the foot). A force
    // held fixed in world-aligned axes does not: per dof c on the support chain, d(wrench)/dq_c lacks the rotation
    // [w_c x f ; p x (w_c x f) + w_c x tau], so  d tau_k / d q_c  gains  J_k . [w_c x f ; p x (w_c x f) + w_c x tau].
    if (lwa)
      for (size_t cf = 0; cf < o.feet.size(); cf++)
      {
        const int f = o.feet[cf], jf = d.foot_joint[f];
        const double * lf = &o.lam[6 * cf];
        for (int c = 0; c < nv; c++)
        {
          if (!m.anc[jf][m.joint_of_dof[c]])
            continue;
          const double wc[3] = {k.J(3, c), k.J(4, c), k.J(5, c)};
          double g[3], pg[3], wt[3];
          cross3(wc, lf, g);
          cross3(k.oMf[f].p, g, pg);
          cross3(wc, lf + 3, wt);
          for (int j = 0; j < nv; j++)
          {
            if (!m.anc[jf][m.joint_of_dof[j]])
              continue;
            double sacc = 0;
            for (int r = 0; r < 3; r++)
              sacc += k.J(r, j) * g[r] + k.J(3 + r, j) * (pg[r] + wt[r]);
            dtau_dq(j, c) += sacc;
          }
        }
      }
    // contact-acceleration partial derivatives
    Mat dc_dq(ncon > 0 ? ncon : 1, nv), dc_dv(ncon > 0 ? ncon : 1, nv);
    for (size_t cf = 0; cf < o.feet.size(); cf++)
    {
      const int f = o.feet[cf], jf = d.foot_joint[f];
      for (int c = 0; c < nv; c++)
      {
        const int ic = m.joint_of_dof[c];
        if (!m.anc[jf][ic])
          continue;
        const int p = d.parent[ic];
        double Jm[6], dq3[3], dv3[3];
        for (int r = 0; r < 6; r++)
          Jm[r] = k.J(r, c);
        if (lwa)
        {
          double dq6[6], dv6[6];
          contact_acc_derivs6(
            k.oMf[f], Kp, Kd, &k.ov[6 * jf], &k.oa[6 * jf], &k.ov[6 * ic], (p >= 0) ? &k.ov[6 * p] : zero6,
            (p >= 0) ? &k.oa[6 * p] : zero6, Jm, dq6, dv6);
          for (int r = 0; r < 6; r++)
          {
            dc_dq(6 * (int)cf + r, c) = dq6[r];
            dc_dv(6 * (int)cf + r, c) = dv6[r];
          }
          continue;
        }
        contact_acc_derivs(
          k.oMf[f].R, k.oMf[f].p, Kp, Kd, &k.ov[6 * jf], &k.oa[6 * jf], &k.ov[6 * ic], (p >= 0) ? &k.ov[6 * p] : zero6,
          (p >= 0) ? &k.oa[6 * p] : zero6, Jm, dq3, dv3);
        for (int r = 0; r < 3; r++)
        {
          dc_dq(3 * (int)cf + r, c) = dq3[r];
          dc_dv(3 * (int)cf + r, c) = dv3[r];
        }
      }
    }
    // implicit differentiation: X1 = M^-1 r1, dlam = G^-1 (r2 - Jc X1), da = X1 + M^-1 Jc^T dlam
    Mat X1(nv, 3 * nv);
    for (int r = 0; r < nv; r++)
    {
      for (int c = 0; c < nv; c++)
      {
        X1(r, c) = -dtau_dq(r, c);
        X1(r, nv + c) = -dtau_dv(r, c);
      }
      X1(r, 2 * nv + r) = 1.0;
    }
    fd_chol_solve(L, nv, X1, 3 * nv);
    Mat DL(ncon > 0 ? ncon : 1, 3 * nv);
    for (int r = 0; r < ncon; r++)
      for (int c = 0; c < 3 * nv; c++)
      {
        double s = 0;
        for (int kk = 0; kk < nv; kk++)
          s += o.Jc(r, kk) * X1(kk, c);
        double r2 = (c < nv) ? -dc_dq(r, c) : (c < 2 * nv ? -dc_dv(r, c - nv) : 0.0);
        DL(r, c) = r2 - s;
      }
    if (ncon)
      fd_chol_solve(LG, ncon, DL, 3 * nv);
    o.da_dq.resize(nv, nv);
    o.da_dv.resize(nv, nv);
    o.da_dtau.resize(nv, nv);
    o.dl_dq.resize(ncon, nv);
    o.dl_dv.resize(ncon, nv);
    o.dl_dtau.resize(ncon, nv);
    for (int r = 0; r < nv; r++)
      for (int c = 0; c < 3 * nv; c++)
      {
        double s = X1(r, c);
        for (int kk = 0; kk < ncon; kk++)
          s += MiJt(r, kk) * DL(kk, c);
        if (c < nv)
          o.da_dq(r, c) = s;
        else if (c < 2 * nv)
          o.da_dv(r, c - nv) = s;
        else
          o.da_dtau(r, c - 2 * nv) = s;
      }
    for (int r = 0; r < ncon; r++)
      for (int c = 0; c < 3 * nv; c++)
      {
        if (c < nv)
          o.dl_dq(r, c) = DL(r, c);
        else if (c < 2 * nv)
          o.dl_dv(r, c - nv) = DL(r, c);
        else
          o.dl_dtau(r, c - 2 * nv) = DL(r, c);
      }
  }

} // namespace orc
