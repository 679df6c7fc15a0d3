// ORACLE — test infrastructure only (see orc_math.hpp header). PARITY UNPINNED.
//
// orc_ocp.hpp — per-knot model evaluation and linearisation:
//   dynamics x+ = f(x,u), every cost (value, gradient, Gauss-Newton Hessian), every constraint
//   (value, Jacobians), in the stage layout the three reference handlers build:
//     KinodynamicsOCP::createStage      reference src/kinodynamics.cpp:40-152, terminal :352-377
//     CentroidalOCP::createStage        reference src/centroidal-dynamics.cpp:39-106, terminal :306-328
//   The arithmetic of the aligator classes those files instantiate (KinodynamicsFwdDynamics,
//   IntegratorSemiImplEuler, CentroidalMomentum(Derivative)Residual, FrameTranslation/Placement/
//   VelocityResidual, Centroidal{Friction,Wrench}ConeResidual, DCMPositionResidual,
//   CentroidalFwdDynamics, IntegratorEuler, ...) is [UPSTREAM] and restated from its published
//   definition.
#pragma once
#include "orc_rbd.hpp"
#include "orc_full.hpp"
#include <stdexcept>
#include <string>

namespace orc
{

  enum CstrType
  {
    CSTR_EQ = 0,
    CSTR_NEG = 1,
    CSTR_BOX = 2
  };

  struct Problem
  {
    Model model;
    smpc_ocp_desc ocp;
    int nx, ndx, nu, nk, fs, N;
    // owned copies of the weights
    Mat w_x, w_u, w_cent, w_centder, w_frame, w_forces, w_com, w_lmom, w_amom, w_lacc, w_aacc;
    Vec qmin, qmax, umin, umax;
    double mass;

    static Mat copy_mat(const double * p, int n)
    {
      Mat m(n, n);
      if (p)
        memcpy(m.d.data(), p, sizeof(double) * n * n);
      return m;
    }
    Problem(const smpc_model_desc & md, const smpc_ocp_desc & od) : model(md), ocp(od)
    {
      nk = model.nfeet;
      fs = ocp.force_size;
      N = ocp.horizon;
      mass = model.total_mass;
      if (fs != 3 && fs != 6)
        throw std::runtime_error("force_size must be 3 or 6");
      if (ocp.kind == SMPC_OCP_KINODYNAMICS)
      {
        nx = model.nq + model.nv;
        ndx = 2 * model.nv;
        nu = model.nv - 6 + fs * nk; // reference src/kinodynamics.cpp:34
      }
      else if (ocp.kind == SMPC_OCP_FULLDYNAMICS)
      {
        nx = model.nq + model.nv;
        ndx = 2 * model.nv;
        nu = model.nv - 6; // reference src/ocp-handler.cpp:18
      }
      else
      {
        nx = ndx = 9;
        nu = fs * nk; // reference src/centroidal-dynamics.cpp:31-32
      }
      const int nj = model.nv - 6;
      if (ocp.kind != SMPC_OCP_CENTROIDAL)
      {
        w_x = copy_mat(ocp.w_x, ndx);
        w_cent = copy_mat(ocp.w_cent, 6);
        w_frame = copy_mat(ocp.w_frame, fs);
        if (ocp.qmin && ocp.qmax)
        {
          qmin.assign(ocp.qmin, ocp.qmin + nj);
          qmax.assign(ocp.qmax, ocp.qmax + nj);
        }
        else if (ocp.kinematics_limits)
          throw std::runtime_error("kinematics_limits set but qmin/qmax missing");
      }
      w_u = copy_mat(ocp.w_u, nu);
      if (ocp.kind == SMPC_OCP_KINODYNAMICS)
        w_centder = copy_mat(ocp.w_centder, 6);
      if (ocp.kind == SMPC_OCP_FULLDYNAMICS)
      {
        w_forces = copy_mat(ocp.w_forces, fs);
        if (ocp.umin && ocp.umax)
        {
          umin.assign(ocp.umin, ocp.umin + nj);
          umax.assign(ocp.umax, ocp.umax + nj);
        }
      }
      if (ocp.kind == SMPC_OCP_CENTROIDAL)
      {
        w_com = copy_mat(ocp.w_com, 3);
        w_lmom = copy_mat(ocp.w_linear_mom, 3);
        w_amom = copy_mat(ocp.w_angular_mom, 3);
        w_lacc = copy_mat(ocp.w_linear_acc, 3);
        w_aacc = copy_mat(ocp.w_angular_acc, 3);
      }
      // the pointers in `ocp` are not owned: null them so nobody dereferences stale memory
      ocp.w_x = ocp.w_u = ocp.w_cent = ocp.w_centder = ocp.w_frame = ocp.w_forces = nullptr;
      ocp.w_com = ocp.w_linear_mom = ocp.w_angular_mom = ocp.w_linear_acc = ocp.w_angular_acc = nullptr;
      ocp.qmin = ocp.qmax = ocp.umin = ocp.umax = nullptr;
    }
    // number of constraint rows of a stage with this slot (order documented in stage_eval)
    int stage_nc(const smpc_slot & s) const
    {
      int nc = 0;
      if (ocp.kind == SMPC_OCP_KINODYNAMICS)
      {
        if (ocp.kinematics_limits)
          nc += model.nv - 6;
        for (int f = 0; f < nk; f++)
          if (s.contact_mask & (1u << f))
          {
            if (ocp.force_cone)
              nc += (fs == 6) ? 17 : 2;
            nc += (fs == 6) ? 6 : 3;
            if (fs == 3 && ocp.land_cstr && (s.land_mask & (1u << f)))
              nc += 1;
          }
      }
      else if (ocp.kind == SMPC_OCP_CENTROIDAL)
      {
        for (int f = 0; f < nk; f++)
          if (s.contact_mask & (1u << f))
            nc += (fs == 6) ? 17 : 2;
      }
      else
      {
        // reference src/fulldynamics.cpp:145-211
        if (ocp.torque_limits)
          nc += model.nv - 6;
        if (ocp.kinematics_limits)
          nc += model.nv - 6;
        for (int f = 0; f < nk; f++)
          if (s.contact_mask & (1u << f))
          {
            if (ocp.force_cone)
              nc += (fs == 6) ? 17 : 5; // MultibodyWrenchCone / 5-row friction pyramid
            if (ocp.land_cstr && (s.land_mask & (1u << f)))
              nc += (fs == 6) ? 6 : 4; // 6-D LWA velocity / 3 velocity rows + z-height
          }
      }
      return nc;
    }
    int max_nc() const
    {
      smpc_slot s;
      memset(&s, 0, sizeof(s));
      s.contact_mask = s.land_mask = (1u << nk) - 1;
      return stage_nc(s);
    }
    int term_nc() const
    {
      if (ocp.kind == SMPC_OCP_CENTROIDAL)
        return 0; // reference src/centroidal-dynamics.cpp:318-328 (disabled)
      return ocp.terminal_constraint ? 3 : 0;
    }
  };

  struct StageEval
  {
    int nc = 0;
    Vec xnext, xdot;
    Vec fc; // contact force of every foot in its contact frame, zero for swing feet (full dynamics; MPC::getContactForces)
    double cost = 0;
    Vec c;
    std::vector<int> ctype;
    Vec lo, hi;
    // derivatives
    Mat A, B;
    Vec lx, lu;
    Mat lxx, lxu, luu;
    Mat Cx, Cu;
  };

  // ---- helpers -------------------------------------------------------------------------------
  // cost += 1/2 r^T W r ; lx += Jx^T W r ; lxx += Jx^T W Jx ... for a residual with Jacobians wrt (x,u)
  inline void add_quad_cost(const Vec & r, const Mat & W, const Mat * Jx, const Mat * Ju, bool derivs, StageEval & o)
  {
    const int nr = (int)r.size();
    Vec Wr(nr, 0.0);
    for (int i = 0; i < nr; i++)
    {
      double s = 0;
      for (int j = 0; j < nr; j++)
        s += W(i, j) * r[j];
      Wr[i] = s;
    }
    double cst = 0;
    for (int i = 0; i < nr; i++)
      cst += r[i] * Wr[i];
    o.cost += 0.5 * cst;
    if (!derivs)
      return;
    const int ndx = Jx ? Jx->c : 0, nu = Ju ? Ju->c : 0;
    Mat WJx(nr, ndx), WJu(nr, nu);
    for (int i = 0; i < nr; i++)
    {
      for (int k = 0; k < nr; k++)
      {
        double w = W(i, k);
        if (w == 0.0)
          continue;
        for (int j = 0; j < ndx; j++)
          WJx(i, j) += w * (*Jx)(k, j);
        for (int j = 0; j < nu; j++)
          WJu(i, j) += w * (*Ju)(k, j);
      }
    }
    // J^T (W J): (a, i, b) loop nest - contiguous inner loop, every entry still accumulates over ascending i;
    // rows of J^T that are structurally zero for this residual are skipped (adding 0 * x changes nothing)
    for (int a = 0; a < ndx; a++)
    {
      double g = 0;
      for (int i = 0; i < nr; i++)
        g += (*Jx)(i, a) * Wr[i];
      o.lx[a] += g;
      Vec rxx(ndx, 0.0), rxu(nu, 0.0);
      bool any = false;
      for (int i = 0; i < nr; i++)
      {
        const double jia = (*Jx)(i, a);
        if (jia == 0.0)
          continue;
        any = true;
        const double * wx = WJx.row(i);
        for (int b2 = 0; b2 < ndx; b2++)
          rxx[b2] += jia * wx[b2];
        if (nu)
        {
          const double * wu = WJu.row(i);
          for (int b2 = 0; b2 < nu; b2++)
            rxu[b2] += jia * wu[b2];
        }
      }
      if (!any)
        continue;
      for (int b2 = 0; b2 < ndx; b2++)
        o.lxx(a, b2) += rxx[b2];
      for (int b2 = 0; b2 < nu; b2++)
        o.lxu(a, b2) += rxu[b2];
    }
    for (int a = 0; a < nu; a++)
    {
      double g = 0;
      for (int i = 0; i < nr; i++)
        g += (*Ju)(i, a) * Wr[i];
      o.lu[a] += g;
      Vec ruu(nu, 0.0);
      bool any = false;
      for (int i = 0; i < nr; i++)
      {
        const double jia = (*Ju)(i, a);
        if (jia == 0.0)
          continue;
        any = true;
        const double * wu = WJu.row(i);
        for (int b2 = 0; b2 < nu; b2++)
          ruu[b2] += jia * wu[b2];
      }
      if (any)
        for (int b2 = 0; b2 < nu; b2++)
          o.luu(a, b2) += ruu[b2];
    }
  }

  // 6x6 LU with partial pivoting; solves A X = B for nrhs columns (B row-major 6 x nrhs, in place).
  inline void lu6_solve(const double * Ain, double * Bm, int nrhs)
  {
    double A[36];
    memcpy(A, Ain, sizeof(A));
    int piv[6];
    for (int k = 0; k < 6; k++)
    {
      int p = k;
      double best = std::fabs(A[6 * k + k]);
      for (int i = k + 1; i < 6; i++)
        if (std::fabs(A[6 * i + k]) > best)
        {
          best = std::fabs(A[6 * i + k]);
          p = i;
        }
      piv[k] = p;
      if (p != k)
      {
        for (int j = 0; j < 6; j++)
          std::swap(A[6 * k + j], A[6 * p + j]);
        for (int j = 0; j < nrhs; j++)
          std::swap(Bm[k * nrhs + j], Bm[p * nrhs + j]);
      }
      double inv = 1.0 / A[6 * k + k];
      for (int i = k + 1; i < 6; i++)
      {
        double l = A[6 * i + k] * inv;
        A[6 * i + k] = l;
        for (int j = k + 1; j < 6; j++)
          A[6 * i + j] -= l * A[6 * k + j];
        for (int j = 0; j < nrhs; j++)
          Bm[i * nrhs + j] -= l * Bm[k * nrhs + j];
      }
    }
    for (int k = 5; k >= 0; k--)
    {
      double inv = 1.0 / A[6 * k + k];
      for (int j = 0; j < nrhs; j++)
      {
        double s = Bm[k * nrhs + j];
        for (int i = k + 1; i < 6; i++)
          s -= A[6 * k + i] * Bm[i * nrhs + j];
        Bm[k * nrhs + j] = s * inv;
      }
    }
    (void)piv;
  }

  // 17 x 6 wrench-cone matrix (Caron et al. 2015, as in aligator CentroidalWrenchConeResidual [UPSTREAM]);
  // L, W are the half sizes of the rectangular sole.
  inline void wrench_cone_matrix(double mu, double L, double W, double * A /*17x6*/)
  {
    const double rows[17][6] = {
      {0, 0, -1, 0, 0, 0},
      {-1, 0, -mu, 0, 0, 0},
      {1, 0, -mu, 0, 0, 0},
      {0, -1, -mu, 0, 0, 0},
      {0, 1, -mu, 0, 0, 0},
      {0, 0, -W, -1, 0, 0},
      {0, 0, -W, 1, 0, 0},
      {0, 0, -L, 0, -1, 0},
      {0, 0, -L, 0, 1, 0},
      {-W, -L, -(L + W) * mu, mu, mu, -1},
      {-W, L, -(L + W) * mu, mu, -mu, -1},
      {W, -L, -(L + W) * mu, -mu, mu, -1},
      {W, L, -(L + W) * mu, -mu, -mu, -1},
      {W, L, -(L + W) * mu, mu, mu, 1},
      {W, -L, -(L + W) * mu, mu, -mu, 1},
      {-W, L, -(L + W) * mu, -mu, mu, 1},
      {-W, -L, -(L + W) * mu, -mu, -mu, 1}};
    memcpy(A, rows, sizeof(rows));
  }

  inline void alloc_eval(const Problem & P, int nc, bool derivs, StageEval & o)
  {
    o.nc = nc;
    o.xnext.assign(P.nx, 0);
    o.xdot.assign(P.ndx, 0);
    o.cost = 0;
    o.c.assign(nc, 0);
    o.ctype.assign(nc, CSTR_EQ);
    o.lo.assign(nc, 0);
    o.hi.assign(nc, 0);
    if (derivs)
    {
      o.A.resize(P.ndx, P.ndx);
      o.B.resize(P.ndx, P.nu);
      o.lx.assign(P.ndx, 0);
      o.lu.assign(P.nu, 0);
      o.lxx.resize(P.ndx, P.ndx);
      o.lxu.resize(P.ndx, P.nu);
      o.luu.resize(P.nu, P.nu);
      o.Cx.resize(nc, P.ndx);
      o.Cu.resize(nc, P.nu);
    }
  }

  // state error x (-) xref and its Jacobian wrt x (Pinocchio dDifference ARG1 on the free-flyer block)
  inline void state_error(const Problem & P, const double * x, const double * xref, Vec & r, Mat * J)
  {
    const Model & m = P.model;
    const int nv = m.nv, nq = m.nq;
    r.assign(2 * nv, 0);
    difference_q(m, xref, x, r.data());
    for (int i = 0; i < nv; i++)
      r[nv + i] = x[nq + i] - xref[nq + i];
    if (J)
    {
      J->resize(2 * nv, 2 * nv);
      double Jl[36];
      Jlog6(r.data(), Jl);
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++)
          (*J)(i, j) = Jl[6 * i + j];
      for (int i = 6; i < 2 * nv; i++)
        (*J)(i, i) = 1.0;
    }
  }

  // <foot>_pose_cost of a quad foot: FramePlacementResidual log6(ref^-1 oMf) with weight w_frame (6 x 6),
  // J = Jlog6(r) * (LOCAL frame Jacobian)   (reference src/kinodynamics.cpp:67-73, src/fulldynamics.cpp:99-104)
  inline void frame_placement_cost(const Problem & P, const smpc_slot & s, const Kin & k, int f, bool derivs, StageEval & o)
  {
    const Model & m = P.model;
    const int nv = m.nv;
    SE3 Mref;
    memcpy(Mref.R, s.foot_ref[f], sizeof(Mref.R));
    memcpy(Mref.p, s.foot_ref[f] + 9, sizeof(Mref.p));
    SE3 rMf = se3_mul(se3_inv(Mref), k.oMf[f]);
    Vec r(6);
    log6(rMf, r.data());
    Mat Jx;
    if (derivs)
    {
      Jx.resize(6, P.ndx);
      double Jl[36], X[36];
      Jlog6(r.data(), Jl);
      se3_actinv_matrix(k.oMf[f], X);
      const int jf = m.d.foot_joint[f];
      for (int j = 0; j < nv; j++)
      {
        if (!m.anc[jf][m.joint_of_dof[j]])
          continue;
        double Jc[6], loc[6], out[6];
        for (int rr = 0; rr < 6; rr++)
          Jc[rr] = k.J(rr, j);
        mat6vec(X, Jc, loc);
        mat6vec(Jl, loc, out);
        for (int rr = 0; rr < 6; rr++)
          Jx(rr, j) = out[rr];
      }
    }
    add_quad_cost(r, P.w_frame, &Jx, nullptr, derivs, o);
  }

  // ---------------------------------------------------------------------------------------------
  // Kinodynamics stage (reference src/kinodynamics.cpp:40-152)
  //   u = [f_0 .. f_{nk-1} (fs each) | a_joint (nv-6)]
  //   hdot_F = [sum f + m g ; sum (p_i - c) x f_i (+ tau_i)]      over feet in contact
  //   a_base = Ag[:, :6]^-1 (hdot_F - dAg v - Ag[:, 6:] a_joint)   (KinodynamicsFwdDynamics [UPSTREAM])
  //   semi-implicit Euler: v+ = v + dt a ; q+ = q (+) dt v+         (IntegratorSemiImplEuler [UPSTREAM])
  // Constraint rows, in the reference's insertion order:
  //   [joint-limit box on q_joints]  then per foot IN CONTACT: [cone <= 0] , LOCAL frame velocity = 0
  //   (3 or 6 rows), [z-height = ref on a landing edge (fs = 3 only)].
  // ---------------------------------------------------------------------------------------------
  inline void kino_stage_eval(const Problem & P, const smpc_slot & s, const double * x, const double * u, bool derivs, StageEval & o)
  {
    const Model & m = P.model;
    const int nv = m.nv, nq = m.nq, ndx = P.ndx, nu = P.nu, nk = P.nk, fs = P.fs;
    const int nj = nv - 6, nf = fs * nk;
    const double dt = P.ocp.timestep;
    const double * q = x;
    const double * v = x + nq;
    const double * aj = u + nf;
    alloc_eval(P, P.stage_nc(s), derivs, o);

    // pass 1: bias (a = 0)
    Kin k;
    centroidal_pass(m, q, v, nullptr, false, k);
    double hdF[6] = {P.mass * P.ocp.gravity[0], P.mass * P.ocp.gravity[1], P.mass * P.ocp.gravity[2], 0, 0, 0};
    for (int f = 0; f < nk; f++)
      if (s.contact_mask & (1u << f))
      {
        const double * ff = u + fs * f;
        double lever[3], t[3];
        for (int r = 0; r < 3; r++)
          lever[r] = k.oMf[f].p[r] - k.com[r];
        cross3(lever, ff, t);
        for (int r = 0; r < 3; r++)
        {
          hdF[r] += ff[r];
          hdF[3 + r] += t[r];
          if (fs == 6)
            hdF[3 + r] += ff[3 + r];
        }
      }
    double Agb[36], rhs[6];
    for (int r = 0; r < 6; r++)
    {
      double sacc = hdF[r] - k.hdg[r];
      for (int j = 0; j < nj; j++)
        sacc -= k.Ag(r, 6 + j) * aj[j];
      rhs[r] = sacc;
      for (int c = 0; c < 6; c++)
        Agb[6 * r + c] = k.Ag(r, c);
    }
    lu6_solve(Agb, rhs, 1);
    Vec a(nv);
    for (int r = 0; r < 6; r++)
      a[r] = rhs[r];
    for (int j = 0; j < nj; j++)
      a[6 + j] = aj[j];
    for (int i = 0; i < nv; i++)
    {
      o.xdot[i] = v[i];
      o.xdot[nv + i] = a[i];
    }
    // integrator
    Vec vn(nv), dq(nv);
    for (int i = 0; i < nv; i++)
    {
      vn[i] = v[i] + dt * a[i];
      dq[i] = dt * vn[i];
    }
    integrate_q(m, q, dq.data(), o.xnext.data());
    for (int i = 0; i < nv; i++)
      o.xnext[nq + i] = vn[i];

    // pass 2: with the acceleration, for the derivatives
    if (derivs)
      centroidal_pass(m, q, v, a.data(), true, k);

    // foot point Jacobians (LOCAL_WORLD_ALIGNED translation) 3 x nv each
    std::vector<Mat> Jp(nk);
    if (derivs)
      for (int f = 0; f < nk; f++)
      {
        Jp[f].resize(3, nv);
        for (int j = 0; j < nv; j++)
        {
          double c3[3];
          point_jac_col(m, k, m.d.foot_joint[f], k.oMf[f].p, j, c3);
          for (int r = 0; r < 3; r++)
            Jp[f](r, j) = c3[r];
        }
      }

    // d hdot_F / dq (6 x nv) and / du (6 x nu)
    Mat dhF_dq(6, nv), dhF_du(6, nu);
    if (derivs)
    {
      for (int f = 0; f < nk; f++)
        if (s.contact_mask & (1u << f))
        {
          const double * ff = u + fs * f;
          for (int j = 0; j < nv; j++)
          {
            double dl[3] = {Jp[f](0, j) - k.Jcom(0, j), Jp[f](1, j) - k.Jcom(1, j), Jp[f](2, j) - k.Jcom(2, j)};
            double t[3];
            cross3(dl, ff, t);
            for (int r = 0; r < 3; r++)
              dhF_dq(3 + r, j) += t[r];
          }
          double lever[3], Lx[9];
          for (int r = 0; r < 3; r++)
            lever[r] = k.oMf[f].p[r] - k.com[r];
          skew3(lever, Lx);
          for (int r = 0; r < 3; r++)
          {
            dhF_du(r, fs * f + r) = 1.0;
            for (int c = 0; c < 3; c++)
              dhF_du(3 + r, fs * f + c) = Lx[3 * r + c];
            if (fs == 6)
              dhF_du(3 + r, fs * f + 3 + r) = 1.0;
          }
        }
      // implicit differentiation of Agb a_b + Agj a_j + dAg v = hdot_F
      Mat da_dq(6, nv), da_dv(6, nv), da_du(6, nu);
      {
        Mat R(6, 2 * nv + nu);
        for (int r = 0; r < 6; r++)
        {
          for (int j = 0; j < nv; j++)
          {
            R(r, j) = dhF_dq(r, j) - k.dhd_dq(r, j);
            R(r, nv + j) = -k.dhd_dv(r, j);
          }
          for (int j = 0; j < nf; j++)
            R(r, 2 * nv + j) = dhF_du(r, j);
          for (int j = 0; j < nj; j++)
            R(r, 2 * nv + nf + j) = -k.Ag(r, 6 + j);
        }
        lu6_solve(Agb, R.d.data(), 2 * nv + nu);
        for (int r = 0; r < 6; r++)
        {
          for (int j = 0; j < nv; j++)
          {
            da_dq(r, j) = R(r, j);
            da_dv(r, j) = R(r, nv + j);
          }
          for (int j = 0; j < nu; j++)
            da_du(r, j) = R(r, 2 * nv + j);
        }
      }
      // continuous-time Jacobians of a (nv x .) : rows 0..5 from above, joint rows select a_joint
      // semi-implicit Euler:
      //   v+ = v + dt a            dv+/dx = [dt da_dq, I + dt da_dv]        dv+/du = dt da_du
      //   q+ = q (+) dt v+         dq+/dx = [Jq, 0] + dt Jv dv+/dx          dq+/du = dt Jv dv+/du
      // with Jq = Ad(exp6(dq_b))^-1 (+) I, Jv = Jexp6(dq_b) (+) I on the free-flyer block.
      double Jq6[36], Jv6[36];
      {
        SE3 E = exp6(dq.data());
        se3_actinv_matrix(E, Jq6);
        Jexp6(dq.data(), Jv6);
      }
      Mat dvn_dx(nv, ndx), dvn_du(nv, nu);
      for (int r = 0; r < 6; r++)
      {
        for (int j = 0; j < nv; j++)
        {
          dvn_dx(r, j) = dt * da_dq(r, j);
          dvn_dx(r, nv + j) = dt * da_dv(r, j);
        }
        for (int j = 0; j < nu; j++)
          dvn_du(r, j) = dt * da_du(r, j);
      }
      for (int j = 0; j < nj; j++)
        dvn_du(6 + j, nf + j) = dt;
      for (int i = 0; i < nv; i++)
        dvn_dx(i, nv + i) += 1.0;
      for (int i = 0; i < nv; i++)
      {
        for (int j = 0; j < ndx; j++)
        {
          o.A(nv + i, j) = dvn_dx(i, j);
          double sacc = 0;
          if (i < 6)
          {
            for (int c = 0; c < 6; c++)
              sacc += Jv6[6 * i + c] * dvn_dx(c, j);
          }
          else
            sacc = dvn_dx(i, j);
          o.A(i, j) = dt * sacc;
        }
        for (int j = 0; j < nu; j++)
        {
          o.B(nv + i, j) = dvn_du(i, j);
          double sacc = 0;
          if (i < 6)
          {
            for (int c = 0; c < 6; c++)
              sacc += Jv6[6 * i + c] * dvn_du(c, j);
          }
          else
            sacc = dvn_du(i, j);
          o.B(i, j) = dt * sacc;
        }
      }
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++)
          o.A(i, j) += Jq6[6 * i + j];
      for (int i = 6; i < nv; i++)
        o.A(i, i) += 1.0;
    }

    // ---------------- costs (reference src/kinodynamics.cpp:56-83) ----------------
    {
      // state_cost
      Vec r;
      Mat J;
      state_error(P, x, s.x_ref, r, derivs ? &J : nullptr);
      add_quad_cost(r, P.w_x, &J, nullptr, derivs, o);
    }
    {
      // control_cost
      Vec r(nu);
      for (int i = 0; i < nu; i++)
        r[i] = u[i] - s.u_ref[i];
      Mat Ju;
      if (derivs)
      {
        Ju.resize(nu, nu);
        for (int i = 0; i < nu; i++)
          Ju(i, i) = 1.0;
      }
      add_quad_cost(r, P.w_u, nullptr, &Ju, derivs, o);
    }
    {
      // centroidal_cost: h(q,v) - 0
      Vec r(k.hg, k.hg + 6);
      Mat Jx;
      if (derivs)
      {
        Jx.resize(6, ndx);
        for (int rr = 0; rr < 6; rr++)
          for (int j = 0; j < nv; j++)
          {
            Jx(rr, j) = k.dh_dq(rr, j);
            Jx(rr, nv + j) = k.Ag(rr, j);
          }
      }
      add_quad_cost(r, P.w_cent, &Jx, nullptr, derivs, o);
    }
    {
      // centroidal_derivative_cost: hdot_F(q,u)
      Vec r(hdF, hdF + 6);
      Mat Jx, Ju;
      if (derivs)
      {
        Jx.resize(6, ndx);
        for (int rr = 0; rr < 6; rr++)
          for (int j = 0; j < nv; j++)
            Jx(rr, j) = dhF_dq(rr, j);
        Ju = dhF_du;
      }
      add_quad_cost(r, P.w_centder, &Jx, &Ju, derivs, o);
    }
    for (int f = 0; f < nk; f++)
    {
      // <foot>_pose_cost
      if (fs == 3)
      {
        Vec r(3);
        for (int rr = 0; rr < 3; rr++)
          r[rr] = k.oMf[f].p[rr] - s.foot_ref[f][9 + rr];
        Mat Jx;
        if (derivs)
        {
          Jx.resize(3, ndx);
          for (int rr = 0; rr < 3; rr++)
            for (int j = 0; j < nv; j++)
              Jx(rr, j) = Jp[f](rr, j);
        }
        add_quad_cost(r, P.w_frame, &Jx, nullptr, derivs, o);
      }
      else
      {
        // FramePlacementResidual: log6(ref^-1 oMf), J = Jlog6 * (LOCAL frame Jacobian)
        frame_placement_cost(P, s, k, f, derivs, o);
      }
    }

    // ---------------- constraints (reference src/kinodynamics.cpp:91-149) ----------------
    int row = 0;
    if (P.ocp.kinematics_limits)
    {
      for (int j = 0; j < nj; j++, row++)
      {
        o.c[row] = q[7 + j];
        o.ctype[row] = CSTR_BOX;
        o.lo[row] = P.qmin[j];
        o.hi[row] = P.qmax[j];
        if (derivs)
          o.Cx(row, 6 + j) = 1.0;
      }
    }
    for (int f = 0; f < nk; f++)
    {
      if (!(s.contact_mask & (1u << f)))
        continue;
      const double * ff = u + fs * f;
      if (P.ocp.force_cone)
      {
        if (fs == 6)
        {
          double Ac[17 * 6];
          wrench_cone_matrix(P.ocp.mu, P.ocp.Lfoot, P.ocp.Wfoot, Ac);
          for (int rr = 0; rr < 17; rr++, row++)
          {
            double sacc = 0;
            for (int c = 0; c < 6; c++)
              sacc += Ac[6 * rr + c] * ff[c];
            o.c[row] = sacc;
            o.ctype[row] = CSTR_NEG;
            if (derivs)
              for (int c = 0; c < 6; c++)
                o.Cu(row, fs * f + c) = Ac[6 * rr + c];
          }
        }
        else
        {
          const double mu = P.ocp.mu, eps = 1e-4; // reference src/kinodynamics.cpp:127
          o.c[row] = -ff[2] + eps;
          o.ctype[row] = CSTR_NEG;
          o.c[row + 1] = -mu * mu * ff[2] * ff[2] + ff[0] * ff[0] + ff[1] * ff[1];
          o.ctype[row + 1] = CSTR_NEG;
          if (derivs)
          {
            o.Cu(row, fs * f + 2) = -1.0;
            o.Cu(row + 1, fs * f + 0) = 2 * ff[0];
            o.Cu(row + 1, fs * f + 1) = 2 * ff[1];
            o.Cu(row + 1, fs * f + 2) = -2 * mu * mu * ff[2];
          }
          row += 2;
        }
      }
      {
        double vf[6];
        Mat dq_, dv_;
        frame_velocity_local(m, k, f, vf, derivs ? &dq_ : nullptr, derivs ? &dv_ : nullptr);
        const int nr = (fs == 6) ? 6 : 3;
        for (int rr = 0; rr < nr; rr++, row++)
        {
          o.c[row] = vf[rr];
          o.ctype[row] = CSTR_EQ;
          if (derivs)
            for (int j = 0; j < nv; j++)
            {
              o.Cx(row, j) = dq_(rr, j);
              o.Cx(row, nv + j) = dv_(rr, j);
            }
        }
      }
      if (fs == 3 && P.ocp.land_cstr && (s.land_mask & (1u << f)))
      {
        o.c[row] = k.oMf[f].p[2] - s.land_z[f];
        o.ctype[row] = CSTR_EQ;
        if (derivs)
          for (int j = 0; j < nv; j++)
            o.Cx(row, j) = Jp[f](2, j);
        row++;
      }
    }
    if (row != o.nc)
      throw std::runtime_error("kino_stage_eval: constraint row count mismatch");
  }

  // ---------------------------------------------------------------------------------------------
  // Full-dynamics stage (reference src/fulldynamics.cpp:78-214), point feet (force_size 3) or quad feet (force_size 6)
  //   u = tau_joint (nv-6), tau = [0; u] (:35-37); a, lam = rigid-contact forward dynamics (orc_full.hpp)
  //   semi-implicit Euler (:140)
  // Costs (:88-137): state, control (target u_ref = 0), centroidal momentum, foot translation per foot,
  //   contact-force lam_f - f_ref per STANCE foot.
  // Constraint rows, in the reference's insertion order (:145-211):
  //   [torque box on u] [joint-limit box on q_joints] then per foot IN CONTACT:
  //   [friction pyramid A lam_f <= 0, 5 rows: aligator MultibodyFrictionConeResidual [UPSTREAM]]
  //   [on a landing edge: LOCAL_WORLD_ALIGNED foot velocity (3 rows) = 0 and z-height = ref].
  // ---------------------------------------------------------------------------------------------
  inline void friction_pyramid(double mu, double * A /*5x3*/)
  {
    const double rows[5][3] = {{0, 0, -1}, {-1, 0, -mu}, {1, 0, -mu}, {0, -1, -mu}, {0, 1, -mu}};
    memcpy(A, rows, sizeof(rows));
  }
  inline void full_stage_eval(const Problem & P, const smpc_slot & s, const double * x, const double * u, bool derivs, StageEval & o)
  {
    const Model & m = P.model;
    const int nv = m.nv, nq = m.nq, ndx = P.ndx, nu = P.nu, nk = P.nk, fs = P.fs;
    const int nj = nv - 6;
    const double dt = P.ocp.timestep;
    const int cd = fs; // rows per contact: 3 (CONTACT_3D, LOCAL) or 6 (CONTACT_6D, LOCAL_WORLD_ALIGNED)
    const double * q = x;
    const double * v = x + nq;
    alloc_eval(P, P.stage_nc(s), derivs, o);
    Vec tau(nv, 0.0);
    for (int j = 0; j < nj; j++)
      tau[6 + j] = u[j];
    FullDyn fd;
    full_dynamics(m, q, v, tau.data(), s.contact_mask, P.ocp.gravity, P.ocp.Kp_correction, P.ocp.Kd_correction, derivs, fd, cd);
    const Kin & k = fd.k;
    const Vec & a = fd.a;
    for (int i = 0; i < nv; i++)
    {
      o.xdot[i] = v[i];
      o.xdot[nv + i] = a[i];
    }
    Vec vn(nv), dq(nv);
    for (int i = 0; i < nv; i++)
    {
      vn[i] = v[i] + dt * a[i];
      dq[i] = dt * vn[i];
    }
    integrate_q(m, q, dq.data(), o.xnext.data());
    for (int i = 0; i < nv; i++)
      o.xnext[nq + i] = vn[i];

    std::vector<Mat> Jp(nk);
    if (derivs)
    {
      for (int f = 0; f < nk; f++)
      {
        Jp[f].resize(3, nv);
        for (int j = 0; j < nv; j++)
        {
          double c3[3];
          point_jac_col(m, k, m.d.foot_joint[f], k.oMf[f].p, j, c3);
          for (int r = 0; r < 3; r++)
            Jp[f](r, j) = c3[r];
        }
      }
      double Jq6[36], Jv6[36];
      {
        SE3 E = exp6(dq.data());
        se3_actinv_matrix(E, Jq6);
        Jexp6(dq.data(), Jv6);
      }
      Mat dvn_dx(nv, ndx), dvn_du(nv, nu);
      for (int r = 0; r < nv; r++)
      {
        for (int j = 0; j < nv; j++)
        {
          dvn_dx(r, j) = dt * fd.da_dq(r, j);
          dvn_dx(r, nv + j) = dt * fd.da_dv(r, j);
        }
        dvn_dx(r, nv + r) += 1.0;
        for (int j = 0; j < nu; j++)
          dvn_du(r, j) = dt * fd.da_dtau(r, 6 + j);
      }
      for (int i = 0; i < nv; i++)
      {
        for (int j = 0; j < ndx; j++)
        {
          o.A(nv + i, j) = dvn_dx(i, j);
          double sacc = 0;
          if (i < 6)
            for (int c = 0; c < 6; c++)
              sacc += Jv6[6 * i + c] * dvn_dx(c, j);
          else
            sacc = dvn_dx(i, j);
          o.A(i, j) = dt * sacc;
        }
        for (int j = 0; j < nu; j++)
        {
          o.B(nv + i, j) = dvn_du(i, j);
          double sacc = 0;
          if (i < 6)
            for (int c = 0; c < 6; c++)
              sacc += Jv6[6 * i + c] * dvn_du(c, j);
          else
            sacc = dvn_du(i, j);
          o.B(i, j) = dt * sacc;
        }
      }
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++)
          o.A(i, j) += Jq6[6 * i + j];
      for (int i = 6; i < nv; i++)
        o.A(i, i) += 1.0;
    }

    // ---------------- costs ----------------
    {
      Vec r;
      Mat J;
      state_error(P, x, s.x_ref, r, derivs ? &J : nullptr);
      add_quad_cost(r, P.w_x, &J, nullptr, derivs, o);
    }
    {
      Vec r(nu);
      for (int i = 0; i < nu; i++)
        r[i] = u[i] - s.u_ref[i];
      Mat Ju;
      if (derivs)
      {
        Ju.resize(nu, nu);
        for (int i = 0; i < nu; i++)
          Ju(i, i) = 1.0;
      }
      add_quad_cost(r, P.w_u, nullptr, &Ju, derivs, o);
    }
    {
      Vec r(k.hg, k.hg + 6);
      Mat Jx;
      if (derivs)
      {
        Jx.resize(6, ndx);
        for (int rr = 0; rr < 6; rr++)
          for (int j = 0; j < nv; j++)
          {
            Jx(rr, j) = k.dh_dq(rr, j);
            Jx(rr, nv + j) = k.Ag(rr, j);
          }
      }
      add_quad_cost(r, P.w_cent, &Jx, nullptr, derivs, o);
    }
    for (int f = 0; f < nk; f++)
    {
      if (fs == 6)
      {
        // FramePlacementResidual (reference src/fulldynamics.cpp:99-104): log6(ref^-1 oMf), J = Jlog6 * (LOCAL frame Jacobian)
        frame_placement_cost(P, s, k, f, derivs, o);
        continue;
      }
      Vec r(3);
      for (int rr = 0; rr < 3; rr++)
        r[rr] = k.oMf[f].p[rr] - s.foot_ref[f][9 + rr];
      Mat Jx;
      if (derivs)
      {
        Jx.resize(3, ndx);
        for (int rr = 0; rr < 3; rr++)
          for (int j = 0; j < nv; j++)
            Jx(rr, j) = Jp[f](rr, j);
      }
      add_quad_cost(r, P.w_frame, &Jx, nullptr, derivs, o);
    }
    o.fc.assign((size_t)P.nk * 3, 0.0);
    // MPC::getContactForces keeps the linear part of the contact force (reference src/mpc.cpp:369)
    for (size_t cf = 0; cf < fd.feet.size(); cf++)
      for (int rr = 0; rr < 3; rr++)
        o.fc[(size_t)fd.feet[cf] * 3 + rr] = fd.lam[cd * cf + rr];
    for (size_t cf = 0; cf < fd.feet.size(); cf++)
    {
      const int f = fd.feet[cf];
      Vec r(cd);
      for (int rr = 0; rr < cd; rr++)
        r[rr] = fd.lam[cd * cf + rr] - s.f_ref[f][rr];
      Mat Jx, Ju;
      if (derivs)
      {
        Jx.resize(cd, ndx);
        Ju.resize(cd, nu);
        for (int rr = 0; rr < cd; rr++)
        {
          for (int j = 0; j < nv; j++)
          {
            Jx(rr, j) = fd.dl_dq(cd * (int)cf + rr, j);
            Jx(rr, nv + j) = fd.dl_dv(cd * (int)cf + rr, j);
          }
          for (int j = 0; j < nu; j++)
            Ju(rr, j) = fd.dl_dtau(cd * (int)cf + rr, 6 + j);
        }
      }
      add_quad_cost(r, P.w_forces, &Jx, &Ju, derivs, o);
    }

    // ---------------- constraints ----------------
    int row = 0;
    if (P.ocp.torque_limits)
      for (int j = 0; j < nj; j++, row++)
      {
        o.c[row] = u[j];
        o.ctype[row] = CSTR_BOX;
        o.lo[row] = P.umin[j];
        o.hi[row] = P.umax[j];
        if (derivs)
          o.Cu(row, j) = 1.0;
      }
    if (P.ocp.kinematics_limits)
      for (int j = 0; j < nj; j++, row++)
      {
        o.c[row] = q[7 + j];
        o.ctype[row] = CSTR_BOX;
        o.lo[row] = P.qmin[j];
        o.hi[row] = P.qmax[j];
        if (derivs)
          o.Cx(row, 6 + j) = 1.0;
      }
    for (size_t cf = 0; cf < fd.feet.size(); cf++)
    {
      const int f = fd.feet[cf], jf = m.d.foot_joint[f];
      if (P.ocp.force_cone)
      {
        // point feet: 5-row friction pyramid on lam_f (MultibodyFrictionConeResidual); quad feet: 17-row wrench cone on the
        // 6-D contact force, which constraintDynamics returns in the LOCAL_WORLD_ALIGNED frame (MultibodyWrenchConeResidual)
        double Ac[17 * 6];
        const int nrows = (fs == 6) ? 17 : 5;
        if (fs == 6)
          wrench_cone_matrix(P.ocp.mu, P.ocp.Lfoot, P.ocp.Wfoot, Ac);
        else
          friction_pyramid(P.ocp.mu, Ac);
        for (int rr = 0; rr < nrows; rr++, row++)
        {
          double sacc = 0;
          for (int c = 0; c < cd; c++)
            sacc += Ac[cd * rr + c] * fd.lam[cd * cf + c];
          o.c[row] = sacc;
          o.ctype[row] = CSTR_NEG;
          if (derivs)
            for (int c = 0; c < cd; c++)
            {
              for (int j = 0; j < nv; j++)
              {
                o.Cx(row, j) += Ac[cd * rr + c] * fd.dl_dq(cd * (int)cf + c, j);
                o.Cx(row, nv + j) += Ac[cd * rr + c] * fd.dl_dv(cd * (int)cf + c, j);
              }
              for (int j = 0; j < nu; j++)
                o.Cu(row, j) += Ac[cd * rr + c] * fd.dl_dtau(cd * (int)cf + c, 6 + j);
            }
        }
      }
      if (fs == 6 && P.ocp.land_cstr && (s.land_mask & (1u << f)))
      {
        // 6-D LOCAL_WORLD_ALIGNED frame velocity = 0 on a landing edge (reference src/fulldynamics.cpp:175-181)
        const double * vi = &k.ov[6 * jf];
        const double * p = k.oMf[f].p;
        double V[6];
        frame_motion_lwa(vi, p, V);
        for (int rr = 0; rr < 6; rr++)
        {
          o.c[row + rr] = V[rr];
          o.ctype[row + rr] = CSTR_EQ;
        }
        if (derivs)
          for (int j = 0; j < nv; j++)
          {
            const int ij = m.joint_of_dof[j];
            if (!m.anc[jf][ij])
              continue;
            const int pj = m.d.parent[ij];
            double Jm[6], dvp[6], dvq[6], dV[6], Jpm[6], t2[3];
            for (int rr = 0; rr < 6; rr++)
            {
              Jm[rr] = k.J(rr, j);
              dvp[rr] = ((pj >= 0) ? k.ov[6 * pj + rr] : 0.0) - vi[rr];
            }
            mxm(dvp, Jm, dvq);
            frame_motion_lwa(dvq, p, dV);
            frame_motion_lwa(Jm, p, Jpm);
            cross3(vi + 3, Jpm, t2);
            for (int rr = 0; rr < 3; rr++)
              dV[rr] += t2[rr];
            for (int rr = 0; rr < 6; rr++)
            {
              o.Cx(row + rr, j) = dV[rr];
              o.Cx(row + rr, nv + j) = Jpm[rr];
            }
          }
        row += 6;
      }
      if (fs == 3 && P.ocp.land_cstr && (s.land_mask & (1u << f)))
      {
        const double * vi = &k.ov[6 * jf];
        const double * p = k.oMf[f].p;
        double wxp[3], pd[3];
        cross3(vi + 3, p, wxp);
        for (int rr = 0; rr < 3; rr++)
          pd[rr] = vi[rr] + wxp[rr];
        for (int rr = 0; rr < 3; rr++)
        {
          o.c[row + rr] = pd[rr];
          o.ctype[row + rr] = CSTR_EQ;
        }
        if (derivs)
          for (int j = 0; j < nv; j++)
          {
            const int ij = m.joint_of_dof[j];
            if (!m.anc[jf][ij])
              continue;
            const int pj = m.d.parent[ij];
            double Jm[6], dvp[6], dvq[6], t1[3], t2[3], Jpj[3] = {Jp[f](0, j), Jp[f](1, j), Jp[f](2, j)};
            for (int rr = 0; rr < 6; rr++)
            {
              Jm[rr] = k.J(rr, j);
              dvp[rr] = ((pj >= 0) ? k.ov[6 * pj + rr] : 0.0) - vi[rr];
            }
            mxm(dvp, Jm, dvq);
            cross3(dvq + 3, p, t1);
            cross3(vi + 3, Jpj, t2);
            for (int rr = 0; rr < 3; rr++)
            {
              o.Cx(row + rr, j) = dvq[rr] + t1[rr] + t2[rr];
              o.Cx(row + rr, nv + j) = Jpj[rr];
            }
          }
        row += 3;
        o.c[row] = p[2] - s.land_z[f];
        o.ctype[row] = CSTR_EQ;
        if (derivs)
          for (int j = 0; j < nv; j++)
            o.Cx(row, j) = Jp[f](2, j);
        row++;
      }
    }
    if (row != o.nc)
      throw std::runtime_error("full_stage_eval: constraint row count mismatch");
  }

  // Terminal node of kinodynamics / full dynamics (reference src/kinodynamics.cpp:352-377):
  // state_cost + centroidal_cost with 10 w_cent; optional DCM equality  c + tau cdot - ref = 0.
  inline void multibody_terminal_eval(const Problem & P, const smpc_terminal & tm, const double * x, bool derivs, StageEval & o)
  {
    const Model & m = P.model;
    const int nv = m.nv, nq = m.nq, ndx = P.ndx;
    alloc_eval(P, P.term_nc(), derivs, o);
    Kin k;
    centroidal_pass(m, x, x + nq, nullptr, derivs, k);
    {
      Vec r;
      Mat J;
      state_error(P, x, tm.x_ref, r, derivs ? &J : nullptr);
      add_quad_cost(r, P.w_x, &J, nullptr, derivs, o);
    }
    {
      Vec r(k.hg, k.hg + 6);
      Mat Jx;
      if (derivs)
      {
        Jx.resize(6, ndx);
        for (int rr = 0; rr < 6; rr++)
          for (int j = 0; j < nv; j++)
          {
            Jx(rr, j) = k.dh_dq(rr, j);
            Jx(rr, nv + j) = k.Ag(rr, j);
          }
      }
      Mat W10 = P.w_cent;
      for (auto & e : W10.d)
        e *= 10.0; // reference src/kinodynamics.cpp:361
      add_quad_cost(r, W10, &Jx, nullptr, derivs, o);
    }
    if (P.term_nc() == 3)
    {
      const double tau = tm.tau; // fixed at createTerminalConstraint (reference src/kinodynamics.cpp:372); setReference only moves com_ref
      for (int rr = 0; rr < 3; rr++)
      {
        o.c[rr] = k.com[rr] + tau * k.vcom[rr] - tm.com_ref[rr];
        o.ctype[rr] = CSTR_EQ;
        if (derivs)
          for (int j = 0; j < nv; j++)
          {
            o.Cx(rr, j) = k.Jcom(rr, j) + tau * k.dh_dq(rr, j) / P.mass;
            o.Cx(rr, nv + j) = tau * k.Jcom(rr, j);
          }
      }
    }
  }

  // ---------------------------------------------------------------------------------------------
  // Centroidal stage (reference src/centroidal-dynamics.cpp:39-106)
  //   x = [c; h_lin; h_ang], u = [f_0 .. f_{nk-1}] ; contact positions p_i come from the slot's foot_ref
  //   xdot = [h_lin/m ; sum f + m g ; sum (p_i - c) x f_i (+ tau_i)]   explicit Euler
  // Costs: com, control, linear/angular momentum, linear/angular acceleration (:72-77).
  // Constraints: wrench (17 rows) or friction cone (2 rows, eps = 1e-4) per foot in contact (:85-104).
  // ---------------------------------------------------------------------------------------------
  inline void centroidal_stage_eval(const Problem & P, const smpc_slot & s, const double * x, const double * u, bool derivs, StageEval & o)
  {
    const int nu = P.nu, nk = P.nk, fs = P.fs;
    const double dt = P.ocp.timestep, mass = P.mass;
    alloc_eval(P, P.stage_nc(s), derivs, o);
    double lin[3] = {mass * P.ocp.gravity[0], mass * P.ocp.gravity[1], mass * P.ocp.gravity[2]};
    double ang[3] = {0, 0, 0};
    Mat dlin_du(3, nu), dang_du(3, nu), dang_dc(3, 3);
    for (int f = 0; f < nk; f++)
      if (s.contact_mask & (1u << f))
      {
        const double * ff = u + fs * f;
        double lever[3], t[3];
        for (int r = 0; r < 3; r++)
          lever[r] = s.foot_ref[f][9 + r] - x[r];
        cross3(lever, ff, t);
        double Lx[9], Fx[9];
        skew3(lever, Lx);
        skew3(ff, Fx);
        for (int r = 0; r < 3; r++)
        {
          lin[r] += ff[r];
          ang[r] += t[r] + (fs == 6 ? ff[3 + r] : 0.0);
          dlin_du(r, fs * f + r) = 1.0;
          for (int c = 0; c < 3; c++)
          {
            dang_du(r, fs * f + c) = Lx[3 * r + c];
            dang_dc(r, c) += Fx[3 * r + c]; // d/dc (p - c) x f = [f]x
          }
          if (fs == 6)
            dang_du(r, fs * f + 3 + r) = 1.0;
        }
      }
    for (int r = 0; r < 3; r++)
    {
      o.xdot[r] = x[3 + r] / mass;
      o.xdot[3 + r] = lin[r];
      o.xdot[6 + r] = ang[r];
    }
    for (int i = 0; i < 9; i++)
      o.xnext[i] = x[i] + dt * o.xdot[i];
    if (derivs)
    {
      for (int i = 0; i < 9; i++)
        o.A(i, i) = 1.0;
      for (int r = 0; r < 3; r++)
      {
        o.A(r, 3 + r) += dt / mass;
        for (int c = 0; c < 3; c++)
          o.A(6 + r, c) += dt * dang_dc(r, c);
        for (int j = 0; j < nu; j++)
        {
          o.B(3 + r, j) = dt * dlin_du(r, j);
          o.B(6 + r, j) = dt * dang_du(r, j);
        }
      }
    }
    // costs
    auto sel = [&](int off) {
      Mat J(3, 9);
      for (int r = 0; r < 3; r++)
        J(r, off + r) = 1.0;
      return J;
    };
    {
      Vec r(3);
      for (int i = 0; i < 3; i++)
        r[i] = x[i] - s.x_ref[i]; // com reference lives in x_ref[0:3]
      Mat J = sel(0);
      add_quad_cost(r, P.w_com, &J, nullptr, derivs, o);
    }
    {
      Vec r(nu);
      for (int i = 0; i < nu; i++)
        r[i] = u[i] - s.u_ref[i];
      Mat Ju(nu, nu);
      for (int i = 0; i < nu; i++)
        Ju(i, i) = 1.0;
      add_quad_cost(r, P.w_u, nullptr, &Ju, derivs, o);
    }
    {
      Vec r(3);
      for (int i = 0; i < 3; i++)
        r[i] = x[3 + i] - s.x_ref[3 + i];
      Mat J = sel(3);
      add_quad_cost(r, P.w_lmom, &J, nullptr, derivs, o);
    }
    {
      Vec r(3);
      for (int i = 0; i < 3; i++)
        r[i] = x[6 + i] - s.x_ref[6 + i];
      Mat J = sel(6);
      add_quad_cost(r, P.w_amom, &J, nullptr, derivs, o);
    }
    {
      // CentroidalAccelerationResidual: (sum f)/m + g  [UPSTREAM]
      Vec r(3);
      for (int i = 0; i < 3; i++)
        r[i] = lin[i] / mass;
      Mat Jx(3, 9), Ju(3, nu);
      for (int rr = 0; rr < 3; rr++)
        for (int j = 0; j < nu; j++)
          Ju(rr, j) = dlin_du(rr, j) / mass;
      add_quad_cost(r, P.w_lacc, &Jx, &Ju, derivs, o);
    }
    {
      // AngularAccelerationResidual: sum (p - c) x f (+ tau)  [UPSTREAM]
      Vec r(ang, ang + 3);
      Mat Jx(3, 9);
      for (int rr = 0; rr < 3; rr++)
        for (int c = 0; c < 3; c++)
          Jx(rr, c) = dang_dc(rr, c);
      add_quad_cost(r, P.w_aacc, &Jx, &dang_du, derivs, o);
    }
    int row = 0;
    for (int f = 0; f < nk; f++)
    {
      if (!(s.contact_mask & (1u << f)))
        continue;
      const double * ff = u + fs * f;
      if (fs == 6)
      {
        double Ac[17 * 6];
        wrench_cone_matrix(P.ocp.mu, P.ocp.Lfoot, P.ocp.Wfoot, Ac);
        for (int rr = 0; rr < 17; rr++, row++)
        {
          double sacc = 0;
          for (int c = 0; c < 6; c++)
            sacc += Ac[6 * rr + c] * ff[c];
          o.c[row] = sacc;
          o.ctype[row] = CSTR_NEG;
          if (derivs)
            for (int c = 0; c < 6; c++)
              o.Cu(row, fs * f + c) = Ac[6 * rr + c];
        }
      }
      else
      {
        const double mu = P.ocp.mu, eps = 1e-4; // reference src/centroidal-dynamics.cpp:99
        o.c[row] = -ff[2] + eps;
        o.ctype[row] = CSTR_NEG;
        o.c[row + 1] = -mu * mu * ff[2] * ff[2] + ff[0] * ff[0] + ff[1] * ff[1];
        o.ctype[row + 1] = CSTR_NEG;
        if (derivs)
        {
          o.Cu(row, fs * f + 2) = -1.0;
          o.Cu(row + 1, fs * f + 0) = 2 * ff[0];
          o.Cu(row + 1, fs * f + 1) = 2 * ff[1];
          o.Cu(row + 1, fs * f + 2) = -2 * mu * mu * ff[2];
        }
        row += 2;
      }
    }
  }
  // Centroidal terminal cost: the two momentum costs (reference src/centroidal-dynamics.cpp:306-316)
  inline void centroidal_terminal_eval(const Problem & P, const smpc_terminal & tm, const double * x, bool derivs, StageEval & o)
  {
    alloc_eval(P, 0, derivs, o);
    for (int part = 0; part < 2; part++)
    {
      int off = 3 + 3 * part;
      Vec r(3);
      for (int i = 0; i < 3; i++)
        r[i] = x[off + i] - tm.x_ref[off + i];
      Mat J(3, 9);
      for (int rr = 0; rr < 3; rr++)
        J(rr, off + rr) = 1.0;
      add_quad_cost(r, part == 0 ? P.w_lmom : P.w_amom, &J, nullptr, derivs, o);
    }
  }

  inline void stage_eval(const Problem & P, const smpc_slot & s, const double * x, const double * u, bool derivs, StageEval & o)
  {
    if (P.ocp.kind == SMPC_OCP_KINODYNAMICS)
      kino_stage_eval(P, s, x, u, derivs, o);
    else if (P.ocp.kind == SMPC_OCP_CENTROIDAL)
      centroidal_stage_eval(P, s, x, u, derivs, o);
    else
      full_stage_eval(P, s, x, u, derivs, o);
  }
  inline void terminal_eval(const Problem & P, const smpc_terminal & tm, const double * x, bool derivs, StageEval & o)
  {
    if (P.ocp.kind == SMPC_OCP_CENTROIDAL)
      centroidal_terminal_eval(P, tm, x, derivs, o);
    else
      multibody_terminal_eval(P, tm, x, derivs, o);
  }

  // state-space operations of the OCP (MultibodyPhaseSpace or R^9)
  inline void state_integrate(const Problem & P, const double * x, const double * dx, double * xo)
  {
    if (P.ocp.kind == SMPC_OCP_CENTROIDAL)
    {
      for (int i = 0; i < 9; i++)
        xo[i] = x[i] + dx[i];
      return;
    }
    const Model & m = P.model;
    integrate_q(m, x, dx, xo);
    for (int i = 0; i < m.nv; i++)
      xo[m.nq + i] = x[m.nq + i] + dx[m.nv + i];
  }
  // d = x1 (-) x0
  inline void state_difference(const Problem & P, const double * x0, const double * x1, double * dd)
  {
    if (P.ocp.kind == SMPC_OCP_CENTROIDAL)
    {
      for (int i = 0; i < 9; i++)
        dd[i] = x1[i] - x0[i];
      return;
    }
    const Model & m = P.model;
    difference_q(m, x0, x1, dd);
    for (int i = 0; i < m.nv; i++)
      dd[m.nv + i] = x1[m.nq + i] - x0[m.nq + i];
  }

} // namespace orc
