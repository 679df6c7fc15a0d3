// full_knot.cuh — one horizon knot of the FULL-DYNAMICS problem, computed by one CTA.
//
// What it replaces (all [UPSTREAM] aligator/pinocchio code reached from the stage that
// FullDynamicsOCP::createStage builds, reference src/fulldynamics.cpp:78-214):
//   MultibodyConstraintFwdDynamics::forward/dForward  (pinocchio::constraintDynamics +
//   computeConstraintDynamicsDerivatives with CONTACT_3D / LOCAL rigid constraints on point feet or CONTACT_6D /
//   LOCAL_WORLD_ALIGNED ones on quad feet, Baumgarte corrector, proximal settings 1e-9 / 1e-10 / 10, reference
//   :39,50-75), IntegratorSemiImplEuler, QuadraticStateCost, QuadraticControlCost, CentroidalMomentumResidual,
//   FrameTranslationResidual / FramePlacementResidual, ContactForceResidual, ControlErrorResidual box (torque limits),
//   StateErrorResidual slice box (joint limits), MultibodyFrictionConeResidual (5-row pyramid) /
//   MultibodyWrenchConeResidual (17 rows), FrameVelocityResidual(LOCAL_WORLD_ALIGNED) and FrameTranslationResidual
//   slice on landing edges, and the per-stage ProxDDP work (dense_knot.cuh).
//   SURVEY.md §8a rows a4, a6, a8, a9, a10. The contact dynamics is computed ONCE per knot and shared by every
//   residual that needs lam or its derivatives (the reference re-runs it inside each residual).
//
//   M a - Jc^T lam = tau - b ,  Jc a + gamma = 0          tau = [0; u]
//   solved through L L^T = M, G = Jc M^-1 Jc^T + mu I (two proximal iterations), then implicit differentiation
//   with the same factors:  X1 = M^-1 [-dtau/dq | -dtau/dv | I],  dlam = G^-1 (r2 - Jc X1),  da = X1 + M^-1 Jc^T dlam.
//   dtau/dq, dtau/dv: world-frame RNEA derivatives (Carpentier & Mansard 2018) with the contact forces as external
//   forces held fixed in the foot frames; d(Jc a + gamma)/dq, /dv in closed form. Same formulas as oracle/orc_full.hpp.
//   Quad feet (FS = 6): 6 rows per contact, c = [R 0; 0 R](a_loc + Kd o v_loc + Kp o log6(oMc)), lam = [f; tau] in
//   world-aligned axes at the contact-frame origin; such a force does not turn with the foot, which adds
//   J_k . [w_c x f ; p x (w_c x f) + w_c x tau] to d tau_k / d q_c (see oracle/orc_full.hpp, `lwa`).
//
// Work distribution: tree levels / dof columns / flat (row, col) index loops as in kino_knot.cuh; the two Cholesky
// factorisations are left-looking with one thread per row (two barriers per column, same dot-product order as the
// oracle); triangular solves run one thread per right-hand side. Constraint Jacobian rows are written straight into
// the knot's C / D blocks (FullScratch::Cx/Cu point there) to keep the shared-memory footprint at ~50 KB.
#pragma once
#include "dense_knot.cuh"

namespace smpc
{

  template <int NJ_, int NV_, int NK_, int FS_ = 3>
  struct FullCfg
  {
    static constexpr int NJ = NJ_, NV = NV_, NK = NK_, FS = FS_;
    static_assert(FS == 3 || FS == 6, "point feet (CONTACT_3D) or quad feet (CONTACT_6D)");
    static constexpr int NQ = NV + 1, NX = NQ + NV, NDX = 2 * NV, NJNT = NV - 6, NU = NJNT;
    static constexpr int CD = FS;        // rows per contact
    static constexpr int NCON = CD * NK; // contact rows when every foot is in contact
    static constexpr int NCOL = 3 * NV;  // columns of the sensitivity systems: q | v | tau
    static constexpr int NZ = NDX + NU;  // columns of a stage Jacobian: x | u
    static constexpr int NCONE = (FS == 6) ? 17 : 5; // wrench cone / friction pyramid rows per stance foot
    static constexpr int NLAND = (FS == 6) ? 6 : 4;  // landing-edge rows: 6-D velocity / 3 velocity rows + height
    static constexpr int NCMAX = 2 * NJNT + NK * (NCONE + NLAND);
  };

  enum FullRowKind
  {
    FROW_TBOX = 0,
    FROW_QBOX = 1,
    FROW_CONE = 2,
    FROW_LVEL = 3,
    FROW_LZ = 4
  };

  template <class Cfg>
  struct FullScratch
  {
    static constexpr int NJ = Cfg::NJ, NV = Cfg::NV, NK = Cfg::NK, NX = Cfg::NX, NU = Cfg::NU, NDX = Cfg::NDX;
    static constexpr int NCON = Cfg::NCON, NCOL = Cfg::NCOL, NZ = Cfg::NZ, NCMAX = Cfg::NCMAX, FS = Cfg::FS;
    static constexpr int Q6 = (FS == 6) ? 6 : 1, Q36 = (FS == 6) ? 36 : 1; // sizes of the quad-foot-only arrays
    double x[NX], u[NU], xn_ws[NX], xnext[NX];
    // kinematics (member names shared with kino_kinematics / kino_accel_pass / kino_centroid)
    double R[NJ][9], p[NJ][3], v[NJ][6], a[NJ][6], bias[NJ][6];
    double Yb[NJ][10], hb[NJ][6];
    double Y[NJ][10], hs[NJ][6], dIo[NJ][6], fs[NJ][6];
    double J[6][NV];
    double fR[NK][9], fp[NK][3], lever[NK][3], vf[NK][6];
    double com[3], vcom[3], h0[6], hd0[6], hg[6], hdg[6];
    double Ag[6][NV], dh_dq[6][NV];
    // foot Jacobians in world-aligned axes at the foot-frame origin: rows 0-2 linear; quad feet also rows 3-5 angular
    double Jp[NK][FS][NV];
    // quad feet: log6 of the foot pose and its Jacobian (Baumgarte position term), pose-cost residual log6(ref^-1 oMf)
    // and T = Jlog6(residual) blkdiag(R^T, R^T), which maps the world-aligned Jacobian to the residual's
    double flog[NK][Q6], fJl[NK][Q36], rpl[NK][Q6], Tpl[NK][Q36];
    // contact dynamics
    int ncon, cidx[NK]; // cidx[f]: index of foot f among the feet in contact, -1 if swinging
    double Mm[NV][NV], dgM[NV];
    double Jc[NCON][NV], MiJt[NV][NCON], G[NCON][NCON], dgG[NCON];
    double tau[NV], nle[NV], af[NV], acc[NV], gamma[NCON], rhs0[NCON], lam[NCON];
    double Fs[NJ][6];
    double X1[NV][NCOL], DL[NCON][NCOL];
    // integrator, state error
    double Jq6[36], Jv6[36], Jl6[36];
    double rs[NDX];
    // costs
    double WJc[6][NDX], Wch[6], WJp[NK][FS][NV], Wfr[NK][FS], WJl[NCON][NZ], Wlr[NCON];
    double costs[3 + 2 * NK];
    // constraint rows
    int rkind[NCMAX], rarg[NCMAX], rsub[NCMAX];
    // dense tail (dense_knot.cuh); Cx / Cu live in the knot's C / D blocks
    double e[NDX], lplus[NDX], xdot[NDX], lx[NDX], lu[NU];
    double (*Cx)[NDX];
    double (*Cu)[NU];
    double cval[NCMAX], lo[NCMAX], hi[NCMAX];
    int ctype[NCMAX];
    double dproj[NCMAX], vplus[NCMAX], rpinf[NCMAX], rcrit[NCMAX];
    int ractive[NCMAX];
    int nc;
    double cost;
    double Lx[NDX], Lu[NU], gx[NDX], gu[NU];
  };

  // Left-looking Cholesky of the leading n x n block of a row-major matrix with leading dimension ld (lower
  // triangle), one thread per row, diagonal staged in dg. flag != NULL receives 1 on a non-positive pivot.
  template <int LD>
  SMPC_DEV void coop_cholesky(double (*A)[LD], double * dg, int n, int * flag)
  {
    for (int j = 0; j < n; j++)
    {
      SMPC_PAR(tid)
      for (int i = j + tid; i < n; i += SMPC_NT)
      {
        double sacc = A[i][j];
        for (int k = 0; k < j; k++)
          sacc -= A[i][k] * A[j][k];
        if (i == j)
        {
          if (!(sacc > 0) && flag)
            *flag = 1;
          dg[j] = sqrt(sacc);
        }
        else
          A[i][j] = sacc;
      }
      SMPC_END
      SMPC_PAR(tid)
      for (int i = j + tid; i < n; i += SMPC_NT)
      {
        if (i == j)
          A[j][j] = dg[j];
        else
          A[i][j] = A[i][j] / dg[j];
      }
      SMPC_END
    }
  }
  // x <- (L L^T)^-1 x for one strided vector (one thread)
  template <int LD>
  SMPC_DEV void chol_solve_vec(const double (*Lm)[LD], int n, double * x, int stride)
  {
    for (int i = 0; i < n; i++)
    {
      double sacc = x[i * stride];
      for (int k = 0; k < i; k++)
        sacc -= Lm[i][k] * x[k * stride];
      x[i * stride] = sacc / Lm[i][i];
    }
    for (int i = n - 1; i >= 0; i--)
    {
      double sacc = x[i * stride];
      for (int k = i + 1; k < n; k++)
        sacc -= Lm[k][i] * x[k * stride];
      x[i * stride] = sacc / Lm[i][i];
    }
  }

  // velocity and classical acceleration of the material point p of a body with world-frame spatial (v, a)
  SMPC_DEV void point_va(const double * v, const double * a, const double * p, double * pd, double * pdd)
  {
    double t[3], t1[3], t2[3];
    v3cross(v + 3, p, t);
    for (int r = 0; r < 3; r++)
      pd[r] = v[r] + t[r];
    v3cross(a + 3, p, t1);
    v3cross(v + 3, pd, t2);
    for (int r = 0; r < 3; r++)
      pdd[r] = a[r] + t1[r] + t2[r];
  }

  // world-axes motion of a frame origin p carried by a body with world-frame spatial motion mo: [lin + ang x p ; ang]
  SMPC_DEV void frame_motion_lwa(const double * mo, const double * p, double * out)
  {
    double t[3];
    v3cross(mo + 3, p, t);
    for (int r = 0; r < 3; r++)
    {
      out[r] = mo[r] + t[r];
      out[3 + r] = mo[3 + r];
    }
  }
  // y = [R 0; 0 R] diag(K) [R^T 0; 0 R^T] x
  SMPC_DEV void rkrt6(const double * R, const double * K, const double * x, double * y)
  {
    double l[6];
    m3tv(R, x, l);
    m3tv(R, x + 3, l + 3);
    for (int r = 0; r < 6; r++)
      l[r] *= K[r];
    m3v(R, l, y);
    m3v(R, l + 3, y + 3);
  }
  SMPC_DEV void mat6v(const double * A, const double * v, double * o)
  {
    for (int r = 0; r < 6; r++)
    {
      double sacc = 0;
      for (int k = 0; k < 6; k++)
        sacc += A[6 * r + k] * v[k];
      o[r] = sacc;
    }
  }

  // ---- quad feet (CONTACT_6D, LOCAL_WORLD_ALIGNED): the pieces of full_stage_knot that differ from point feet ----
  // contact-row value gamma (6) of foot f at zero joint acceleration; keeps log6(oMf) and its Jacobian
  template <class Cfg, bool DERIVS>
  SMPC_DEV void full6_gamma(const DevOcp & O, FullScratch<Cfg> & s, int f, int jf, double * gamma6)
  {
    double V[6], Aw[6], t1[6], ke[6], t2[6], e6[6];
    frame_motion_lwa(s.v[jf], s.fp[f], V);
    frame_motion_lwa(s.a[jf], s.fp[f], Aw);
    log6(s.fR[f], s.fp[f], e6);
    for (int r = 0; r < 6; r++)
      s.flog[f][r % FullScratch<Cfg>::Q6] = e6[r];
    if (DERIVS)
    {
      double Jl[36];
      Jlog6(e6, Jl);
      for (int r = 0; r < 36; r++)
        s.fJl[f][r % FullScratch<Cfg>::Q36] = Jl[r];
    }
    rkrt6(s.fR[f], O.Kd, V, t1);
    for (int r = 0; r < 6; r++)
      ke[r] = O.Kp[r] * e6[r];
    m3v(s.fR[f], ke, t2);
    m3v(s.fR[f], ke + 3, t2 + 3);
    for (int r = 0; r < 6; r++)
      gamma6[r] = Aw[r] + t1[r] + t2[r];
  }
  // partial derivatives of the 6-D contact row of foot f with respect to q_c and v_c (dof c on the support chain)
  template <class Cfg>
  SMPC_DEV void full6_contact_derivs(const DevModel & M, const DevOcp & O, const FullScratch<Cfg> & s, int f, int jf, int c, double * dq6, double * dv6)
  {
    const smpc_model_desc & d = M.d;
    const int ic = M.joint_of_dof[c];
    const int par = d.parent[ic];
    const double z6[6] = {0, 0, 0, 0, 0, 0};
    const double * vp = (par >= 0) ? s.v[par] : z6;
    const double * ap = (par >= 0) ? s.a[par] : z6;
    const double * vi = s.v[jf];
    const double * ai = s.a[jf];
    const double * vim = s.v[ic];
    const double * p = s.fp[f];
    const double * Rf = s.fR[f];
    double Jm[6], dvp[6], dap[6], dvi[6], vs[6], dv_q[6], w[6], da1[6], da2[6], da_q[6], da_v[6];
    for (int r = 0; r < 6; r++)
    {
      Jm[r] = s.J[r][c];
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
    double Jpm[6], V[6], dV_q[6], dA_q[6], dA_v[6], t[3];
    frame_motion_lwa(Jm, p, Jpm);
    const double * wm = Jm + 3;
    frame_motion_lwa(vi, p, V);
    frame_motion_lwa(dv_q, p, dV_q);
    v3cross(vi + 3, Jpm, t);
    for (int r = 0; r < 3; r++)
      dV_q[r] += t[r];
    frame_motion_lwa(da_q, p, dA_q);
    v3cross(ai + 3, Jpm, t);
    for (int r = 0; r < 3; r++)
      dA_q[r] += t[r];
    frame_motion_lwa(da_v, p, dA_v);
    // d/dq [R Kd R^T V] = w_m x (R Kd R^T V) + R Kd R^T (dV - w_m x V)
    double kv[6], wxV[6], dVr[6], t2[6], c1[6];
    rkrt6(Rf, O.Kd, V, kv);
    v3cross(wm, V, wxV);
    v3cross(wm, V + 3, wxV + 3);
    for (int r = 0; r < 6; r++)
      dVr[r] = dV_q[r] - wxV[r];
    rkrt6(Rf, O.Kd, dVr, t2);
    v3cross(wm, kv, c1);
    v3cross(wm, kv + 3, c1 + 3);
    // d/dq [R (Kp o e)] = w_m x (R Kp e) + R Kp Jlog6(e) (X_F^-1 J_m),  X_F^-1 J_m = blkdiag(R^T) [J_lin + J_ang x p ; J_ang]
    double ke[6], rke[6], c2[6], Jloc[6], de[6], rde[6], Jl[36];
    for (int r = 0; r < 6; r++)
      ke[r] = O.Kp[r] * s.flog[f][r % FullScratch<Cfg>::Q6];
    for (int r = 0; r < 36; r++)
      Jl[r] = s.fJl[f][r % FullScratch<Cfg>::Q36];
    m3v(Rf, ke, rke);
    m3v(Rf, ke + 3, rke + 3);
    v3cross(wm, rke, c2);
    v3cross(wm, rke + 3, c2 + 3);
    m3tv(Rf, Jpm, Jloc);
    m3tv(Rf, Jpm + 3, Jloc + 3);
    mat6v(Jl, Jloc, de);
    for (int r = 0; r < 6; r++)
      de[r] *= O.Kp[r];
    m3v(Rf, de, rde);
    m3v(Rf, de + 3, rde + 3);
    double t3[6];
    rkrt6(Rf, O.Kd, Jpm, t3);
    for (int r = 0; r < 6; r++)
    {
      dq6[r] = dA_q[r] + c1[r] + t2[r] + c2[r] + rde[r];
      dv6[r] = dA_v[r] + t3[r];
    }
  }
  // <foot>_pose_cost of a quad foot: quad_pose_cost of kino_knot.cuh (shared with the kinodynamics knot)
  template <class Cfg, bool DERIVS>
  SMPC_DEV double full6_pose_cost(const DevOcp & O, const smpc_slot & slot, FullScratch<Cfg> & s, int f)
  {
    return quad_pose_cost<Cfg, DERIVS>(O, slot, s, f);
  }

  template <class Cfg, bool DERIVS>
  SMPC_DEV void full_stage_knot(const DevModel & M, const DevOcp & O, const KnotLayout & L, const KnotIO & io, FullScratch<Cfg> & s)
  {
    constexpr int NV = Cfg::NV, NQ = Cfg::NQ, NX = Cfg::NX, NDX = Cfg::NDX, NU = Cfg::NU, NK = Cfg::NK, NJ = Cfg::NJ;
    constexpr int NJNT = Cfg::NJNT, NCOL = Cfg::NCOL, NZ = Cfg::NZ, FS = Cfg::FS, CD = Cfg::CD;
    const smpc_model_desc & d = M.d;
    const smpc_slot & slot = *io.slot;
    const unsigned cmask = slot.contact_mask;
    const double dt = O.dt;
    const double g6[6] = {O.gravity[0], O.gravity[1], O.gravity[2], 0, 0, 0};

    // ---- load (and, for a trial, step) the iterate; contact list; constraint row table ----
    SMPC_PAR(tid)
    {
      if (io.dx == nullptr)
      {
        for (int k = tid; k < NX; k += SMPC_NT)
        {
          s.x[k] = io.x[k];
          s.xn_ws[k] = io.xn[k];
        }
        for (int k = tid; k < NU; k += SMPC_NT)
          s.u[k] = io.u[k];
      }
      else
      {
        if (tid == 0)
        {
          double sd[NDX];
          for (int k = 0; k < NDX; k++)
            sd[k] = io.alpha * io.dx[k];
          integrate_q(NV, io.x, sd, s.x);
          for (int k = 0; k < NV; k++)
            s.x[NQ + k] = io.x[NQ + k] + sd[NV + k];
          for (int k = 0; k < NX; k++)
            io.tx[k] = s.x[k];
        }
        else if (tid == 1)
        {
          double sd[NDX];
          for (int k = 0; k < NDX; k++)
            sd[k] = io.alpha * io.dxn[k];
          integrate_q(NV, io.xn, sd, s.xn_ws);
          for (int k = 0; k < NV; k++)
            s.xn_ws[NQ + k] = io.xn[NQ + k] + sd[NV + k];
        }
        for (int k = tid; k < NU; k += SMPC_NT)
        {
          s.u[k] = io.u[k] + io.alpha * io.du[k];
          io.tu[k] = s.u[k];
        }
      }
      if (tid == 2)
      {
        int n = 0;
        for (int f = 0; f < NK; f++)
          s.cidx[f] = (cmask & (1u << f)) ? n++ : -1;
        s.ncon = CD * n;
        s.Cx = reinterpret_cast<double(*)[NDX]>(io.knot + L.oC);
        s.Cu = reinterpret_cast<double(*)[NU]>(io.knot + L.oD);
        // constraint rows in the reference's insertion order (src/fulldynamics.cpp:145-211)
        int row = 0;
        if (O.torque_limits)
          for (int j = 0; j < NJNT; j++, row++)
            s.rkind[row] = FROW_TBOX, s.rarg[row] = j, s.rsub[row] = 0;
        if (O.kinematics_limits)
          for (int j = 0; j < NJNT; j++, row++)
            s.rkind[row] = FROW_QBOX, s.rarg[row] = j, s.rsub[row] = 0;
        for (int f = 0; f < NK; f++)
        {
          if (!(cmask & (1u << f)))
            continue;
          if (O.force_cone)
            for (int r = 0; r < Cfg::NCONE; r++, row++)
              s.rkind[row] = FROW_CONE, s.rarg[row] = f, s.rsub[row] = r;
          if (O.land_cstr && (slot.land_mask & (1u << f)))
          {
            // point feet: 3 LWA velocity rows + the height; quad feet: the 6-D LWA velocity (reference src/fulldynamics.cpp:175-209)
            for (int r = 0; r < (FS == 6 ? 6 : 3); r++, row++)
              s.rkind[row] = FROW_LVEL, s.rarg[row] = f, s.rsub[row] = r;
            if (FS == 3)
              s.rkind[row] = FROW_LZ, s.rarg[row] = f, s.rsub[row] = 2, row++;
          }
        }
        s.nc = row;
      }
      for (int e = tid; e < NV * NV; e += SMPC_NT)
        s.Mm[e / NV][e % NV] = 0.0;
    }
    SMPC_END

    kino_kinematics<Cfg>(M, s);
    kino_accel_pass<Cfg>(M, s, nullptr);
    kino_centroid<Cfg>(M, s);
    const int ncon = s.ncon;

    // ---- joint-space inertia, nonlinear effects, contact Jacobian and drift ----
    SMPC_PAR(tid)
    {
      for (int c = tid; c < NV; c += SMPC_NT)
      {
        const int i = M.joint_of_dof[c];
        double Jc6[6], F[6], Yg[6];
        for (int r = 0; r < 6; r++)
          Jc6[r] = s.J[r][c];
        inertia_apply(s.Y[i], Jc6, F);
        for (int j = 0; j <= c; j++)
          if ((M.anc_mask[i] >> M.joint_of_dof[j]) & 1u)
          {
            double sacc = 0;
            for (int r = 0; r < 6; r++)
              sacc += s.J[r][j] * F[r];
            s.Mm[j][c] = sacc;
            s.Mm[c][j] = sacc;
          }
        inertia_apply(s.Y[i], g6, Yg);
        double sacc = 0;
        for (int r = 0; r < 6; r++)
          sacc += Jc6[r] * (s.fs[i][r] - Yg[r]);
        s.nle[c] = sacc;
        s.tau[c] = (c < 6) ? 0.0 : s.u[c - 6];
      }
      for (int t = tid; t < NK * NV + NK; t += SMPC_NT)
      {
        if (t < NK * NV)
        {
          const int f = t / NV, c = t % NV;
          const int cf = s.cidx[f];
          if (cf < 0)
            continue;
          const int jf = d.foot_joint[f];
          const bool on = (M.anc_mask[jf] >> M.joint_of_dof[c]) & 1u;
          const double Jm[6] = {s.J[0][c], s.J[1][c], s.J[2][c], s.J[3][c], s.J[4][c], s.J[5][c]};
          if (FS == 6)
          {
            double Jl[6] = {0, 0, 0, 0, 0, 0};
            if (on)
              frame_motion_lwa(Jm, s.fp[f], Jl);
            for (int r = 0; r < CD; r++)
              s.Jc[CD * cf + r][c] = Jl[r];
            continue;
          }
          double l3[3] = {0, 0, 0};
          if (on)
          {
            double wxp[3], c3[3];
            v3cross(Jm + 3, s.fp[f], wxp);
            for (int r = 0; r < 3; r++)
              c3[r] = Jm[r] + wxp[r];
            m3tv(s.fR[f], c3, l3);
          }
          for (int r = 0; r < 3; r++)
            s.Jc[3 * cf + r][c] = l3[r];
        }
        else
        {
          const int f = t - NK * NV;
          const int cf = s.cidx[f];
          if (cf < 0)
            continue;
          const int jf = d.foot_joint[f];
          if (FS == 6)
          {
            double g6[6];
            full6_gamma<Cfg, DERIVS>(O, s, f, jf, g6);
            for (int r = 0; r < CD; r++)
              s.gamma[CD * cf + r] = g6[r];
            continue;
          }
          double pd[3], pdd[3], a1[3], a2[3], a3[3];
          point_va(s.v[jf], s.a[jf], s.fp[f], pd, pdd);
          m3tv(s.fR[f], pdd, a1);
          m3tv(s.fR[f], pd, a2);
          m3tv(s.fR[f], s.fp[f], a3);
          for (int r = 0; r < 3; r++)
            s.gamma[3 * cf + r] = a1[r] + O.Kd[r] * a2[r] + O.Kp[r] * a3[r];
        }
      }
    }
    SMPC_END

    coop_cholesky<NV>(s.Mm, s.dgM, NV, nullptr);

    // ---- M^-1 Jc^T (one thread per contact row) and the free acceleration ----
    SMPC_PAR(tid)
    for (int t = tid; t <= ncon; t += SMPC_NT)
    {
      if (t < ncon)
      {
        for (int k = 0; k < NV; k++)
          s.MiJt[k][t] = s.Jc[t][k];
        chol_solve_vec<NV>(s.Mm, NV, &s.MiJt[0][t], Cfg::NCON);
      }
      else
      {
        for (int k = 0; k < NV; k++)
          s.af[k] = s.tau[k] - s.nle[k];
        chol_solve_vec<NV>(s.Mm, NV, s.af, 1);
      }
    }
    SMPC_END
    // ---- regularised Delassus matrix and the right-hand side ----
    SMPC_PAR(tid)
    for (int e = tid; e < ncon * ncon + ncon; e += SMPC_NT)
    {
      if (e < ncon * ncon)
      {
        const int r = e / ncon, c = e % ncon;
        double sacc = 0;
        for (int k = 0; k < NV; k++)
          sacc += s.Jc[r][k] * s.MiJt[k][c];
        s.G[r][c] = sacc + (r == c ? 1e-10 : 0.0); // proximal mu, reference src/fulldynamics.cpp:39
      }
      else
      {
        const int r = e - ncon * ncon;
        double sacc = s.gamma[r];
        for (int k = 0; k < NV; k++)
          sacc += s.Jc[r][k] * s.af[k];
        s.rhs0[r] = -sacc;
      }
    }
    SMPC_END
    coop_cholesky<Cfg::NCON>(s.G, s.dgG, ncon, nullptr);
    SMPC_PAR(tid)
    if (tid == 0)
    {
      for (int r = 0; r < ncon; r++)
        s.lam[r] = 0.0;
      for (int it = 0; it < 2; it++)
      {
        double xv[Cfg::NCON];
        for (int r = 0; r < ncon; r++)
          xv[r] = s.rhs0[r] + 1e-10 * s.lam[r];
        chol_solve_vec<Cfg::NCON>(s.G, ncon, xv, 1);
        for (int r = 0; r < ncon; r++)
          s.lam[r] = xv[r];
      }
    }
    SMPC_END
    SMPC_PAR(tid)
    for (int c = tid; c < NV; c += SMPC_NT)
    {
      double sacc = s.af[c];
      for (int r = 0; r < ncon; r++)
        sacc += s.MiJt[c][r] * s.lam[r];
      s.acc[c] = sacc;
    }
    SMPC_END

    // ---- integrator, dynamics slack, state error ----
    SMPC_PAR(tid)
    {
      if (tid == 0)
      {
        const double * q = s.x;
        const double * v = s.x + NQ;
        double dq[NV];
        for (int k = 0; k < NV; k++)
        {
          const double vn = v[k] + dt * s.acc[k];
          s.xnext[NQ + k] = vn;
          dq[k] = dt * vn;
        }
        integrate_q(NV, q, dq, s.xnext);
        difference_q(NV, s.xn_ws, s.xnext, s.e);
        for (int k = 0; k < NV; k++)
          s.e[NV + k] = s.xnext[NQ + k] - s.xn_ws[NQ + k];
        if (DERIVS)
        {
          double E[9], ep[3];
          exp6(dq, E, ep);
          actinv6(E, ep, s.Jq6);
          Jexp6(dq, s.Jv6);
        }
      }
      else if (tid == 1)
      {
        difference_q(NV, slot.x_ref, s.x, s.rs);
        for (int k = 0; k < NV; k++)
          s.rs[NV + k] = s.x[NQ + k] - slot.x_ref[NQ + k];
        if (DERIVS)
          Jlog6(s.rs, s.Jl6);
      }
      for (int k = tid; k < NDX; k += SMPC_NT)
      {
        s.xdot[k] = (k < NV) ? s.x[NQ + k] : s.acc[k - NV];
        io.xdot[k] = s.xdot[k]; // MPC::getStateDerivative (reference src/mpc.cpp:346-352)
      }
      // MPC::getContactForces (reference src/mpc.cpp:354-380): contact force of every foot in its contact frame, zero
      // rows for swing feet
      if (io.fc)
        for (int k = tid; k < 3 * NK; k += SMPC_NT)
        {
          const int cf = s.cidx[k / 3];
          io.fc[k] = (cf >= 0) ? s.lam[CD * cf + k % 3] : 0.0; // linear part (reference src/mpc.cpp:369)
        }
    }
    SMPC_END

    if (DERIVS)
    {
      kino_accel_pass<Cfg>(M, s, s.acc);
      // ---- subtree forces with gravity and contact forces; momentum derivative and foot Jacobian columns ----
      SMPC_PAR(tid)
      {
        for (int i = tid; i < NJ; i += SMPC_NT)
        {
          double Yg[6], F[6];
          inertia_apply(s.Y[i], g6, Yg);
          for (int r = 0; r < 6; r++)
            F[r] = s.fs[i][r] - Yg[r];
          for (int f = 0; f < NK; f++)
          {
            const int cf = s.cidx[f];
            if (cf < 0 || !((M.anc_mask[d.foot_joint[f]] >> i) & 1u))
              continue;
            double fw[3], nw[3];
            if (FS == 6)
            {
              for (int r = 0; r < 3; r++)
                fw[r] = s.lam[CD * cf + r];
              v3cross(s.fp[f], fw, nw);
              for (int r = 0; r < 3; r++)
                nw[r] += s.lam[CD * cf + (FS == 6 ? 3 : 0) + r];
            }
            else
            {
              m3v(s.fR[f], &s.lam[3 * cf], fw);
              v3cross(s.fp[f], fw, nw);
            }
            for (int r = 0; r < 3; r++)
            {
              F[r] -= fw[r];
              F[3 + r] -= nw[r];
            }
          }
          for (int r = 0; r < 6; r++)
            s.Fs[i][r] = F[r];
        }
        for (int j = tid; j < NV; j += SMPC_NT)
        {
          const int i = M.joint_of_dof[j];
          const int par = d.parent[i];
          const double z6[6] = {0, 0, 0, 0, 0, 0};
          const double * vp = (par >= 0) ? s.v[par] : z6;
          double Jm[6], t1[6], w[6], t2[6], dh0[6], o1[6];
          for (int r = 0; r < 6; r++)
            Jm[r] = s.J[r][j];
          mxf(Jm, s.hs[i], t1);
          mxm(vp, Jm, w);
          inertia_apply(s.Y[i], w, t2);
          for (int r = 0; r < 6; r++)
            dh0[r] = t1[r] + t2[r];
          shift_to_com(s.com, dh0, o1);
          const double mass = s.Y[0][0];
          const double Jcj[3] = {s.Ag[0][j] / mass, s.Ag[1][j] / mass, s.Ag[2][j] / mass};
          double c1[3];
          v3cross(Jcj, s.h0, c1);
          for (int r = 0; r < 3; r++)
            o1[3 + r] -= c1[r];
          for (int r = 0; r < 6; r++)
            s.dh_dq[r][j] = o1[r];
          for (int f = 0; f < NK; f++)
          {
            double jp[6] = {0, 0, 0, 0, 0, 0};
            if ((M.anc_mask[d.foot_joint[f]] >> i) & 1u)
              frame_motion_lwa(Jm, s.fp[f], jp);
            for (int r = 0; r < FS; r++)
              s.Jp[f][r][j] = jp[r];
          }
        }
      }
      SMPC_END

      // ---- right-hand sides of the sensitivity systems: RNEA partials (pairs j, c) and contact-acceleration partials ----
      SMPC_PAR(tid)
      {
        for (int e = tid; e < NV * NV; e += SMPC_NT)
        {
          const int j = e / NV, c = e % NV;
          const int ic = M.joint_of_dof[c], ij = M.joint_of_dof[j];
          int is;
          bool transport = false;
          double vq = 0, vv = 0;
          bool related = true;
          if ((M.anc_mask[ic] >> ij) & 1u)
          {
            is = ic;
            transport = (ic != ij);
          }
          else if ((M.anc_mask[ij] >> ic) & 1u)
            is = ij;
          else
          {
            is = 0;
            related = false;
          }
          if (related)
          {
            const int par = d.parent[ic];
            const double z6[6] = {0, 0, 0, 0, 0, 0};
            const double * vp = (par >= 0) ? s.v[par] : z6;
            double apg[6], Jm[6], dVdq[6], dAdq[6], dAdv[6], t[6], vs[6];
            for (int r = 0; r < 6; r++)
            {
              apg[r] = ((par >= 0) ? s.a[par][r] : 0.0) - g6[r];
              Jm[r] = s.J[r][c];
              vs[r] = s.v[ic][r] + vp[r];
            }
            mxm(vp, Jm, dVdq);
            mxm(apg, Jm, dAdq);
            mxm(vp, dVdq, t);
            for (int r = 0; r < 6; r++)
              dAdq[r] += t[r];
            mxm(vs, Jm, dAdv);
            double Tq[6], Tv[6], u1[6], u2[6], u3[6];
            inertia_apply(s.Y[is], dAdq, u1);
            dinertia_apply(s.hs[is], s.dIo[is], dVdq, u2);
            mxf(dVdq, s.hs[is], u3);
            for (int r = 0; r < 6; r++)
              Tq[r] = u1[r] + u2[r] + u3[r];
            if (transport)
            {
              mxf(Jm, s.Fs[is], u1);
              for (int r = 0; r < 6; r++)
                Tq[r] += u1[r];
            }
            inertia_apply(s.Y[is], dAdv, u1);
            dinertia_apply(s.hs[is], s.dIo[is], Jm, u2);
            mxf(Jm, s.hs[is], u3);
            for (int r = 0; r < 6; r++)
              Tv[r] = u1[r] + u2[r] + u3[r];
            for (int r = 0; r < 6; r++)
            {
              vq += s.J[r][j] * Tq[r];
              vv += s.J[r][j] * Tv[r];
            }
          }
          if (FS == 6)
          {
            // forces held in world-aligned axes do not turn with the foot (see the header)
            for (int f = 0; f < NK; f++)
            {
              const int cf = s.cidx[f];
              const unsigned am = M.anc_mask[d.foot_joint[f]];
              if (cf < 0 || !((am >> ic) & 1u) || !((am >> ij) & 1u))
                continue;
              const double * lf = &s.lam[CD * cf];
              const double wc[3] = {s.J[3][c], s.J[4][c], s.J[5][c]};
              double g[3], pg[3], wt[3];
              v3cross(wc, lf, g);
              v3cross(s.fp[f], g, pg);
              v3cross(wc, lf + (FS == 6 ? 3 : 0), wt);
              for (int r = 0; r < 3; r++)
                vq += s.J[r][j] * g[r] + s.J[3 + r][j] * (pg[r] + wt[r]);
            }
          }
          s.X1[j][c] = -vq;
          s.X1[j][NV + c] = -vv;
          s.X1[j][2 * NV + c] = (j == c) ? 1.0 : 0.0;
        }
        for (int t = tid; t < NK * NV; t += SMPC_NT)
        {
          const int f = t / NV, c = t % NV;
          const int cf = s.cidx[f];
          if (cf < 0)
            continue;
          const int jf = d.foot_joint[f];
          const int ic = M.joint_of_dof[c];
          if (FS == 6)
          {
            double dq6[6] = {0, 0, 0, 0, 0, 0}, dv6[6] = {0, 0, 0, 0, 0, 0};
            if ((M.anc_mask[jf] >> ic) & 1u)
              full6_contact_derivs<Cfg>(M, O, s, f, jf, c, dq6, dv6);
            for (int r = 0; r < CD; r++)
            {
              s.DL[CD * cf + r][c] = -dq6[r];
              s.DL[CD * cf + r][NV + c] = -dv6[r];
              s.DL[CD * cf + r][2 * NV + c] = 0.0;
            }
            continue;
          }
          double dcq[3] = {0, 0, 0}, dcv[3] = {0, 0, 0};
          if ((M.anc_mask[jf] >> ic) & 1u)
          {
            const int par = d.parent[ic];
            const double z6[6] = {0, 0, 0, 0, 0, 0};
            const double * vp = (par >= 0) ? s.v[par] : z6;
            const double * ap = (par >= 0) ? s.a[par] : z6;
            const double * vi = s.v[jf];
            const double * ai = s.a[jf];
            const double * vim = s.v[ic];
            const double * p = s.fp[f];
            const double * Rf = s.fR[f];
            double Jm[6];
            for (int r = 0; r < 6; r++)
              Jm[r] = s.J[r][c];
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
            double jp[3], tt[3];
            v3cross(Jm + 3, p, tt);
            for (int r = 0; r < 3; r++)
              jp[r] = Jm[r] + tt[r];
            double dpd_q[3], dpdd_q[3], dpdd_v[3], t1[3], t2[3], t3[3], t4[3];
            v3cross(dv_q + 3, p, t1);
            v3cross(vi + 3, jp, t2);
            for (int r = 0; r < 3; r++)
              dpd_q[r] = dv_q[r] + t1[r] + t2[r];
            v3cross(da_q + 3, p, t1);
            v3cross(ai + 3, jp, t2);
            v3cross(dv_q + 3, pd, t3);
            v3cross(vi + 3, dpd_q, t4);
            for (int r = 0; r < 3; r++)
              dpdd_q[r] = da_q[r] + t1[r] + t2[r] + t3[r] + t4[r];
            v3cross(da_v + 3, p, t1);
            v3cross(Jm + 3, pd, t2);
            v3cross(vi + 3, jp, t3);
            for (int r = 0; r < 3; r++)
              dpdd_v[r] = da_v[r] + t1[r] + t2[r] + t3[r];
            double wa[3], wv[3], u1[3], u2[3], o1[3], o2[3], o3[3], o4[3];
            v3cross(Jm + 3, pdd, wa);
            v3cross(Jm + 3, pd, wv);
            for (int r = 0; r < 3; r++)
            {
              u1[r] = dpdd_q[r] - wa[r];
              u2[r] = dpd_q[r] - wv[r];
            }
            m3tv(Rf, u1, o1);
            m3tv(Rf, u2, o2);
            m3tv(Rf, Jm, o3);
            for (int r = 0; r < 3; r++)
              dcq[r] = o1[r] + O.Kd[r] * o2[r] + O.Kp[r] * o3[r];
            m3tv(Rf, dpdd_v, o1);
            m3tv(Rf, jp, o4);
            for (int r = 0; r < 3; r++)
              dcv[r] = o1[r] + O.Kd[r] * o4[r];
          }
          for (int r = 0; r < 3; r++)
          {
            s.DL[3 * cf + r][c] = -dcq[r];
            s.DL[3 * cf + r][NV + c] = -dcv[r];
            s.DL[3 * cf + r][2 * NV + c] = 0.0;
          }
        }
      }
      SMPC_END
      // ---- X1 = M^-1 r1 (one thread per column) ----
      SMPC_PAR(tid)
      for (int c = tid; c < NCOL; c += SMPC_NT)
        chol_solve_vec<NV>(s.Mm, NV, &s.X1[0][c], NCOL);
      SMPC_END
      // ---- DL = G^-1 (r2 - Jc X1) ----
      SMPC_PAR(tid)
      for (int e = tid; e < ncon * NCOL; e += SMPC_NT)
      {
        const int r = e / NCOL, c = e % NCOL;
        double sacc = 0;
        for (int k = 0; k < NV; k++)
          sacc += s.Jc[r][k] * s.X1[k][c];
        s.DL[r][c] = s.DL[r][c] - sacc;
      }
      SMPC_END
      SMPC_PAR(tid)
      for (int c = tid; c < NCOL; c += SMPC_NT)
        chol_solve_vec<Cfg::NCON>(s.G, ncon, &s.DL[0][c], NCOL);
      SMPC_END
      // ---- da = X1 + M^-1 Jc^T dlam (in place) ----
      SMPC_PAR(tid)
      for (int e = tid; e < NV * NCOL; e += SMPC_NT)
      {
        const int r = e / NCOL, c = e % NCOL;
        double sacc = s.X1[r][c];
        for (int k = 0; k < ncon; k++)
          sacc += s.MiJt[r][k] * s.DL[k][c];
        s.X1[r][c] = sacc;
      }
      SMPC_END

      // ---- A and B, one thread per column; weighted Jacobians of the residual costs ----
      SMPC_PAR(tid)
      {
        for (int c = tid; c < NZ; c += SMPC_NT)
        {
          const bool isA = c < NDX;
          const int j = isA ? c : c - NDX;
          const int rc = isA ? j : 2 * NV + 6 + j; // column of X1 holding d a / d(.)
          double * out = io.knot + (isA ? L.oA : L.oB);
          const int ld = isA ? NDX : NU;
          double dvn6[6];
          for (int r = 0; r < 6; r++)
            dvn6[r] = dt * s.X1[r][rc] + ((isA && j == NV + r) ? 1.0 : 0.0);
          for (int row = 0; row < NDX; row++)
          {
            double val;
            if (row < 6)
            {
              double sacc = 0;
              for (int k = 0; k < 6; k++)
                sacc += s.Jv6[6 * row + k] * dvn6[k];
              val = dt * sacc + ((isA && j < 6) ? s.Jq6[6 * row + j] : 0.0);
            }
            else if (row < NV)
            {
              const double dvn = dt * s.X1[row][rc] + ((isA && j == NV + row) ? 1.0 : 0.0);
              val = dt * dvn + ((isA && j == row) ? 1.0 : 0.0);
            }
            else
              val = dt * s.X1[row - NV][rc] + ((isA && j == row) ? 1.0 : 0.0);
            out[row * ld + j] = val;
          }
        }
        for (int e = tid; e < 6 * NDX; e += SMPC_NT)
        {
          const int r = e / NDX, a = e % NDX;
          double sacc = 0;
          for (int k = 0; k < 6; k++)
            sacc += O.w_cent[6 * r + k] * (a < NV ? s.dh_dq[k][a] : s.Ag[k][a - NV]);
          s.WJc[r][a] = sacc;
        }
        if (FS == 3)
          for (int e = tid; e < NK * 3 * NV; e += SMPC_NT)
          {
            const int f = e / (3 * NV), r = (e / NV) % 3, a = e % NV;
            double sacc = 0;
            for (int k = 0; k < 3; k++)
              sacc += O.w_frame[3 * r + k] * s.Jp[f][k][a];
            s.WJp[f][r][a] = sacc;
          }
        for (int e = tid; e < ncon * NZ; e += SMPC_NT)
        {
          const int row = e / NZ, a = e % NZ;
          const int cf = row / CD, r = row % CD;
          const int col = (a < NDX) ? a : 2 * NV + 6 + (a - NDX);
          double sacc = 0;
          for (int k = 0; k < CD; k++)
            sacc += O.w_forces[CD * r + k] * s.DL[CD * cf + k][col];
          s.WJl[row][a] = sacc;
        }
      }
      SMPC_END
    }

    // ---- cost values and weighted residuals ----
    SMPC_PAR(tid)
    {
      if (tid == 0)
      {
        double c0 = 0;
        for (int k = 0; k < NDX; k++)
          c0 += s.rs[k] * (O.w_x[k] * s.rs[k]);
        s.costs[0] = 0.5 * c0;
        double c1 = 0;
        for (int k = 0; k < NU; k++)
        {
          const double du = s.u[k] - slot.u_ref[k];
          c1 += du * (O.w_u[k] * du);
        }
        s.costs[1] = 0.5 * c1;
      }
      else if (tid == 1)
      {
        double c2 = 0;
        for (int r = 0; r < 6; r++)
        {
          double a = 0;
          for (int k = 0; k < 6; k++)
            a += O.w_cent[6 * r + k] * s.hg[k];
          s.Wch[r] = a;
          c2 += s.hg[r] * a;
        }
        s.costs[2] = 0.5 * c2;
      }
      if (tid >= 2 && tid < 2 + NK)
      {
        const int f = tid - 2;
        if (FS == 6)
          s.costs[3 + f] = full6_pose_cost<Cfg, DERIVS>(O, slot, s, f);
        else
        {
          double rf[3], cst = 0;
          for (int r = 0; r < 3; r++)
            rf[r] = s.fp[f][r] - slot.foot_ref[f][9 + r];
          for (int r = 0; r < 3; r++)
          {
            double a = 0;
            for (int k = 0; k < 3; k++)
              a += O.w_frame[3 * r + k] * rf[k];
            s.Wfr[f][r] = a;
            cst += rf[r] * a;
          }
          s.costs[3 + f] = 0.5 * cst;
        }
      }
      if (tid >= 2 + NK && tid < 2 + 2 * NK)
      {
        const int f = tid - 2 - NK;
        const int cf = s.cidx[f];
        double cst = 0;
        if (cf >= 0)
        {
          double rl[CD];
          for (int r = 0; r < CD; r++)
            rl[r] = s.lam[CD * cf + r] - slot.f_ref[f][r];
          for (int r = 0; r < CD; r++)
          {
            double a = 0;
            for (int k = 0; k < CD; k++)
              a += O.w_forces[CD * r + k] * rl[k];
            s.Wlr[CD * cf + r] = a;
            cst += rl[r] * a;
          }
        }
        s.costs[3 + NK + f] = 0.5 * cst;
      }
    }
    SMPC_END

    // ---- constraint values (and full Jacobian rows into the knot's C / D blocks) ----
    SMPC_PAR(tid)
    {
      if (tid == 0)
      {
        double cost = 0;
        for (int k = 0; k < 3 + 2 * NK; k++)
          cost += s.costs[k];
        s.cost = cost;
      }
      for (int row = tid; row < s.nc; row += SMPC_NT)
      {
        const int kind = s.rkind[row], f = s.rarg[row], sub = s.rsub[row];
        double c = 0, lo = 0, hi = 0;
        int type = 0;
        if (kind == FROW_TBOX)
        {
          c = s.u[f];
          lo = O.umin[f];
          hi = O.umax[f];
          type = 2;
        }
        else if (kind == FROW_QBOX)
        {
          c = s.x[7 + f];
          lo = O.qmin[f];
          hi = O.qmax[f];
          type = 2;
        }
        else if (kind == FROW_CONE)
        {
          const int cf = s.cidx[f];
          const double * Ac = (FS == 6) ? O.wcone : O.pyramid; // 17 x 6 wrench cone on [f; tau] / 5 x 3 friction pyramid
          for (int k = 0; k < CD; k++)
            c += Ac[CD * sub + k] * s.lam[CD * cf + k];
          type = 1;
        }
        else if (kind == FROW_LVEL)
        {
          const int jf = d.foot_joint[f];
          double V[6];
          frame_motion_lwa(s.v[jf], s.fp[f], V);
          c = V[sub];
        }
        else
          c = s.fp[f][2] - slot.land_z[f];
        s.cval[row] = c;
        s.lo[row] = lo;
        s.hi[row] = hi;
        s.ctype[row] = type;
      }
      if (DERIVS && FS == 6)
      {
        // pose cost of a quad foot: residual Jacobian T Jp (T from full6_pose_cost, previous phase); weighted: (T^T W T) Jp
        constexpr int Q36 = FullScratch<Cfg>::Q36;
        for (int e = tid; e < NK * FS * NV; e += SMPC_NT)
        {
          const int f = e / (FS * NV), r = (e / NV) % FS, a = e % NV;
          double sacc = 0;
          for (int i2 = 0; i2 < 6; i2++)
          {
            double wtj = 0; // (W T Jp)(i2, a)
            for (int j2 = 0; j2 < 6; j2++)
            {
              double tj = 0; // (T Jp)(j2, a)
              for (int k = 0; k < 6; k++)
                tj += s.Tpl[f][(6 * j2 + k) % Q36] * s.Jp[f][k % FS][a];
              wtj += O.w_frame[6 * i2 + j2] * tj;
            }
            sacc += s.Tpl[f][(6 * i2 + r) % Q36] * wtj;
          }
          s.WJp[f][r][a] = sacc;
        }
      }
      if (DERIVS)
      {
        for (int e = tid; e < s.nc * NZ; e += SMPC_NT)
        {
          const int row = e / NZ, a = e % NZ;
          const int kind = s.rkind[row], f = s.rarg[row], sub = s.rsub[row];
          double val = 0;
          if (kind == FROW_TBOX)
            val = (a == NDX + f) ? 1.0 : 0.0;
          else if (kind == FROW_QBOX)
            val = (a == 6 + f) ? 1.0 : 0.0;
          else if (kind == FROW_CONE)
          {
            const int cf = s.cidx[f];
            const int col = (a < NDX) ? a : 2 * NV + 6 + (a - NDX);
            const double * Ac = (FS == 6) ? O.wcone : O.pyramid;
            for (int k = 0; k < CD; k++)
              val += Ac[CD * sub + k] * s.DL[CD * cf + k][col];
          }
          else if (kind == FROW_LVEL)
          {
            if (a < NV)
            {
              const int jf = d.foot_joint[f];
              const int ia = M.joint_of_dof[a];
              if ((M.anc_mask[jf] >> ia) & 1u)
              {
                const int par = d.parent[ia];
                const double * vi = s.v[jf];
                double Jm[6], dvp[6], dvq[6], t1[3], t2[3];
                const double jp[3] = {s.Jp[f][0][a], s.Jp[f][1][a], s.Jp[f][2][a]};
                for (int r = 0; r < 6; r++)
                {
                  Jm[r] = s.J[r][a];
                  dvp[r] = ((par >= 0) ? s.v[par][r] : 0.0) - vi[r];
                }
                mxm(dvp, Jm, dvq);
                if (sub < 3)
                {
                  v3cross(dvq + 3, s.fp[f], t1);
                  v3cross(vi + 3, jp, t2);
                  val = dvq[sub] + t1[sub] + t2[sub];
                }
                else
                  val = dvq[sub]; // angular rows of the 6-D landing velocity
              }
            }
            else if (a < NDX)
              val = s.Jp[f][sub % FS][a - NV];
          }
          else
            val = (a < NV) ? s.Jp[f][2][a] : 0.0;
          if (a < NDX)
            s.Cx[row][a] = val;
          else
            s.Cu[row][a - NDX] = val;
        }
      }
    }
    SMPC_END

    if (DERIVS)
    {
      // ---- Gauss-Newton Q S R and the cost gradient ----
      SMPC_PAR(tid)
      {
        double * Q = io.knot + L.oQ;
        for (int e = tid; e < NDX * NDX; e += SMPC_NT)
        {
          const int a = e / NDX, b = e % NDX;
          double acc = 0;
          if (a < 6 && b < 6)
          {
            double t = 0;
            for (int r = 0; r < 6; r++)
              t += s.Jl6[6 * r + a] * (O.w_x[r] * s.Jl6[6 * r + b]);
            acc += t;
          }
          else if (a == b)
            acc += O.w_x[a];
          {
            double t = 0;
            for (int r = 0; r < 6; r++)
              t += (a < NV ? s.dh_dq[r][a] : s.Ag[r][a - NV]) * s.WJc[r][b];
            acc += t;
          }
          if (a < NV && b < NV)
            for (int f = 0; f < NK; f++)
            {
              double tf = 0;
              for (int r = 0; r < FS; r++)
                tf += s.Jp[f][r][a] * s.WJp[f][r][b];
              acc += tf;
            }
          {
            double t = 0;
            for (int r = 0; r < ncon; r++)
              t += s.DL[r][a] * s.WJl[r][b];
            acc += t;
          }
          Q[e] = acc;
        }
        double * S = io.knot + L.oS;
        for (int e = tid; e < NDX * NU; e += SMPC_NT)
        {
          const int a = e / NU, j = e % NU;
          double t = 0;
          for (int r = 0; r < ncon; r++)
            t += s.DL[r][a] * s.WJl[r][NDX + j];
          S[e] = t;
        }
        double * Rm = io.knot + L.oR;
        for (int e = tid; e < NU * NU; e += SMPC_NT)
        {
          const int i = e / NU, j = e % NU;
          double acc = (i == j) ? O.w_u[i] : 0.0;
          double t = 0;
          for (int r = 0; r < ncon; r++)
            t += s.DL[r][2 * NV + 6 + i] * s.WJl[r][NDX + j];
          Rm[e] = acc + t;
        }
        for (int a = tid; a < NDX; a += SMPC_NT)
        {
          double g = 0;
          if (a < 6)
          {
            for (int r = 0; r < 6; r++)
              g += s.Jl6[6 * r + a] * (O.w_x[r] * s.rs[r]);
          }
          else
            g = O.w_x[a] * s.rs[a];
          {
            double t = 0;
            for (int r = 0; r < 6; r++)
              t += (a < NV ? s.dh_dq[r][a] : s.Ag[r][a - NV]) * s.Wch[r];
            g += t;
          }
          if (a < NV)
            for (int f = 0; f < NK; f++)
            {
              double tf = 0;
              for (int r = 0; r < FS; r++)
                tf += s.Jp[f][r][a] * s.Wfr[f][r];
              g += tf;
            }
          {
            double t = 0;
            for (int r = 0; r < ncon; r++)
              t += s.DL[r][a] * s.Wlr[r];
            g += t;
          }
          s.lx[a] = g;
        }
        for (int j = tid; j < NU; j += SMPC_NT)
        {
          double g = O.w_u[j] * (s.u[j] - slot.u_ref[j]);
          double t = 0;
          for (int r = 0; r < ncon; r++)
            t += s.DL[r][2 * NV + 6 + j] * s.Wlr[r];
          s.lu[j] = g + t;
        }
      }
      SMPC_END
    }

    dense_knot_tail<FullScratch<Cfg>, DERIVS, false>(L, io, s);
  }

} // namespace smpc
