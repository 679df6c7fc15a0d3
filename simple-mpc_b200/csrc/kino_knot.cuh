// kino_knot.cuh — one horizon knot of the KINODYNAMICS problem, computed by one CTA:
//   values:       x+ = f(x,u), stage cost, constraint values, AL multiplier projections, merit pieces
//   linearisation: A, B, Gauss-Newton Q S R q r, active-set C D d  written once, in the dense knot
//                 layout the Riccati kernel streams (dev_types.h: KnotLayout)
//
// What it replaces (all [UPSTREAM] aligator/pinocchio code reached from the stage that
// KinodynamicsOCP::createStage builds, reference src/kinodynamics.cpp:40-152, terminal :352-377):
//   KinodynamicsFwdDynamics::forward/dForward, IntegratorSemiImplEuler, QuadraticStateCost,
//   QuadraticControlCost, CentroidalMomentumResidual, CentroidalMomentumDerivativeResidual,
//   FrameTranslationResidual, FrameVelocityResidual(LOCAL), CentroidalFrictionConeResidual,
//   StateErrorResidual slice (joint limits), DCMPositionResidual, and the ProxDDP per-stage work
//   (multiplier projection, LQ knot assembly, Lagrangian gradients).  SURVEY.md §8a rows a4,a5,a8,a9,a10.
//
// Algebra: everything in the WORLD frame about the world origin with compact 10-parameter inertias
// (dev_math.cuh); the centroidal-dynamics derivatives are the closed forms
//   dh0/dq_j   = J_j x* hs_i + Y_i (v_p x J_j)
//   dhd0/dv_j  = dY_i J_j + Y_i ((v_p + v_i) x J_j) + J_j x* hs_i
//   dhd0/dq_j  = J_j x* fs_i + Y_i (a_p x J_j) + dY_i w - Y_i (w x v_p) + w x* hs_i ,  w = v_p x J_j
// (i = joint of dof j, p = its parent, Y/hs/fs/dY = subtree sums), shifted to the centre of mass.
//
// Work distribution: joints of one tree level in parallel, then one thread per dof column for every
// 6 x nv / 3 x nv Jacobian, one thread per column of A/B (rows written coalesced), flat-index loops
// for Q/S/R/C/D. State that crosses a phase lives in KinoScratch (shared memory).
#pragma once
#include <cstddef>
#include "dev_math.cuh"
#include "dev_types.h"

namespace smpc
{

  template <int NJ_, int NV_, int NK_, int FS_>
  struct KinoCfg
  {
    static constexpr int NJ = NJ_, NV = NV_, NK = NK_, FS = FS_;
    static constexpr int NQ = NV + 1, NX = NQ + NV, NDX = 2 * NV, NF = FS * NK, NJNT = NV - 6, NU = NJNT + NF;
    static constexpr int NRHS = 2 * NV + NU;
    static constexpr int VR = (FS == 6) ? 6 : 3; // frame-velocity rows per foot
    static constexpr int PR = FS;                  // rows of the <foot>_pose_cost residual: translation / log6 placement
    static constexpr int NCMAX = NJNT + NK * (((FS == 6) ? 17 : 2) + VR + ((FS == 3) ? 1 : 0));
  };

  enum RowKind
  {
    ROW_BOX = 0,
    ROW_CONE0 = 1,
    ROW_CONE1 = 2,
    ROW_VEL = 3,
    ROW_LAND = 4,
    ROW_DCM = 5,
    ROW_WCONE = 6 // one of the 17 rows of the wrench cone of a quad foot (rsub = row)
  };

  template <class Cfg>
  struct KinoScratch
  {
    static constexpr int NJ = Cfg::NJ, NV = Cfg::NV, NK = Cfg::NK, NX = Cfg::NX, NU = Cfg::NU, NDX = Cfg::NDX;
    double x[NX], u[NU > 0 ? NU : 1], xn_ws[NX];
    // Shared memory bounds the resident knots per SM (one warp per knot; profiles/r01e, r01f: throughput follows the
    // resident warps), so arrays whose lifetimes do not overlap share storage.
    //   U1, three eras: (1) kinematics, acceleration passes and the per-dof derivative phase; (2) the right-hand sides
    //   of the implicit differentiation, consumed by the A / B assembly; (3) weighted Jacobians and constraint rows.
    union
    {
      struct
      {
        double R[NJ][9], p[NJ][3], v[NJ][6], a[NJ][6];
        double Y[NJ][10], hs[NJ][6], dIo[NJ][6], fs[NJ][6]; // subtree sums
        double J[6][NV];
      };
      struct
      {
        double rhs[6][Cfg::NRHS];
        double lamn[NDX]; // co-state lams[t+1], staged for the A / B assembly loops
        double Jq6[36], Jv6[36]; // free-flyer blocks of dIntegrate (built in the right-hand-side phase, read by the A / B assembly)
      };
      struct
      {
        double WJc[6][NDX], WJcdx[6][NV], WJcdu[6][Cfg::NF], WJp[NK][Cfg::PR][NV];
        double cval[Cfg::NCMAX], dproj[Cfg::NCMAX], vplus[Cfg::NCMAX], rpinf[Cfg::NCMAX], rcrit[Cfg::NCMAX];
        double vsm[Cfg::NCMAX]; // constraint multipliers of this knot, staged for the gradient loops
      };
    };
    //   U2: own-body inertia / momentum (until the last acceleration pass) | d(hdot)/d(q, v) (per-dof phase until the
    //   right-hand sides are built) | Lagrangian-gradient vectors (written by the A / B assembly and later)
    union
    {
      struct
      {
        double Yb[NJ][10], hb[NJ][6];
      };
      struct
      {
        double dhd_dq[6][NV], dhd_dv[6][NV];
      };
      struct
      {
        double Lx[NDX], Lu[NU > 0 ? NU : 1], gx[NDX], gu[NU > 0 ? NU : 1];
      };
    };
    static constexpr int Q6 = (Cfg::FS == 6) ? 6 : 1, Q36 = (Cfg::FS == 6) ? 36 : 1; // sizes of the quad-foot-only arrays
    double Wch[6], Wcdh[6], Wfr[NK][Cfg::PR]; // weighted residuals (written while era-1 arrays are still read at the terminal node)
    double mass;                          // Y[0][0], kept past era 1
    double fR[NK][9], fp[NK][3], lever[NK][3], vf[NK][6];
    double com[3], vcom[3], h0[6], hd0[6], hg[6], hdg[6], hdF[6];
    double Ag[6][NV];
    double acc[NV];
    double lu[36];
    int piv[6];
    double rs[NDX], xnext[NX], e[NDX], lplus[NDX];
    // constraint rows
    int nc;
    // row table as bytes (kind < 7, joint / foot index < 64, sub-row < 17): shared memory bounds the resident knots per SM
    unsigned char rkind[Cfg::NCMAX], rarg[Cfg::NCMAX], rsub[Cfg::NCMAX], ractive[Cfg::NCMAX];
    double costs[4 + NK];
    double rpl[NK][Q6]; // quad feet: pose residual log6(ref^-1 oMf) (written by the value-only instantiation too)
    // ---- linearisation only (DERIVS = true). Everything from dh_dq on is never touched by the value-only (trial)
    // instantiation, which is launched with trial_bytes() of shared memory and therefore fits more CTAs per SM ----
    double dh_dq[6][NV];
    double Jp[NK][Cfg::PR][NV];       // world-aligned foot Jacobian: translation rows (+ rotation rows for quad feet)
    double Tpl[NK][Q36];              // quad feet: T = Jlog6(r) blkdiag(R^T, R^T) of the pose residual r
    // d(LOCAL foot velocity)/dq. The partial w.r.t. v is R^T Jp (vf_dv below) and d(hdot_F)/du is [e_k ; lever x e_k]
    // (dhF_du below): both are rebuilt where they are read - 2.3 KB of scratch less, which is what lets 12 knots share an SM
    double vf_dq[NK][Cfg::VR][NV];
    double dhF_dq[3][NV];
    double Jl6[36];
  };
  template <class Cfg>
  constexpr size_t kino_trial_bytes()
  {
    return (offsetof(KinoScratch<Cfg>, dh_dq) + 15) & ~(size_t)15;
  }

  // Global-memory view of one knot.
  struct KnotIO
  {
    const double * x;      // xs[t]
    const double * u;      // us[t]          (NULL at the terminal node)
    const double * xn;     // xs[t+1]        (NULL at the terminal node)
    // trial mode: x <- x (+) alpha dx, u <- u + alpha du, xn <- xn (+) alpha dxn ; results to tx, tu
    const double * dx, *du, *dxn;
    double alpha;
    double * tx, *tu;
    const smpc_slot * slot;
    const smpc_terminal * term;
    const double * vs, *prev_vs;         // multipliers of this knot's constraints
    const double * lam_t, *lam_n;        // co-states lams[t], lams[t+1]
    const double * prev_lam_n;           // estimate for lams[t+1]
    double mu, preg;
    // outputs
    double * knot; // dense LQ knot (KnotLayout)
    int * nc_out;
    double * gx, *gu;   // merit gradient pieces (first-order multipliers), without the -lam_plus[t] term
    double * lplus_n;   // lam_plus[t+1] = prev_lam[t+1] + e/mu
    double * xdot;
    double * fc = nullptr; // contact forces [nfeet][3] (full dynamics only)
    KnotScalars * sc;
  };

  // d(LOCAL velocity of foot f, row r)/dv_a = (R_f^T Jp_f)(r, a): translation rows from the point Jacobian, rotation rows
  // (quad feet) from the world-aligned angular Jacobian kept in rows 3..5 of Jp
  template <class Cfg>
  SMPC_DEV double kino_vf_dv(const KinoScratch<Cfg> & s, int f, int r, int a)
  {
    constexpr int PR = Cfg::PR;
    const int r3 = r % 3, k0 = (r >= 3) ? 3 : 0;
    const double * R = s.fR[f];
    return R[r3] * s.Jp[f][k0 % PR][a] + R[3 + r3] * s.Jp[f][(k0 + 1) % PR][a] + R[6 + r3] * s.Jp[f][(k0 + 2) % PR][a];
  }
  // d(hdot_F)(r)/du_jf for the contact-wrench entry jf = FS f + k: [e_k ; lever_f x e_k] for a force, [0 ; e_k] for a torque
  template <class Cfg>
  SMPC_DEV double kino_dhF_du(const KinoScratch<Cfg> & s, unsigned cmask, int r, int jf)
  {
    constexpr int FS = Cfg::FS;
    const int f = jf / FS, k = jf % FS;
    if (!(cmask & (1u << f)))
      return 0.0;
    if (k >= 3 || r < 3)
      return (r == k) ? 1.0 : 0.0;
    const double * l = s.lever[f];
    const int i = r - 3; // (l x e_k)(i)
    if (i == k)
      return 0.0;
    const int m = 3 - i - k; // the remaining axis
    return ((k - i + 3) % 3 == 1) ? -l[m] : l[m];
  }

  // ------------------------------------------------------------------------------------------
  // kinematics + centroidal quantities shared by stage and terminal knots
  // ------------------------------------------------------------------------------------------
  template <class Cfg, class S>
  SMPC_DEV void kino_kinematics(const DevModel & M, S & s)
  {
    constexpr int NV = Cfg::NV, NQ = Cfg::NQ;
    const smpc_model_desc & d = M.d;
    const double * q = s.x;
    const double * v = s.x + NQ;
    for (int l = 0; l < M.nlevels; l++)
    {
      SMPC_PAR(tid)
      for (int idx = M.level_start[l] + tid; idx < M.level_start[l + 1]; idx += SMPC_NT)
      {
        const int i = M.level_joints[idx];
        const int par = d.parent[i];
        const int iv = d.idx_v[i];
        const bool ff = d.jtype[i] == SMPC_JOINT_FREEFLYER;
        double Rj[9], pj[3] = {0, 0, 0};
        if (ff)
        {
          quat_to_R(q + d.idx_q[i] + 3, Rj);
          pj[0] = q[d.idx_q[i]];
          pj[1] = q[d.idx_q[i] + 1];
          pj[2] = q[d.idx_q[i] + 2];
        }
        else
        {
          const double ang = q[d.idx_q[i]];
          const double sn = sin(ang), cs = cos(ang);
          double W[9], W2[9];
          skew(d.axis[i], W);
          m3m(W, W, W2);
          for (int k = 0; k < 9; k++)
            Rj[k] = ((k % 4 == 0) ? 1.0 : 0.0) + sn * W[k] + (1 - cs) * W2[k];
        }
        double Rl[9], pl[3];
        m3m(d.place_R[i], Rj, Rl);
        m3v(d.place_R[i], pj, pl);
        for (int k = 0; k < 3; k++)
          pl[k] += d.place_p[i][k];
        double * oR = s.R[i];
        double * op = s.p[i];
        if (par < 0)
        {
          for (int k = 0; k < 9; k++)
            oR[k] = Rl[k];
          for (int k = 0; k < 3; k++)
            op[k] = pl[k];
        }
        else
        {
          m3m(s.R[par], Rl, oR);
          double t[3];
          m3v(s.R[par], pl, t);
          for (int k = 0; k < 3; k++)
            op[k] = t[k] + s.p[par][k];
        }
        // Jacobian columns and joint velocity
        double vJ[6] = {0, 0, 0, 0, 0, 0};
        if (ff)
        {
          for (int c = 0; c < 6; c++)
          {
            double col[6];
            const double axc[3] = {oR[c % 3], oR[3 + c % 3], oR[6 + c % 3]};
            if (c < 3)
            {
              col[0] = axc[0], col[1] = axc[1], col[2] = axc[2];
              col[3] = col[4] = col[5] = 0;
            }
            else
            {
              v3cross(op, axc, col);
              col[3] = axc[0], col[4] = axc[1], col[5] = axc[2];
            }
            for (int r = 0; r < 6; r++)
            {
              s.J[r][iv + c] = col[r];
              vJ[r] += col[r] * v[iv + c];
            }
          }
        }
        else
        {
          double col[6];
          m3v(oR, d.axis[i], col + 3);
          v3cross(op, col + 3, col);
          for (int r = 0; r < 6; r++)
          {
            s.J[r][iv] = col[r];
            vJ[r] = col[r] * v[iv];
          }
        }
        double * vi = s.v[i];
        for (int r = 0; r < 6; r++)
          vi[r] = (par >= 0 ? s.v[par][r] : 0.0) + vJ[r];
        // world-frame inertia about the origin
        double cw[3], T[9], Ic[9];
        m3v(oR, d.lever[i], cw);
        for (int k = 0; k < 3; k++)
          cw[k] += op[k];
        m3m(oR, d.inertia[i], T);
        m3mt(T, oR, Ic);
        const double m = d.mass[i];
        const double c2 = v3dot(cw, cw);
        double * Y = s.Yb[i];
        Y[0] = m;
        Y[1] = m * cw[0], Y[2] = m * cw[1], Y[3] = m * cw[2];
        Y[4] = Ic[0] + m * (c2 - cw[0] * cw[0]);
        Y[5] = Ic[1] - m * cw[0] * cw[1];
        Y[6] = Ic[2] - m * cw[0] * cw[2];
        Y[7] = Ic[4] + m * (c2 - cw[1] * cw[1]);
        Y[8] = Ic[5] - m * cw[1] * cw[2];
        Y[9] = Ic[8] + m * (c2 - cw[2] * cw[2]);
        inertia_apply(Y, vi, s.hb[i]);
        // dIo = W Io + (W Io)^T - mc vl^T - vl mc^T + 2 (vl . mc) I
        {
          const double S9[9] = {Y[4], Y[5], Y[6], Y[5], Y[7], Y[8], Y[6], Y[8], Y[9]};
          double W[9], WS[9];
          skew(vi + 3, W);
          m3m(W, S9, WS);
          const double * mc = Y + 1;
          const double vm = v3dot(vi, mc);
          double * D = s.dIo[i];
          D[0] = 2 * WS[0] - 2 * mc[0] * vi[0] + 2 * vm;
          D[1] = WS[1] + WS[3] - mc[0] * vi[1] - vi[0] * mc[1];
          D[2] = WS[2] + WS[6] - mc[0] * vi[2] - vi[0] * mc[2];
          D[3] = 2 * WS[4] - 2 * mc[1] * vi[1] + 2 * vm;
          D[4] = WS[5] + WS[7] - mc[1] * vi[2] - vi[1] * mc[2];
          D[5] = 2 * WS[8] - 2 * mc[2] * vi[2] + 2 * vm;
        }
        for (int k = 0; k < 10; k++)
          s.Y[i][k] = Y[k];
        for (int k = 0; k < 6; k++)
          s.hs[i][k] = s.hb[i][k];
      }
      SMPC_END
    }
    // subtree sums of Y (10), hs (6), dIo (6): one thread per component, children before parents
    SMPC_PAR(tid)
    if (tid < 22)
    {
      for (int i = d.njoints - 1; i > 0; i--)
      {
        const int par = d.parent[i];
        if (tid < 10)
          s.Y[par][tid] += s.Y[i][tid];
        else if (tid < 16)
          s.hs[par][tid - 10] += s.hs[i][tid - 10];
        else
          s.dIo[par][tid - 16] += s.dIo[i][tid - 16];
      }
    }
    SMPC_END
    (void)NV;
  }

  // spatial accelerations (gravity-free) and inertial forces for joint accelerations `acc` (NULL = 0),
  // then the subtree force sums fs.
  template <class Cfg, class S>
  SMPC_DEV void kino_accel_pass(const DevModel & M, S & s, const double * acc)
  {
    const smpc_model_desc & d = M.d;
    for (int l = 0; l < M.nlevels; l++)
    {
      SMPC_PAR(tid)
      for (int idx = M.level_start[l] + tid; idx < M.level_start[l + 1]; idx += SMPC_NT)
      {
        const int i = M.level_joints[idx];
        const int par = d.parent[i];
        const int iv = d.idx_v[i];
        const int ndof = (d.jtype[i] == SMPC_JOINT_FREEFLYER) ? 6 : 1;
        double * ai = s.a[i];
        // velocity-product term v_i x (S_i qdot_i), rebuilt from the Jacobian columns (it used to be kept in scratch)
        const double * vq = s.x + Cfg::NQ;
        double vJ[6], bias[6];
        for (int r = 0; r < 6; r++)
        {
          double t = 0;
          for (int c = 0; c < ndof; c++)
            t += s.J[r][iv + c] * vq[iv + c];
          vJ[r] = t;
        }
        mxm(s.v[i], vJ, bias);
        for (int r = 0; r < 6; r++)
        {
          double aJ = 0;
          if (acc)
            for (int c = 0; c < ndof; c++)
              aJ += s.J[r][iv + c] * acc[iv + c];
          ai[r] = (par >= 0 ? s.a[par][r] : 0.0) + aJ + bias[r];
        }
        double Ia[6], vxh[6];
        inertia_apply(s.Yb[i], ai, Ia);
        mxf(s.v[i], s.hb[i], vxh);
        for (int r = 0; r < 6; r++)
          s.fs[i][r] = Ia[r] + vxh[r];
      }
      SMPC_END
    }
    SMPC_PAR(tid)
    if (tid < 6)
      for (int i = d.njoints - 1; i > 0; i--)
        s.fs[d.parent[i]][tid] += s.fs[i][tid];
    SMPC_END
  }

  SMPC_DEV void shift_to_com(const double * com, const double * in, double * out)
  {
    double cxf[3];
    v3cross(com, in, cxf);
    for (int r = 0; r < 3; r++)
    {
      out[r] = in[r];
      out[3 + r] = in[3 + r] - cxf[r];
    }
  }

  // com, momenta, Ag columns, foot placements / velocities
  template <class Cfg, class S>
  SMPC_DEV void kino_centroid(const DevModel & M, S & s)
  {
    constexpr int NV = Cfg::NV, NK = Cfg::NK;
    const smpc_model_desc & d = M.d;
    SMPC_PAR(tid)
    if (tid == 0)
    {
      const double mass = s.Y[0][0];
      for (int r = 0; r < 3; r++)
        s.com[r] = s.Y[0][1 + r] / mass;
      for (int r = 0; r < 6; r++)
      {
        s.h0[r] = s.hs[0][r];
        s.hd0[r] = s.fs[0][r];
      }
      for (int r = 0; r < 3; r++)
        s.vcom[r] = s.h0[r] / mass;
      shift_to_com(s.com, s.h0, s.hg);
      shift_to_com(s.com, s.hd0, s.hdg);
    }
    SMPC_END
    SMPC_PAR(tid)
    {
      if (tid < NV)
      {
        const int i = M.joint_of_dof[tid];
        double Jc[6], A0[6], Ag[6];
        for (int r = 0; r < 6; r++)
          Jc[r] = s.J[r][tid];
        inertia_apply(s.Y[i], Jc, A0);
        shift_to_com(s.com, A0, Ag);
        for (int r = 0; r < 6; r++)
          s.Ag[r][tid] = Ag[r];
      }
      else if (tid - NV < NK)
      {
        const int f = tid - NV;
        const int jf = d.foot_joint[f];
        m3m(s.R[jf], d.foot_R[f], s.fR[f]);
        double t[3];
        m3v(s.R[jf], d.foot_p[f], t);
        for (int r = 0; r < 3; r++)
        {
          s.fp[f][r] = t[r] + s.p[jf][r];
          s.lever[f][r] = s.fp[f][r] - s.com[r];
        }
        double wxp[3], vl[3];
        v3cross(s.v[jf] + 3, s.fp[f], wxp);
        for (int r = 0; r < 3; r++)
          vl[r] = s.v[jf][r] + wxp[r];
        m3tv(s.fR[f], vl, s.vf[f]);
        m3tv(s.fR[f], s.v[jf] + 3, s.vf[f] + 3);
      }
    }
    SMPC_END
  }

  // 6x6 LU with partial pivoting, in place (one thread)
  SMPC_DEV void lu6_factor(double * A, int * piv)
  {
    for (int k = 0; k < 6; k++)
    {
      int p = k;
      double best = fabs(A[6 * k + k]);
      for (int i = k + 1; i < 6; i++)
        if (fabs(A[6 * i + k]) > best)
        {
          best = fabs(A[6 * i + k]);
          p = i;
        }
      piv[k] = p;
      if (p != k)
        for (int j = 0; j < 6; j++)
        {
          const double t = A[6 * k + j];
          A[6 * k + j] = A[6 * p + j];
          A[6 * p + j] = t;
        }
      const double inv = 1.0 / A[6 * k + k];
      for (int i = k + 1; i < 6; i++)
      {
        const double l = A[6 * i + k] * inv;
        A[6 * i + k] = l;
        for (int j = k + 1; j < 6; j++)
          A[6 * i + j] -= l * A[6 * k + j];
      }
    }
  }
  SMPC_DEV void lu6_solve(const double * A, const int * piv, double * b)
  {
    for (int k = 0; k < 6; k++)
    {
      const int p = piv[k];
      if (p != k)
      {
        const double t = b[k];
        b[k] = b[p];
        b[p] = t;
      }
      for (int i = k + 1; i < 6; i++)
        b[i] -= A[6 * i + k] * b[k];
    }
    for (int k = 5; k >= 0; k--)
    {
      double sacc = b[k];
      for (int i = k + 1; i < 6; i++)
        sacc -= A[6 * k + i] * b[i];
      b[k] = sacc / A[6 * k + k];
    }
  }

  SMPC_DEV double ncone_proj(int type, double z, double lo, double hi, int & active)
  {
    if (type == 0) // equality
    {
      active = 1;
      return z;
    }
    if (type == 1) // negative orthant
    {
      active = z > 0;
      return active ? z : 0.0;
    }
    if (z < lo)
    {
      active = 1;
      return z - lo;
    }
    if (z > hi)
    {
      active = 1;
      return z - hi;
    }
    active = 0;
    return 0.0;
  }

  // <foot>_pose_cost of a quad foot: FramePlacementResidual r = log6(ref^-1 oMf), weight w_frame (6 x 6). The residual
  // Jacobian is T Jp with T = Jlog6(r) blkdiag(R^T, R^T) and Jp the world-aligned foot Jacobian; stores r, T and the
  // gradient weights Wfr = T^T W r. Returns 1/2 r^T W r. S = KinoScratch or FullScratch (reference src/kinodynamics.cpp:67-73,
  // src/fulldynamics.cpp:99-104).
  template <class Cfg, bool DERIVS, class S>
  SMPC_DEV double quad_pose_cost(const DevOcp & O, const smpc_slot & slot, S & s, int f)
  {
    constexpr int Q6 = S::Q6, Q36 = S::Q36, FS = Cfg::FS;
    const double * Rr = slot.foot_ref[f];
    double Rrel[9], dp[3], prel[3], rr[6], Wr[6], cst = 0;
    m3tm(Rr, s.fR[f], Rrel);
    for (int r = 0; r < 3; r++)
      dp[r] = s.fp[f][r] - slot.foot_ref[f][9 + r];
    m3tv(Rr, dp, prel);
    log6(Rrel, prel, rr);
    for (int r = 0; r < 6; r++)
    {
      s.rpl[f][r % Q6] = rr[r];
      double a = 0;
      for (int k = 0; k < 6; k++)
        a += O.w_frame[6 * r + k] * rr[k];
      Wr[r] = a;
      cst += rr[r] * a;
    }
    if (DERIVS)
    {
      double Jl[36], T[36];
      Jlog6(rr, Jl);
      for (int r = 0; r < 6; r++)
        for (int half = 0; half < 2; half++)
          for (int k = 0; k < 3; k++)
          {
            double a = 0;
            for (int m2 = 0; m2 < 3; m2++)
              a += Jl[6 * r + 3 * half + m2] * s.fR[f][3 * k + m2];
            T[6 * r + 3 * half + k] = a;
          }
      for (int r = 0; r < 36; r++)
        s.Tpl[f][r % Q36] = T[r];
      for (int k = 0; k < 6; k++)
      {
        double a = 0;
        for (int r = 0; r < 6; r++)
          a += T[6 * r + k] * Wr[r];
        s.Wfr[f][k % FS] = a;
      }
    }
    return 0.5 * cst;
  }

  // ------------------------------------------------------------------------------------------
  // stage knot
  // ------------------------------------------------------------------------------------------
  template <class Cfg, bool DERIVS>
  SMPC_DEV void kino_stage_knot(const DevModel & M, const DevOcp & O, const KnotLayout & L, const KnotIO & io, KinoScratch<Cfg> & s)
  {
    constexpr int NV = Cfg::NV, NQ = Cfg::NQ, NX = Cfg::NX, NDX = Cfg::NDX, NU = Cfg::NU, NK = Cfg::NK, FS = Cfg::FS;
    constexpr int NF = Cfg::NF, NJNT = Cfg::NJNT, NRHS = Cfg::NRHS, VR = Cfg::VR;
    static_assert(FS == 3 || FS == 6, "point feet (3-D forces) or quad feet (6-D wrenches)");
    constexpr int PR = Cfg::PR;
    const smpc_slot & slot = *io.slot;
    const unsigned cmask = slot.contact_mask;
    const double dt = O.dt, mu = io.mu, inv_mu = 1.0 / io.mu;

    // ---- load (and, for a trial, step) the iterate ----
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
    }
    SMPC_END

    kino_kinematics<Cfg>(M, s);
    kino_accel_pass<Cfg>(M, s, nullptr);
    kino_centroid<Cfg>(M, s);

    // ---- forward dynamics value, integrator, state error (independent serial tasks on separate threads) ----
    SMPC_PAR(tid)
    {
      if (tid == 0)
      {
        const double * u = s.u;
        const double * aj = s.u + NF;
        double hdF[6] = {O.mass * O.gravity[0], O.mass * O.gravity[1], O.mass * O.gravity[2], 0, 0, 0};
        for (int f = 0; f < NK; f++)
          if (cmask & (1u << f))
          {
            double t[3];
            v3cross(s.lever[f], u + FS * f, t);
            for (int r = 0; r < 3; r++)
            {
              hdF[r] += u[FS * f + r];
              hdF[3 + r] += t[r];
              if (FS == 6)
                hdF[3 + r] += u[FS * f + 3 + r]; // contact torque of a quad foot
            }
          }
        double b[6];
        for (int r = 0; r < 6; r++)
        {
          s.hdF[r] = hdF[r];
          double sacc = hdF[r] - s.hdg[r];
          for (int j = 0; j < NJNT; j++)
            sacc -= s.Ag[r][6 + j] * aj[j];
          b[r] = sacc;
          for (int c = 0; c < 6; c++)
            s.lu[6 * r + c] = s.Ag[r][c];
        }
        lu6_factor(s.lu, s.piv);
        lu6_solve(s.lu, s.piv, b);
        for (int r = 0; r < 6; r++)
          s.acc[r] = b[r];
        for (int j = 0; j < NJNT; j++)
          s.acc[6 + j] = aj[j];
        // semi-implicit Euler
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
        // dynamics slack e = x+ (-) xs[t+1], first-order co-state lam_plus = lam_e + e/mu
        difference_q(NV, s.xn_ws, s.xnext, s.e);
        for (int k = 0; k < NV; k++)
          s.e[NV + k] = s.xnext[NQ + k] - s.xn_ws[NQ + k];
        for (int k = 0; k < NDX; k++)
          s.lplus[k] = io.prev_lam_n[k] + s.e[k] * inv_mu;
      }
      else if (tid == 1)
      {
        difference_q(NV, slot.x_ref, s.x, s.rs);
        for (int k = 0; k < NV; k++)
          s.rs[NV + k] = s.x[NQ + k] - slot.x_ref[NQ + k];
        if (DERIVS)
          Jlog6(s.rs, s.Jl6);
      }
      else if (tid == 2)
      {
        // constraint row table, in the reference's insertion order (src/kinodynamics.cpp:91-149)
        int row = 0;
        if (O.kinematics_limits)
          for (int j = 0; j < NJNT; j++, row++)
          {
            s.rkind[row] = ROW_BOX;
            s.rarg[row] = j;
            s.rsub[row] = 0;
          }
        for (int f = 0; f < NK; f++)
        {
          if (!(cmask & (1u << f)))
            continue;
          if (O.force_cone)
          {
            if (FS == 6)
            {
              // CentroidalWrenchConeResidual: 17 rows A [f; tau] <= 0 (reference src/kinodynamics.cpp:104-111)
              for (int r = 0; r < 17; r++, row++)
                s.rkind[row] = ROW_WCONE, s.rarg[row] = f, s.rsub[row] = r;
            }
            else
            {
              s.rkind[row] = ROW_CONE0, s.rarg[row] = f, s.rsub[row] = 0, row++;
              s.rkind[row] = ROW_CONE1, s.rarg[row] = f, s.rsub[row] = 0, row++;
            }
          }
          for (int r = 0; r < VR; r++, row++)
          {
            s.rkind[row] = ROW_VEL;
            s.rarg[row] = f;
            s.rsub[row] = r;
          }
          if (FS == 3 && O.land_cstr && (slot.land_mask & (1u << f)))
          {
            s.rkind[row] = ROW_LAND, s.rarg[row] = f, s.rsub[row] = 2, row++;
          }
        }
        s.nc = row;
      }
    }
    SMPC_END

    if (DERIVS)
    {
      kino_accel_pass<Cfg>(M, s, s.acc);
      SMPC_PAR(tid)
      if (tid == 0)
        for (int r = 0; r < 6; r++)
          s.hd0[r] = s.fs[0][r];
      SMPC_END

      // ---- one thread per dof column: centroidal derivatives, foot Jacobians, frame-velocity partials ----
      SMPC_PAR(tid)
      if (tid < NV)
      {
        const int j = tid;
        const smpc_model_desc & d = M.d;
        const int i = M.joint_of_dof[j];
        const int par = d.parent[i];
        const double z6[6] = {0, 0, 0, 0, 0, 0};
        const double * vp = (par >= 0) ? s.v[par] : z6;
        const double * ap = (par >= 0) ? s.a[par] : z6;
        const double * vi = s.v[i];
        const double * Y = s.Y[i];
        const double * hs = s.hs[i];
        const double * fsub = s.fs[i];
        const double * dmc = s.hs[i]; // d(m c)/dt of the subtree = its linear momentum
        const double * dIo = s.dIo[i];
        double Jc[6];
        for (int r = 0; r < 6; r++)
          Jc[r] = s.J[r][j];
        double t1[6], w[6], t2[6], dh0[6];
        mxf(Jc, hs, t1);
        mxm(vp, Jc, w);
        inertia_apply(Y, w, t2);
        for (int r = 0; r < 6; r++)
          dh0[r] = t1[r] + t2[r];
        double vpi[6], t3[6], t4[6], t5[6], dhdv0[6];
        for (int r = 0; r < 6; r++)
          vpi[r] = vp[r] + vi[r];
        dinertia_apply(dmc, dIo, Jc, t3);
        mxm(vpi, Jc, t4);
        inertia_apply(Y, t4, t5);
        for (int r = 0; r < 6; r++)
          dhdv0[r] = t3[r] + t5[r] + t1[r];
        double u1[6], u2[6], u3[6], u4[6], u5[6], u6[6], u7[6], dhdq0[6];
        mxf(Jc, fsub, u1);
        mxm(ap, Jc, u2);
        inertia_apply(Y, u2, u3);
        dinertia_apply(dmc, dIo, w, u4);
        mxm(w, vp, u5);
        inertia_apply(Y, u5, u6);
        mxf(w, hs, u7);
        for (int r = 0; r < 6; r++)
          dhdq0[r] = u1[r] + u3[r] + u4[r] - u6[r] + u7[r];
        double o1[6], o2[6], o3[6];
        shift_to_com(s.com, dh0, o1);
        shift_to_com(s.com, dhdq0, o2);
        shift_to_com(s.com, dhdv0, o3);
        const double mass = s.Y[0][0];
        const double Jcj[3] = {s.Ag[0][j] / mass, s.Ag[1][j] / mass, s.Ag[2][j] / mass};
        double c1[3], c2[3];
        v3cross(Jcj, s.h0, c1);
        v3cross(Jcj, s.hd0, c2);
        for (int r = 0; r < 3; r++)
        {
          o1[3 + r] -= c1[r];
          o2[3 + r] -= c2[r];
        }
        for (int r = 0; r < 6; r++)
        {
          s.dh_dq[r][j] = o1[r];
          s.dhd_dq[r][j] = o2[r];
          s.dhd_dv[r][j] = o3[r];
        }
        double dF[3] = {0, 0, 0};
        for (int f = 0; f < NK; f++)
        {
          const int jf = d.foot_joint[f];
          const bool on = (M.anc_mask[jf] >> i) & 1u;
          double jp[3] = {0, 0, 0}, wl[3] = {0, 0, 0};
          if (on)
          {
            double t[3];
            v3cross(Jc + 3, s.fp[f], t);
            for (int r = 0; r < 3; r++)
              jp[r] = Jc[r] + t[r];
            v3cross(w + 3, s.fp[f], t);
            for (int r = 0; r < 3; r++)
              wl[r] = w[r] + t[r];
          }
          double a3[3], b3[3];
          m3tv(s.fR[f], wl, a3);
          m3tv(s.fR[f], jp, b3);
          for (int r = 0; r < 3; r++)
          {
            s.Jp[f][r][j] = jp[r];
            s.vf_dq[f][r][j] = a3[r];
          }
          if (FS == 6)
          {
            // rotation rows: world-aligned Jacobian J_ang, LOCAL angular velocity partials R^T (v_p x J_j)_ang, R^T J_ang
            const double z3[3] = {0, 0, 0};
            m3tv(s.fR[f], on ? w + 3 : z3, a3);
            m3tv(s.fR[f], on ? Jc + 3 : z3, b3);
            for (int r = 0; r < 3; r++)
            {
              s.Jp[f][(3 + r) % PR][j] = on ? Jc[3 + r] : 0.0;
              s.vf_dq[f][(3 + r) % VR][j] = a3[r];
            }
          }
          if (cmask & (1u << f))
          {
            const double dl[3] = {jp[0] - Jcj[0], jp[1] - Jcj[1], jp[2] - Jcj[2]};
            double t[3];
            v3cross(dl, s.u + FS * f, t);
            for (int r = 0; r < 3; r++)
              dF[r] += t[r];
          }
        }
        for (int r = 0; r < 3; r++)
          s.dhF_dq[r][j] = dF[r];
      }
      SMPC_END

      // ---- implicit differentiation: Agb [da_dq | da_dv | da_du] = [dhF_dq - dhd_dq | -dhd_dv | dhF_du | -Ag_joint] ----
      SMPC_PAR(tid)
      for (int c = tid; c < NRHS; c += SMPC_NT)
      {
        double b[6];
        if (c < NV)
        {
          for (int r = 0; r < 3; r++)
          {
            b[r] = -s.dhd_dq[r][c];
            b[3 + r] = s.dhF_dq[r][c] - s.dhd_dq[3 + r][c];
          }
        }
        else if (c < 2 * NV)
        {
          for (int r = 0; r < 6; r++)
            b[r] = -s.dhd_dv[r][c - NV];
        }
        else if (c < 2 * NV + NF)
        {
          const int jf = c - 2 * NV;
          const int f = jf / FS, k = jf % FS;
          for (int r = 0; r < 6; r++)
            b[r] = 0;
          if (cmask & (1u << f))
          {
            // d/df_k [f ; lever x f] = [e_k ; lever x e_k] ; contact torque of a quad foot: d/dtau_k = [0 ; e_k]
            b[k] = 1.0;
            if (k < 3)
            {
              double ek[3] = {0, 0, 0}, t[3];
              ek[k] = 1.0;
              v3cross(s.lever[f], ek, t);
              b[3] = t[0], b[4] = t[1], b[5] = t[2];
            }
          }
        }
        else
        {
          const int jj = c - 2 * NV - NF;
          for (int r = 0; r < 6; r++)
            b[r] = -s.Ag[r][6 + jj];
        }
        lu6_solve(s.lu, s.piv, b);
        for (int r = 0; r < 6; r++)
          s.rhs[r][c] = b[r];
      }
      // multipliers read inside the A / B assembly loops: one coalesced fetch instead of a global load per iteration
      for (int k = tid; k < NDX; k += SMPC_NT)
        s.lamn[k] = io.lam_n[k];
      // dIntegrate blocks of q+ = q (+) dt v+ on the free-flyer: by the lane with the fewest right-hand sides
      if (tid == SMPC_NT - 1)
      {
        double dq6[6], E[9], ep[3];
        for (int k = 0; k < 6; k++)
          dq6[k] = dt * s.xnext[NQ + k];
        exp6(dq6, E, ep);
        actinv6(E, ep, s.Jq6);
        Jexp6(dq6, s.Jv6);
      }
      SMPC_END

      // ---- A and B, one thread per column, rows written coalesced; Lagrangian-gradient columns on the fly ----
      SMPC_PAR(tid)
      for (int c = tid; c < NDX + NU; c += SMPC_NT)
      {
        const bool isA = c < NDX;
        const int j = isA ? c : c - NDX;
        const int rc = isA ? j : 2 * NV + j; // column of rhs holding d a_base / d(.)
        double dvn[6];
        for (int r = 0; r < 6; r++)
          dvn[r] = dt * s.rhs[r][rc] + ((isA && j == NV + r) ? 1.0 : 0.0);
        double * out = io.knot + (isA ? L.oA : L.oB);
        const int ld = isA ? NDX : NU;
        double Lacc = 0, gacc = 0;
        for (int row = 0; row < NDX; row++)
        {
          double val;
          if (row < 6)
          {
            double sacc = 0;
            for (int k = 0; k < 6; k++)
              sacc += s.Jv6[6 * row + k] * dvn[k];
            val = dt * sacc + ((isA && j < 6) ? s.Jq6[6 * row + j] : 0.0);
          }
          else if (row < NV)
          {
            // joint rows of q+: dt * dvn(row, .) + identity
            const double dv = isA ? ((j == NV + row) ? 1.0 : 0.0) : ((j == NF + row - 6) ? dt : 0.0);
            val = dt * dv + ((isA && j == row) ? 1.0 : 0.0);
          }
          else if (row < NV + 6)
            val = dvn[row - NV];
          else
          {
            const int r = row - NV;
            val = isA ? ((j == NV + r) ? 1.0 : 0.0) : ((j == NF + r - 6) ? dt : 0.0);
          }
          out[row * ld + j] = val;
          Lacc += val * s.lamn[row];
          gacc += val * s.lplus[row];
        }
        if (isA)
        {
          s.Lx[j] = Lacc;
          s.gx[j] = gacc;
        }
        else
        {
          s.Lu[j] = Lacc;
          s.gu[j] = gacc;
        }
      }
      SMPC_END

      // ---- weighted Jacobians of the residual costs ----
      SMPC_PAR(tid)
      {
        for (int e = tid; e < 6 * NDX; e += SMPC_NT)
        {
          const int r = e / NDX, a = e % NDX;
          double sacc = 0;
          for (int k = 0; k < 6; k++)
            sacc += O.w_cent[6 * r + k] * (a < NV ? s.dh_dq[k][a] : s.Ag[k][a - NV]);
          s.WJc[r][a] = sacc;
        }
        for (int e = tid; e < 6 * NV; e += SMPC_NT)
        {
          const int r = e / NV, a = e % NV;
          double sacc = 0;
          for (int k = 0; k < 3; k++)
            sacc += O.w_centder[6 * r + 3 + k] * s.dhF_dq[k][a];
          s.WJcdx[r][a] = sacc;
        }
        for (int e = tid; e < 6 * NF; e += SMPC_NT)
        {
          const int r = e / NF, a = e % NF;
          double sacc = 0;
          for (int k = 0; k < 6; k++)
            sacc += O.w_centder[6 * r + k] * kino_dhF_du<Cfg>(s, cmask, k, a);
          s.WJcdu[r][a] = sacc;
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
        double c2 = 0, c3 = 0;
        for (int r = 0; r < 6; r++)
        {
          double a = 0, b = 0;
          for (int k = 0; k < 6; k++)
          {
            a += O.w_cent[6 * r + k] * s.hg[k];
            b += O.w_centder[6 * r + k] * s.hdF[k];
          }
          s.Wch[r] = a;
          s.Wcdh[r] = b;
          c2 += s.hg[r] * a;
          c3 += s.hdF[r] * b;
        }
        s.costs[2] = 0.5 * c2;
        s.costs[3] = 0.5 * c3;
      }
      if (tid >= 2 && tid < 2 + NK)
      {
        const int f = tid - 2;
        if (FS == 6)
          s.costs[4 + f] = quad_pose_cost<Cfg, DERIVS>(O, slot, s, f);
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
          s.costs[4 + f] = 0.5 * cst;
        }
      }
    }
    SMPC_END

    // ---- constraint values, multiplier projection ----
    SMPC_PAR(tid)
    {
    if (DERIVS && FS == 6)
    {
      // pose cost of a quad foot: residual Jacobian T Jp (T from kino6_pose_cost, previous phase); weighted: (T^T W T) Jp
      constexpr int Q36 = KinoScratch<Cfg>::Q36;
      for (int e = tid; e < NK * PR * NV; e += SMPC_NT)
      {
        const int f = e / (PR * NV), r = (e / NV) % PR, a = e % NV;
        double sacc = 0;
        for (int i2 = 0; i2 < 6; i2++)
        {
          double wtj = 0; // (W T Jp)(i2, a)
          for (int j2 = 0; j2 < 6; j2++)
          {
            double tj = 0; // (T Jp)(j2, a)
            for (int k = 0; k < 6; k++)
              tj += s.Tpl[f][(6 * j2 + k) % Q36] * s.Jp[f][k % PR][a];
            wtj += O.w_frame[6 * i2 + j2] * tj;
          }
          sacc += s.Tpl[f][(6 * i2 + r) % Q36] * wtj;
        }
        s.WJp[f][r][a] = sacc;
      }
    }
    if (DERIVS)
      for (int k = tid; k < O.ncmax; k += SMPC_NT) // the multiplier arrays are strided by the run-time ncmax
        s.vsm[k] = io.vs[k];
    for (int row = tid; row < s.nc; row += SMPC_NT)
    {
      const int kind = s.rkind[row], f = s.rarg[row];
      double c = 0, lo = 0, hi = 0;
      int type = 0;
      if (kind == ROW_BOX)
      {
        c = s.x[7 + f];
        lo = O.qmin[f];
        hi = O.qmax[f];
        type = 2;
      }
      else if (FS == 6 && kind == ROW_WCONE)
      {
        const double * Ac = O.wcone + 6 * s.rsub[row];
        for (int k = 0; k < 6; k++)
          c += Ac[k] * s.u[FS * f + (k % FS)];
        type = 1;
      }
      else if (kind == ROW_CONE0)
      {
        c = -s.u[FS * f + 2] + 1e-4; // reference src/kinodynamics.cpp:127
        type = 1;
      }
      else if (kind == ROW_CONE1)
      {
        const double * ff = s.u + FS * f;
        c = -O.mu * O.mu * ff[2] * ff[2] + ff[0] * ff[0] + ff[1] * ff[1];
        type = 1;
      }
      else if (kind == ROW_VEL)
        c = s.vf[f][s.rsub[row]];
      else
        c = s.fp[f][2] - slot.land_z[f];
      const double z = c + mu * io.prev_vs[row];
      int act;
      const double pr = ncone_proj(type, z, lo, hi, act);
      s.cval[row] = c;
      s.dproj[row] = pr;
      s.ractive[row] = act;
      s.vplus[row] = pr * inv_mu;
      s.rpinf[row] = fabs(pr - mu * io.prev_vs[row]);
      s.rcrit[row] = fabs(pr - mu * io.vs[row]);
    }
    }
    SMPC_END

    if (DERIVS)
    {
      // ---- Q S R (Gauss-Newton), q r, flat-index, coalesced ----
      SMPC_PAR(tid)
      {
        double * Q = io.knot + L.oQ;
        for (int e = tid; e < NDX * NDX; e += SMPC_NT)
        {
          const int a = e / NDX, b = e % NDX;
          double acc = 0;
          // state_cost
          if (a < 6 && b < 6)
          {
            double t = 0;
            for (int r = 0; r < 6; r++)
              t += s.Jl6[6 * r + a] * (O.w_x[r] * s.Jl6[6 * r + b]);
            acc += t;
          }
          else if (a == b)
            acc += O.w_x[a];
          // centroidal_cost
          {
            double t = 0;
            for (int r = 0; r < 6; r++)
              t += (a < NV ? s.dh_dq[r][a] : s.Ag[r][a - NV]) * s.WJc[r][b];
            acc += t;
          }
          if (a < NV && b < NV)
          {
            // centroidal_derivative_cost (angular rows depend on q)
            double t = 0;
            for (int r = 0; r < 3; r++)
              t += s.dhF_dq[r][a] * s.WJcdx[3 + r][b];
            acc += t;
            // <foot>_pose_cost
            for (int f = 0; f < NK; f++)
            {
              double tf = 0;
              for (int r = 0; r < PR; r++)
                tf += s.Jp[f][r][a] * s.WJp[f][r][b];
              acc += tf;
            }
          }
          if (a == b)
            acc += io.preg;
          Q[e] = acc;
        }
        double * S = io.knot + L.oS;
        for (int e = tid; e < NDX * NU; e += SMPC_NT)
        {
          const int a = e / NU, j = e % NU;
          double acc = 0;
          if (a < NV && j < NF)
            for (int r = 0; r < 3; r++)
              acc += s.dhF_dq[r][a] * s.WJcdu[3 + r][j];
          S[e] = acc;
        }
        double * Rm = io.knot + L.oR;
        for (int e = tid; e < NU * NU; e += SMPC_NT)
        {
          const int i = e / NU, j = e % NU;
          double acc = (i == j) ? O.w_u[i] : 0.0;
          if (i < NF && j < NF)
          {
            double t = 0;
            for (int r = 0; r < 6; r++)
              t += kino_dhF_du<Cfg>(s, cmask, r, i) * s.WJcdu[r][j];
            acc += t;
          }
          if (i == j)
            acc += io.preg;
          Rm[e] = acc;
        }
        // gradients
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
          {
            double t = 0;
            for (int r = 0; r < 3; r++)
              t += s.dhF_dq[r][a] * s.Wcdh[3 + r];
            g += t;
            for (int f = 0; f < NK; f++)
            {
              double tf = 0;
              for (int r = 0; r < PR; r++)
                tf += s.Jp[f][r][a] * s.Wfr[f][r];
              g += tf;
            }
          }
          io.knot[L.oq + a] = g;
          // constraint Jacobian transposed times multipliers (all rows, current and first-order multipliers)
          double Lc = 0, gc = 0;
          for (int row = 0; row < s.nc; row++)
          {
            const int kind = s.rkind[row], f = s.rarg[row];
            double cx = 0;
            if (kind == ROW_BOX)
              cx = (a == 6 + f) ? 1.0 : 0.0;
            else if (kind == ROW_VEL)
              cx = (a < NV) ? s.vf_dq[f][s.rsub[row]][a] : kino_vf_dv<Cfg>(s, f, s.rsub[row], a - NV);
            else if (kind == ROW_LAND)
              cx = (a < NV) ? s.Jp[f][2][a] : 0.0;
            Lc += cx * s.vsm[row];
            gc += cx * s.vplus[row];
          }
          s.Lx[a] = g + s.Lx[a] + Lc - io.lam_t[a];
          s.gx[a] = g + s.gx[a] + gc;
          io.knot[L.of + a] = s.e[a] + mu * io.prev_lam_n[a];
        }
        for (int j = tid; j < NU; j += SMPC_NT)
        {
          double g = O.w_u[j] * (s.u[j] - slot.u_ref[j]);
          if (j < NF)
          {
            double t = 0;
            for (int r = 0; r < 6; r++)
              t += kino_dhF_du<Cfg>(s, cmask, r, j) * s.Wcdh[r];
            g += t;
          }
          io.knot[L.or_ + j] = g;
          double Lc = 0, gc = 0;
          for (int row = 0; row < s.nc; row++)
          {
            const int kind = s.rkind[row], f = s.rarg[row];
            double cu = 0;
            if (FS == 6 && kind == ROW_WCONE)
              cu = (j >= FS * f && j < FS * f + FS) ? O.wcone[6 * s.rsub[row] + (j - FS * f) % 6] : 0.0;
            else if (kind == ROW_CONE0)
              cu = (j == FS * f + 2) ? -1.0 : 0.0;
            else if (kind == ROW_CONE1)
            {
              if (j == FS * f)
                cu = 2 * s.u[FS * f];
              else if (j == FS * f + 1)
                cu = 2 * s.u[FS * f + 1];
              else if (j == FS * f + 2)
                cu = -2 * O.mu * O.mu * s.u[FS * f + 2];
            }
            Lc += cu * s.vsm[row];
            gc += cu * s.vplus[row];
          }
          s.Lu[j] = g + s.Lu[j] + Lc;
          s.gu[j] = g + s.gu[j] + gc;
        }
        // active-set constraint rows
        double * Cm = io.knot + L.oC;
        for (int e = tid; e < L.ncmax * NDX; e += SMPC_NT)
        {
          const int row = e / NDX, a = e % NDX;
          double cx = 0;
          if (row < s.nc && s.ractive[row])
          {
            const int kind = s.rkind[row], f = s.rarg[row];
            if (kind == ROW_BOX)
              cx = (a == 6 + f) ? 1.0 : 0.0;
            else if (kind == ROW_VEL)
              cx = (a < NV) ? s.vf_dq[f][s.rsub[row]][a] : kino_vf_dv<Cfg>(s, f, s.rsub[row], a - NV);
            else if (kind == ROW_LAND)
              cx = (a < NV) ? s.Jp[f][2][a] : 0.0;
          }
          Cm[e] = cx;
        }
        double * Dm = io.knot + L.oD;
        for (int e = tid; e < L.ncmax * NU; e += SMPC_NT)
        {
          const int row = e / NU, j = e % NU;
          double cu = 0;
          if (row < s.nc && s.ractive[row])
          {
            const int kind = s.rkind[row], f = s.rarg[row];
            if (FS == 6 && kind == ROW_WCONE)
              cu = (j >= FS * f && j < FS * f + FS) ? O.wcone[6 * s.rsub[row] + (j - FS * f) % 6] : 0.0;
            else if (kind == ROW_CONE0)
              cu = (j == FS * f + 2) ? -1.0 : 0.0;
            else if (kind == ROW_CONE1)
            {
              if (j == FS * f)
                cu = 2 * s.u[FS * f];
              else if (j == FS * f + 1)
                cu = 2 * s.u[FS * f + 1];
              else if (j == FS * f + 2)
                cu = -2 * O.mu * O.mu * s.u[FS * f + 2];
            }
          }
          Dm[e] = cu;
        }
        for (int row = tid; row < L.ncmax; row += SMPC_NT)
          io.knot[L.od + row] = (row < s.nc) ? s.dproj[row] : 0.0;
      }
      SMPC_END
    }

    // ---- per-knot scalars and small vectors ----
    SMPC_PAR(tid)
    {
      if (tid == 0)
      {
        double cost = 0;
        for (int k = 0; k < 4 + NK; k++)
          cost += s.costs[k];
        double sq = 0, pinf = 0, crit = 0;
        for (int k = 0; k < NDX; k++)
        {
          sq += s.lplus[k] * s.lplus[k];
          pinf = fmax(pinf, fabs(s.e[k]));
          crit = fmax(crit, fabs(mu * (s.lplus[k] - io.lam_n[k])));
        }
        double pen = 0.5 * mu * sq;
        double sqv = 0;
        for (int row = 0; row < s.nc; row++)
        {
          sqv += s.vplus[row] * s.vplus[row];
          pinf = fmax(pinf, s.rpinf[row]);
          crit = fmax(crit, s.rcrit[row]);
        }
        pen += 0.5 * mu * sqv;
        double dinf = 0;
        if (DERIVS)
        {
          // the x-stationarity row of node 0 defines lam_0 (dx_0 = 0 is imposed): skipped by the caller via sc->pad
          for (int k = 0; k < NDX; k++)
            dinf = fmax(dinf, fabs(s.Lx[k]));
          double dinfu = 0;
          for (int k = 0; k < NU; k++)
            dinfu = fmax(dinfu, fabs(s.Lu[k]));
          io.sc->pad[0] = dinf;  // |Lx|
          io.sc->pad[1] = dinfu; // |Lu|
          dinf = fmax(dinf, dinfu);
        }
        io.sc->cost = cost;
        io.sc->pen = pen;
        io.sc->pinf = pinf;
        io.sc->dinf = dinf;
        io.sc->crit = crit;
        *io.nc_out = s.nc;
      }
      for (int k = tid; k < NDX; k += SMPC_NT)
      {
        io.xdot[k] = (k < NV) ? s.x[NQ + k] : s.acc[k - NV];
        io.lplus_n[k] = s.lplus[k];
        if (DERIVS)
          io.gx[k] = s.gx[k];
      }
      if (DERIVS)
        for (int k = tid; k < NU; k += SMPC_NT)
          io.gu[k] = s.gu[k];
    }
    SMPC_END
  }

  // ------------------------------------------------------------------------------------------
  // terminal knot: state_cost + 10 x centroidal_cost, optional DCM equality
  // (reference src/kinodynamics.cpp:352-377)
  // ------------------------------------------------------------------------------------------
  template <class Cfg, bool DERIVS>
  SMPC_DEV void kino_terminal_knot(const DevModel & M, const DevOcp & O, const KnotLayout & L, const KnotIO & io, KinoScratch<Cfg> & s)
  {
    constexpr int NV = Cfg::NV, NQ = Cfg::NQ, NX = Cfg::NX, NDX = Cfg::NDX;
    const smpc_terminal & tm = *io.term;
    const double mu = io.mu, inv_mu = 1.0 / io.mu;
    SMPC_PAR(tid)
    {
      if (io.dx == nullptr)
      {
        for (int k = tid; k < NX; k += SMPC_NT)
          s.x[k] = io.x[k];
      }
      else if (tid == 0)
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
    }
    SMPC_END
    kino_kinematics<Cfg>(M, s);
    kino_accel_pass<Cfg>(M, s, nullptr);
    kino_centroid<Cfg>(M, s);
    const int nct = O.nct;
    SMPC_PAR(tid)
    {
      if (tid == 0)
      {
        difference_q(NV, tm.x_ref, s.x, s.rs);
        for (int k = 0; k < NV; k++)
          s.rs[NV + k] = s.x[NQ + k] - tm.x_ref[NQ + k];
        if (DERIVS)
          Jlog6(s.rs, s.Jl6);
        double c0 = 0;
        for (int k = 0; k < NDX; k++)
          c0 += s.rs[k] * (O.w_x[k] * s.rs[k]);
        double c2 = 0;
        for (int r = 0; r < 6; r++)
        {
          double a = 0;
          for (int k = 0; k < 6; k++)
            a += (10.0 * O.w_cent[6 * r + k]) * s.hg[k]; // reference src/kinodynamics.cpp:361
          s.Wch[r] = a;
          c2 += s.hg[r] * a;
        }
        s.costs[0] = 0.5 * c0 + 0.5 * c2;
        s.nc = nct;
        s.mass = s.Y[0][0]; // the subtree sums are overwritten by the weighted Jacobians below
      }
      if (DERIVS && tid < NV)
      {
        // dh/dq column (only the momentum derivative is needed at the terminal node)
        const int j = tid;
        const smpc_model_desc & d = M.d;
        const int i = M.joint_of_dof[j];
        const int par = d.parent[i];
        const double z6[6] = {0, 0, 0, 0, 0, 0};
        const double * vp = (par >= 0) ? s.v[par] : z6;
        double Jc[6], t1[6], w[6], t2[6], dh0[6], o1[6];
        for (int r = 0; r < 6; r++)
          Jc[r] = s.J[r][j];
        mxf(Jc, s.hs[i], t1);
        mxm(vp, Jc, w);
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
      }
    }
    SMPC_END
    // DCM rows: c + tau cdot - ref
    const double tau = tm.tau; // fixed at createTerminalConstraint (reference src/kinodynamics.cpp:372); updateTerminalConstraint only moves com_ref
    SMPC_PAR(tid)
    {
      for (int row = tid; row < nct; row += SMPC_NT)
      {
        const double c = s.com[row] + tau * s.vcom[row] - tm.com_ref[row];
        const double z = c + mu * io.prev_vs[row];
        s.cval[row] = c;
        s.dproj[row] = z;
        s.ractive[row] = 1;
        s.vplus[row] = z * inv_mu;
        s.rpinf[row] = fabs(z - mu * io.prev_vs[row]);
        s.rcrit[row] = fabs(z - mu * io.vs[row]);
      }
      if (DERIVS)
        for (int e = tid; e < 6 * NDX; e += SMPC_NT)
        {
          const int r = e / NDX, a = e % NDX;
          double sacc = 0;
          for (int k = 0; k < 6; k++)
            sacc += (10.0 * O.w_cent[6 * r + k]) * (a < NV ? s.dh_dq[k][a] : s.Ag[k][a - NV]);
          s.WJc[r][a] = sacc;
        }
    }
    SMPC_END
    if (DERIVS)
    {
      SMPC_PAR(tid)
      {
        const double mass = s.mass;
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
          double t = 0;
          for (int r = 0; r < 6; r++)
            t += (a < NV ? s.dh_dq[r][a] : s.Ag[r][a - NV]) * s.WJc[r][b];
          acc += t;
          if (a == b)
            acc += io.preg;
          Q[e] = acc;
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
          double t = 0;
          for (int r = 0; r < 6; r++)
            t += (a < NV ? s.dh_dq[r][a] : s.Ag[r][a - NV]) * s.Wch[r];
          g += t;
          io.knot[L.oq + a] = g;
          double Lc = 0, gc = 0;
          for (int row = 0; row < nct; row++)
          {
            const double cx = (a < NV) ? (s.Ag[row][a] / mass + tau * s.dh_dq[row][a] / mass) : tau * (s.Ag[row][a - NV] / mass);
            Lc += cx * io.vs[row];
            gc += cx * s.vplus[row];
          }
          s.Lx[a] = g + Lc - io.lam_t[a];
          s.gx[a] = g + gc;
        }
        double * Cm = io.knot + L.oC;
        for (int e = tid; e < L.ncmax * NDX; e += SMPC_NT)
        {
          const int row = e / NDX, a = e % NDX;
          double cx = 0;
          if (row < nct)
            cx = (a < NV) ? (s.Ag[row][a] / mass + tau * s.dh_dq[row][a] / mass) : tau * (s.Ag[row][a - NV] / mass);
          Cm[e] = cx;
        }
        for (int row = tid; row < L.ncmax; row += SMPC_NT)
          io.knot[L.od + row] = (row < nct) ? s.dproj[row] : 0.0;
      }
      SMPC_END
    }
    SMPC_PAR(tid)
    {
      if (tid == 0)
      {
        double pen = 0, pinf = 0, crit = 0, sqv = 0;
        for (int row = 0; row < nct; row++)
        {
          sqv += s.vplus[row] * s.vplus[row];
          pinf = fmax(pinf, s.rpinf[row]);
          crit = fmax(crit, s.rcrit[row]);
        }
        pen = 0.5 * mu * sqv;
        double dinf = 0;
        if (DERIVS)
          for (int k = 0; k < NDX; k++)
            dinf = fmax(dinf, fabs(s.Lx[k]));
        io.sc->cost = s.costs[0];
        io.sc->pen = pen;
        io.sc->pinf = pinf;
        io.sc->dinf = dinf;
        io.sc->crit = crit;
        io.sc->pad[0] = dinf;
        io.sc->pad[1] = 0;
        *io.nc_out = nct;
      }
      if (DERIVS)
        for (int k = tid; k < NDX; k += SMPC_NT)
          io.gx[k] = s.gx[k];
    }
    SMPC_END
  }

} // namespace smpc
