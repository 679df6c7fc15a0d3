// centro_knot.cuh — one horizon knot of the CENTROIDAL problem, one warp per knot.
//
// What it replaces (all [UPSTREAM] aligator code reached from the stage that CentroidalOCP::createStage
// builds, reference src/centroidal-dynamics.cpp:39-106, terminal cost :306-316):
//   CentroidalFwdDynamics + IntegratorEuler, CentroidalCoMResidual, ControlErrorResidual,
//   LinearMomentumResidual, AngularMomentumResidual, CentroidalAccelerationResidual,
//   AngularAccelerationResidual, CentroidalWrenchConeResidual / CentroidalFrictionConeResidual (eps = 1e-4,
//   reference src/centroidal-dynamics.cpp:99), and the per-stage ProxDDP work (dense_knot.cuh).
//
//   x = [c; h_lin; h_ang] in R^9, u = [f_0 .. f_{nk-1}] (fs each); contact points p_i are the slot's
//   foot references (setReferencePose writes them into the dynamics' contact map, reference :151-169)
//   xdot = [h_lin/m ; sum_i f_i + m g ; sum_i (p_i - c) x f_i (+ tau_i)],   x+ = x + dt xdot
// Every Jacobian is closed form; the knot is written dense in KnotLayout order.
#pragma once
#include "dense_knot.cuh"

namespace smpc
{

  template <int NK_, int FS_>
  struct CentroCfg
  {
    static constexpr int NK = NK_, FS = FS_;
    static constexpr int NX = 9, NDX = 9, NU = NK * FS;
    static constexpr int ROWS = (FS == 6) ? 17 : 2;
    static constexpr int NCMAX = NK * ROWS;
  };

  template <class Cfg>
  struct CentroScratch : DenseScratch<9, 9, Cfg::NU, Cfg::NCMAX>
  {
    double lin[3], ang[3];
    double Jc[6][3];        // d[lin/m ; ang]/dc   (rows 0-2 are zero)
    double Ju[6][Cfg::NU];  // d[lin/m ; ang]/du
    double WJc[6][3], WJu[6][Cfg::NU], Wr[6];
    double racc[6];
  };

  template <class Cfg, bool DERIVS>
  SMPC_DEV void centro_stage_knot(const DevOcp & O, const KnotLayout & L, const KnotIO & io, CentroScratch<Cfg> & s)
  {
    constexpr int NU = Cfg::NU, NK = Cfg::NK, FS = Cfg::FS, NDX = 9, ROWS = Cfg::ROWS;
    const smpc_slot & slot = *io.slot;
    const unsigned cmask = slot.contact_mask;
    const double dt = O.dt, mass = O.mass;

    // ---- load (and, for a trial, step) the iterate: the state space is R^9 ----
    SMPC_PAR(tid)
    {
      for (int k = tid; k < 9; k += SMPC_NT)
      {
        if (io.dx == nullptr)
        {
          s.x[k] = io.x[k];
          s.xn_ws[k] = io.xn[k];
        }
        else
        {
          s.x[k] = io.x[k] + io.alpha * io.dx[k];
          s.xn_ws[k] = io.xn[k] + io.alpha * io.dxn[k];
          io.tx[k] = s.x[k];
        }
      }
      for (int k = tid; k < NU; k += SMPC_NT)
      {
        if (io.dx == nullptr)
          s.u[k] = io.u[k];
        else
        {
          s.u[k] = io.u[k] + io.alpha * io.du[k];
          io.tu[k] = s.u[k];
        }
      }
    }
    SMPC_END

    // ---- dynamics, acceleration residual Jacobians, constraint rows ----
    SMPC_PAR(tid)
    {
      if (tid == 0)
      {
        double lin[3] = {mass * O.gravity[0], mass * O.gravity[1], mass * O.gravity[2]};
        double ang[3] = {0, 0, 0};
        double dc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int r = 0; r < 6; r++)
          for (int j = 0; j < NU; j++)
            s.Ju[r][j] = 0.0;
        for (int f = 0; f < NK; f++)
          if (cmask & (1u << f))
          {
            const double * ff = s.u + FS * f;
            double lever[3], t[3];
            for (int r = 0; r < 3; r++)
              lever[r] = slot.foot_ref[f][9 + r] - s.x[r];
            v3cross(lever, ff, t);
            // [lever]x and [f]x
            const double Lx[9] = {0, -lever[2], lever[1], lever[2], 0, -lever[0], -lever[1], lever[0], 0};
            const double Fx[9] = {0, -ff[2], ff[1], ff[2], 0, -ff[0], -ff[1], ff[0], 0};
            for (int r = 0; r < 3; r++)
            {
              lin[r] += ff[r];
              ang[r] += t[r] + (FS == 6 ? ff[3 + r] : 0.0);
              s.Ju[r][FS * f + r] = 1.0 / mass;
              for (int c = 0; c < 3; c++)
              {
                s.Ju[3 + r][FS * f + c] = Lx[3 * r + c];
                dc[3 * r + c] += Fx[3 * r + c]; // d/dc (p - c) x f = [f]x
              }
              if (FS == 6)
                s.Ju[3 + r][FS * f + 3 + r] = 1.0;
            }
          }
        for (int r = 0; r < 3; r++)
        {
          s.lin[r] = lin[r];
          s.ang[r] = ang[r];
          s.xdot[r] = s.x[3 + r] / mass;
          s.xdot[3 + r] = lin[r];
          s.xdot[6 + r] = ang[r];
          s.racc[r] = lin[r] / mass;
          s.racc[3 + r] = ang[r];
          for (int c = 0; c < 3; c++)
          {
            s.Jc[r][c] = 0.0;
            s.Jc[3 + r][c] = dc[3 * r + c];
          }
        }
        for (int i = 0; i < 9; i++)
        {
          s.xnext[i] = s.x[i] + dt * s.xdot[i];
          s.e[i] = s.xnext[i] - s.xn_ws[i];
        }
      }
      else if (tid == 1)
      {
        // constraint rows, feet in contact in insertion order (reference src/centroidal-dynamics.cpp:85-104)
        int row = 0;
        for (int f = 0; f < NK; f++)
        {
          if (!(cmask & (1u << f)))
            continue;
          const double * ff = s.u + FS * f;
          if (FS == 6)
          {
            for (int rr = 0; rr < 17; rr++, row++)
            {
              double sacc = 0;
              for (int c = 0; c < 6; c++)
                sacc += O.wcone[6 * rr + c] * ff[c];
              s.cval[row] = sacc;
              s.ctype[row] = 1;
              s.lo[row] = s.hi[row] = 0;
              if (DERIVS)
              {
                for (int a = 0; a < NDX; a++)
                  s.Cx[row][a] = 0.0;
                for (int j = 0; j < NU; j++)
                  s.Cu[row][j] = 0.0;
                for (int c = 0; c < 6; c++)
                  s.Cu[row][FS * f + c] = O.wcone[6 * rr + c];
              }
            }
          }
          else
          {
            const double muf = O.mu, eps = 1e-4; // reference src/centroidal-dynamics.cpp:99
            s.cval[row] = -ff[2] + eps;
            s.cval[row + 1] = -muf * muf * ff[2] * ff[2] + ff[0] * ff[0] + ff[1] * ff[1];
            for (int k = 0; k < 2; k++)
            {
              s.ctype[row + k] = 1;
              s.lo[row + k] = s.hi[row + k] = 0;
              if (DERIVS)
              {
                for (int a = 0; a < NDX; a++)
                  s.Cx[row + k][a] = 0.0;
                for (int j = 0; j < NU; j++)
                  s.Cu[row + k][j] = 0.0;
              }
            }
            if (DERIVS)
            {
              s.Cu[row][FS * f + 2] = -1.0;
              s.Cu[row + 1][FS * f + 0] = 2 * ff[0];
              s.Cu[row + 1][FS * f + 1] = 2 * ff[1];
              s.Cu[row + 1][FS * f + 2] = -2 * muf * muf * ff[2];
            }
            row += 2;
          }
        }
        s.nc = row;
        (void)ROWS;
      }
    }
    SMPC_END

    // ---- cost value; weighted acceleration Jacobians ----
    SMPC_PAR(tid)
    {
      if (tid == 0)
      {
        // insertion order of the reference: com, control, linear mom, angular mom, linear acc, angular acc
        double cost = 0;
        const double * Ws[3] = {O.w_com, O.w_lmom, O.w_amom};
        for (int part = 0; part < 3; part++)
        {
          double r[3], c = 0;
          for (int i = 0; i < 3; i++)
            r[i] = s.x[3 * part + i] - slot.x_ref[3 * part + i];
          for (int i = 0; i < 3; i++)
          {
            double a = 0;
            for (int k = 0; k < 3; k++)
              a += Ws[part][3 * i + k] * r[k];
            s.lx[3 * part + i] = a; // J = selector: lx block = W r
            c += r[i] * a;
          }
          cost += 0.5 * c;
        }
        double cu = 0;
        for (int k = 0; k < NU; k++)
        {
          const double du = s.u[k] - slot.u_ref[k];
          cu += du * (O.w_u[k] * du);
        }
        cost += 0.5 * cu;
        for (int blk = 0; blk < 2; blk++)
        {
          const double * W = blk == 0 ? O.w_lacc : O.w_aacc;
          double c = 0;
          for (int i = 0; i < 3; i++)
          {
            double a = 0;
            for (int k = 0; k < 3; k++)
              a += W[3 * i + k] * s.racc[3 * blk + k];
            s.Wr[3 * blk + i] = a;
            c += s.racc[3 * blk + i] * a;
          }
          cost += 0.5 * c;
        }
        s.cost = cost;
      }
      if (DERIVS)
      {
        for (int e = tid; e < 6 * NU; e += SMPC_NT)
        {
          const int r = e / NU, j = e % NU, blk = r / 3;
          const double * W = blk == 0 ? O.w_lacc : O.w_aacc;
          double a = 0;
          for (int k = 0; k < 3; k++)
            a += W[3 * (r % 3) + k] * s.Ju[3 * blk + k][j];
          s.WJu[r][j] = a;
        }
        for (int e = tid; e < 6 * 3; e += SMPC_NT)
        {
          const int r = e / 3, c = e % 3, blk = r / 3;
          const double * W = blk == 0 ? O.w_lacc : O.w_aacc;
          double a = 0;
          for (int k = 0; k < 3; k++)
            a += W[3 * (r % 3) + k] * s.Jc[3 * blk + k][c];
          s.WJc[r][c] = a;
        }
      }
    }
    SMPC_END

    if (DERIVS)
    {
      SMPC_PAR(tid)
      {
        double * Am = io.knot + L.oA;
        for (int e = tid; e < 81; e += SMPC_NT)
        {
          const int r = e / 9, c = e % 9;
          double v = (r == c) ? 1.0 : 0.0;
          if (r < 3 && c == 3 + r)
            v += dt / mass;
          if (r >= 6 && c < 3)
            v += dt * s.Jc[3 + (r - 6)][c];
          Am[e] = v;
        }
        double * Bm = io.knot + L.oB;
        for (int e = tid; e < 9 * NU; e += SMPC_NT)
        {
          const int r = e / NU, j = e % NU;
          double v = 0;
          if (r >= 3 && r < 6)
            v = dt * (s.Ju[r - 3][j] * mass); // d lin / du
          else if (r >= 6)
            v = dt * s.Ju[3 + (r - 6)][j];
          Bm[e] = v;
        }
        double * Q = io.knot + L.oQ;
        for (int e = tid; e < 81; e += SMPC_NT)
        {
          const int a = e / 9, b = e % 9;
          double acc = 0;
          if (a / 3 == b / 3)
          {
            const double * W = (a < 3) ? O.w_com : (a < 6 ? O.w_lmom : O.w_amom);
            acc += W[3 * (a % 3) + (b % 3)];
          }
          if (a < 3 && b < 3)
          {
            double t = 0;
            for (int r = 0; r < 6; r++)
              t += s.Jc[r][a] * s.WJc[r][b];
            acc += t;
          }
          Q[e] = acc;
        }
        double * S = io.knot + L.oS;
        for (int e = tid; e < 9 * NU; e += SMPC_NT)
        {
          const int a = e / NU, j = e % NU;
          double acc = 0;
          if (a < 3)
            for (int r = 0; r < 6; r++)
              acc += s.Jc[r][a] * s.WJu[r][j];
          S[e] = acc;
        }
        double * Rm = io.knot + L.oR;
        for (int e = tid; e < NU * NU; e += SMPC_NT)
        {
          const int i = e / NU, j = e % NU;
          double acc = (i == j) ? O.w_u[i] : 0.0;
          double t = 0;
          for (int r = 0; r < 6; r++)
            t += s.Ju[r][i] * s.WJu[r][j];
          Rm[e] = acc + t;
        }
        // cost gradients (lx blocks of the three selector costs were set above)
        for (int a = tid; a < 3; a += SMPC_NT)
        {
          double t = 0;
          for (int r = 0; r < 6; r++)
            t += s.Jc[r][a] * s.Wr[r];
          s.lx[a] += t;
        }
        for (int j = tid; j < NU; j += SMPC_NT)
        {
          double g = O.w_u[j] * (s.u[j] - slot.u_ref[j]);
          double t = 0;
          for (int r = 0; r < 6; r++)
            t += s.Ju[r][j] * s.Wr[r];
          s.lu[j] = g + t;
        }
      }
      SMPC_END
    }
    dense_knot_tail<CentroScratch<Cfg>, DERIVS, false>(L, io, s);
  }

  // terminal node: linear and angular momentum costs only, no constraint
  // (reference src/centroidal-dynamics.cpp:306-328)
  template <class Cfg, bool DERIVS>
  SMPC_DEV void centro_terminal_knot(const DevOcp & O, const KnotLayout & L, const KnotIO & io, CentroScratch<Cfg> & s)
  {
    const smpc_terminal & tm = *io.term;
    SMPC_PAR(tid)
    {
      for (int k = tid; k < 9; k += SMPC_NT)
      {
        if (io.dx == nullptr)
          s.x[k] = io.x[k];
        else
        {
          s.x[k] = io.x[k] + io.alpha * io.dx[k];
          io.tx[k] = s.x[k];
        }
      }
    }
    SMPC_END
    SMPC_PAR(tid)
    {
      if (tid == 0)
      {
        double cost = 0;
        for (int i = 0; i < 3; i++)
          s.lx[i] = 0.0;
        for (int part = 1; part < 3; part++)
        {
          const double * W = part == 1 ? O.w_lmom : O.w_amom;
          double r[3], c = 0;
          for (int i = 0; i < 3; i++)
            r[i] = s.x[3 * part + i] - tm.x_ref[3 * part + i];
          for (int i = 0; i < 3; i++)
          {
            double a = 0;
            for (int k = 0; k < 3; k++)
              a += W[3 * i + k] * r[k];
            s.lx[3 * part + i] = a;
            c += r[i] * a;
          }
          cost += 0.5 * c;
        }
        s.cost = cost;
        s.nc = 0;
      }
      if (DERIVS)
      {
        double * Q = io.knot + L.oQ;
        for (int e = tid; e < 81; e += SMPC_NT)
        {
          const int a = e / 9, b = e % 9;
          double acc = 0;
          if (a >= 3 && a / 3 == b / 3)
          {
            const double * W = (a < 6) ? O.w_lmom : O.w_amom;
            acc = W[3 * (a % 3) + (b % 3)];
          }
          Q[e] = acc;
        }
      }
    }
    SMPC_END
    dense_knot_tail<CentroScratch<Cfg>, DERIVS, true>(L, io, s);
  }

} // namespace smpc
