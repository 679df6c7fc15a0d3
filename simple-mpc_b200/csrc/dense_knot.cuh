// dense_knot.cuh — the model-independent half of one horizon knot for problems whose stage data is
// small enough to be held DENSE in shared memory (centroidal: ndx 9, nu 12; full dynamics later).
//
// The model code (centro_knot.cuh, ...) fills a DenseScratch with
//     x+ = f(x,u) and the dynamics slack e, the stage cost and its gradient (lx, lu),
//     constraint values / set types / bounds and the FULL constraint Jacobians (Cx, Cu),
// and writes A, B and the Gauss-Newton Q, S, R blocks (without the proximal term) straight into the
// dense LQ knot. dense_knot_tail() then does what aligator's SolverProxDDP does per stage
// [UPSTREAM: computeMultipliers / updateLQSubproblem / computeInfeasibilities / computeCriterion]:
//     first-order multipliers lam+ = lam_e + e/mu, nu+ = Pi_N(nu_e + c/mu); active sets;
//     q, r, f, active-set C, D, d; proximal terms on Q, R; Lagrangian gradients with the current and
//     the first-order multipliers; merit / infeasibility / inner-criterion pieces.
// The operation order follows oracle/orc_solver.hpp (multipliers, build_lq, lagrangian_grad).
#pragma once
#include "dev_math.cuh"
#include "dev_types.h"
#include "kino_knot.cuh" // KnotIO, ncone_proj

namespace smpc
{

  template <int NX_, int NDX_, int NU_, int NCMAX_>
  struct DenseScratch
  {
    static constexpr int NX = NX_, NDX = NDX_, NU = NU_, NCMAX = NCMAX_;
    double x[NX], u[NU > 0 ? NU : 1], xn_ws[NX], xnext[NX];
    double e[NDX], lplus[NDX], xdot[NDX];
    double lx[NDX], lu[NU > 0 ? NU : 1];
    double Cx[NCMAX][NDX], Cu[NCMAX][NU > 0 ? NU : 1];
    double cval[NCMAX], lo[NCMAX], hi[NCMAX];
    int ctype[NCMAX]; // 0 equality, 1 negative orthant, 2 box
    double dproj[NCMAX], vplus[NCMAX], rpinf[NCMAX], rcrit[NCMAX];
    int ractive[NCMAX];
    int nc;
    double cost;
    double Lx[NDX], Lu[NU > 0 ? NU : 1], gx[NDX], gu[NU > 0 ? NU : 1];
  };

  // TERMINAL: no control, no dynamics (only Q, q, C, d of the knot are meaningful)
  template <class S, bool DERIVS, bool TERMINAL>
  SMPC_DEV void dense_knot_tail(const KnotLayout & L, const KnotIO & io, S & s)
  {
    constexpr int NDX = S::NDX, NU = S::NU;
    const double mu = io.mu, inv_mu = 1.0 / io.mu;
    const int nc = s.nc;
    // ---- first-order multipliers ----
    SMPC_PAR(tid)
    {
      if (!TERMINAL)
        for (int k = tid; k < NDX; k += SMPC_NT)
          s.lplus[k] = io.prev_lam_n[k] + s.e[k] * inv_mu;
      for (int row = tid; row < nc; row += SMPC_NT)
      {
        const double z = s.cval[row] + mu * io.prev_vs[row];
        int act;
        const double pr = ncone_proj(s.ctype[row], z, s.lo[row], s.hi[row], act);
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
      SMPC_PAR(tid)
      {
        const double * Am = io.knot + L.oA;
        const double * Bm = io.knot + L.oB;
        for (int a = tid; a < NDX; a += SMPC_NT)
        {
          io.knot[L.oQ + a * NDX + a] += io.preg;
          io.knot[L.oq + a] = s.lx[a];
          double La = s.lx[a], ga = s.lx[a];
          if (!TERMINAL)
          {
            io.knot[L.of + a] = s.e[a] + mu * io.prev_lam_n[a];
            double t1 = 0, t2 = 0;
            for (int b = 0; b < NDX; b++)
            {
              t1 += Am[b * NDX + a] * io.lam_n[b];
              t2 += Am[b * NDX + a] * s.lplus[b];
            }
            La += t1;
            ga += t2;
          }
          double Lc = 0, gc = 0;
          for (int row = 0; row < nc; row++)
          {
            Lc += s.Cx[row][a] * io.vs[row];
            gc += s.Cx[row][a] * s.vplus[row];
          }
          s.Lx[a] = La + Lc - io.lam_t[a];
          s.gx[a] = ga + gc;
        }
        if (!TERMINAL)
          for (int j = tid; j < NU; j += SMPC_NT)
          {
            io.knot[L.oR + j * NU + j] += io.preg;
            io.knot[L.or_ + j] = s.lu[j];
            double t1 = 0, t2 = 0;
            for (int b = 0; b < NDX; b++)
            {
              t1 += Bm[b * NU + j] * io.lam_n[b];
              t2 += Bm[b * NU + j] * s.lplus[b];
            }
            double Lc = 0, gc = 0;
            for (int row = 0; row < nc; row++)
            {
              Lc += s.Cu[row][j] * io.vs[row];
              gc += s.Cu[row][j] * s.vplus[row];
            }
            s.Lu[j] = s.lu[j] + t1 + Lc;
            s.gu[j] = s.lu[j] + t2 + gc;
          }
      }
      SMPC_END
      // (phase boundary: the full-dynamics knot keeps its Cx / Cu rows in the knot's own C / D blocks, so the
      //  masking below must not start before every thread has finished reading them above)
      SMPC_PAR(tid)
      {
        double * Cm = io.knot + L.oC;
        for (int e = tid; e < L.ncmax * NDX; e += SMPC_NT)
        {
          const int row = e / NDX, a = e % NDX;
          Cm[e] = (row < nc && s.ractive[row]) ? s.Cx[row][a] : 0.0;
        }
        if (!TERMINAL)
        {
          double * Dm = io.knot + L.oD;
          for (int e = tid; e < L.ncmax * NU; e += SMPC_NT)
          {
            const int row = e / NU, j = e % NU;
            Dm[e] = (row < nc && s.ractive[row]) ? s.Cu[row][j] : 0.0;
          }
        }
        for (int row = tid; row < L.ncmax; row += SMPC_NT)
          io.knot[L.od + row] = (row < nc) ? s.dproj[row] : 0.0;
      }
      SMPC_END
    }

    SMPC_PAR(tid)
    {
      if (tid == 0)
      {
        double pen = 0, pinf = 0, crit = 0;
        if (!TERMINAL)
        {
          double sq = 0;
          for (int k = 0; k < NDX; k++)
          {
            sq += s.lplus[k] * s.lplus[k];
            pinf = fmax(pinf, fabs(s.e[k]));
            crit = fmax(crit, fabs(mu * (s.lplus[k] - io.lam_n[k])));
          }
          pen = 0.5 * mu * sq;
        }
        double sqv = 0;
        for (int row = 0; row < nc; row++)
        {
          sqv += s.vplus[row] * s.vplus[row];
          pinf = fmax(pinf, s.rpinf[row]);
          crit = fmax(crit, s.rcrit[row]);
        }
        pen += 0.5 * mu * sqv;
        double dinf = 0, dinfu = 0;
        if (DERIVS)
        {
          for (int k = 0; k < NDX; k++)
            dinf = fmax(dinf, fabs(s.Lx[k]));
          if (!TERMINAL)
            for (int k = 0; k < NU; k++)
              dinfu = fmax(dinfu, fabs(s.Lu[k]));
          io.sc->pad[0] = dinf;
          io.sc->pad[1] = dinfu;
          dinf = fmax(dinf, dinfu);
        }
        io.sc->cost = s.cost;
        io.sc->pen = pen;
        io.sc->pinf = pinf;
        io.sc->dinf = dinf;
        io.sc->crit = crit;
        *io.nc_out = nc;
      }
      if (!TERMINAL)
        for (int k = tid; k < NDX; k += SMPC_NT)
        {
          io.xdot[k] = s.xdot[k];
          io.lplus_n[k] = s.lplus[k];
        }
      if (DERIVS)
      {
        for (int k = tid; k < NDX; k += SMPC_NT)
          io.gx[k] = s.gx[k];
        if (!TERMINAL)
          for (int k = tid; k < NU; k += SMPC_NT)
            io.gu[k] = s.gu[k];
      }
    }
    SMPC_END
  }

} // namespace smpc
