// riccati.cuh — stage-wise proximal Riccati recursion, one MPC instance per CTA.
//
// Replaces aligator::gar::ProximalRiccatiSolver::backward(mudyn, mueq) / forward(...)
// ([UPSTREAM] aligator 0.16.0, selected at reference src/mpc.cpp:46-52) for the LQ problem that
// SolverProxDDP::updateLQSubproblem assembles.
//
// Per stage (backward, sequential in t):
//   1. P~ = (I + mudyn P+)^-1 P+ , p~ = (I + mudyn P+)^-1 p+        Cholesky + two triangular sweeps
//   2. w = P~ f + p~ ;  [PA PB] = P~ [A B]
//   3. Qh = Q + A^T PA ; Sh = S + A^T PB ; Rh = R + B^T PB ; qh = q + A^T w ; rh = r + B^T w
//   4. quasi-definite KKT [[Rh, D^T],[D, -mu I]] [u; nu] = -[[rh Sh^T],[d C]] by block LDL^T:
//        Rh = L L^T ; Y = L^-1 [rh Sh^T] ; W = L^-1 D^T ; Sg = mu I + W^T W = Ls Ls^T ;
//        G = Sg^-1 ([d C] - W^T Y) ;  [k K] = -L^-T (Y + W G) ;  [z Z] = G
//   5. P = sym(Qh + Sh K + C^T Z) ; p = qh + Sh k + C^T z
//
// Every output element is produced by ONE thread with fused multiply-adds in ascending (for the
// backward substitutions: descending) k — the canonical order of oracle/orc_solver.hpp — so the
// gains agree with the oracle to rounding of the inputs. The factorizations are right-looking
// (all threads update the trailing block after each pivot), which applies the same per-element
// update sequence as the oracle's left-looking loops.
//
// Shared memory (doubles): P|p [ndx(ndx+1)] · M→Qh [ndx²] · AB→{Y,W,Sg} · PAB→G · Sh Rh qh rh w
#pragma once
#include "par.cuh"
#include "dev_types.h"

namespace smpc
{

  struct RicSmem
  {
    int oP, oM, oAB, oPAB, oSh, oRh, oqh, orh, ow, oY, oW, oSg, oG, oflag, size;
  };
  inline RicSmem make_ric_smem(int ndx, int nu, int ncmax)
  {
    RicSmem R;
    int o = 0;
    R.oP = o, o += ndx * (ndx + 1);
    R.oM = o, o += ndx * ndx;
    const int szAB = ndx * (ndx + nu);
    const int szYW = nu * (ndx + 1) + nu * ncmax + ncmax * ncmax;
    R.oAB = o;
    R.oY = o;
    R.oW = o + nu * (ndx + 1);
    R.oSg = R.oW + nu * ncmax;
    o += (szAB > szYW ? szAB : szYW);
    const int szG = ncmax * (ndx + 1);
    R.oPAB = o;
    R.oG = o;
    o += (szAB > szG ? szAB : szG);
    R.oSh = o, o += ndx * nu;
    R.oRh = o, o += nu * nu;
    R.oqh = o, o += ndx;
    R.orh = o, o += nu;
    R.ow = o, o += ndx;
    R.oflag = o, o += 2;
    R.size = o;
    return R;
  }

  // In-place lower Cholesky of the n x n row-major matrix M (leading dimension ld). Right-looking.
  // `bad` (shared int) is set to 1 if a pivot is not positive.
  SMPC_DEV void chol_lower_par(double * M, int n, int ld, int * bad)
  {
    for (int j = 0; j < n; j++)
    {
      SMPC_PAR(tid)
      {
        const double s = M[j * ld + j];
        if (tid == 0 && !(s > 0))
          *bad = 1;
        const double l = sqrt(s);
        const double inv = 1.0 / l;
        for (int i = j + 1 + tid; i < n; i += SMPC_NT)
          M[i * ld + j] = M[i * ld + j] * inv;
      }
      SMPC_END
      SMPC_PAR(tid)
      {
        if (tid == 0)
          M[j * ld + j] = sqrt(M[j * ld + j]);
        // trailing update: M(i,c) -= L(i,j) L(c,j) for j < c <= i
        const int m = n - j - 1;
        for (int e = tid; e < m * m; e += SMPC_NT)
        {
          const int i = j + 1 + e / m, c = j + 1 + e % m;
          if (c <= i)
            M[i * ld + c] = fma(-M[i * ld + j], M[c * ld + j], M[i * ld + c]);
        }
      }
      SMPC_END
    }
  }
  // X <- L^-1 X : L lower n x n (ld ldl), X n x m (ld ldx). Right-looking over rows.
  SMPC_DEV void solve_lower_par(const double * L, int n, int ldl, double * X, int m, int ldx)
  {
    for (int k = 0; k < n; k++)
    {
      SMPC_PAR(tid)
      {
        const double ilkk = 1.0 / L[k * ldl + k];
        for (int c = tid; c < m; c += SMPC_NT)
          X[k * ldx + c] = X[k * ldx + c] * ilkk;
      }
      SMPC_END
      SMPC_PAR(tid)
      {
        const int rows = n - k - 1;
        for (int e = tid; e < rows * m; e += SMPC_NT)
        {
          const int i = k + 1 + e / m, c = e % m;
          X[i * ldx + c] = fma(-L[i * ldl + k], X[k * ldx + c], X[i * ldx + c]);
        }
      }
      SMPC_END
    }
  }
  // X <- L^-T X
  SMPC_DEV void solve_lower_T_par(const double * L, int n, int ldl, double * X, int m, int ldx)
  {
    for (int k = n - 1; k >= 0; k--)
    {
      SMPC_PAR(tid)
      {
        const double ilkk = 1.0 / L[k * ldl + k];
        for (int c = tid; c < m; c += SMPC_NT)
          X[k * ldx + c] = X[k * ldx + c] * ilkk;
      }
      SMPC_END
      SMPC_PAR(tid)
      {
        for (int e = tid; e < k * m; e += SMPC_NT)
        {
          const int i = e / m, c = e % m;
          X[i * ldx + c] = fma(-L[k * ldl + i], X[k * ldx + c], X[i * ldx + c]);
        }
      }
      SMPC_END
    }
  }

  // Backward sweep for one instance. knots: (N+1) dense knots; ncs: (N+1) row counts; sol: N stage solutions
  // + terminal (Z,z only). P0/p0 (optional) receive the value function of node 0.
  SMPC_DEV void riccati_backward_inst(
    const KnotLayout & L, const SolLayout & SL, const RicSmem & RS, int N, const double * __restrict__ knots,
    const int * __restrict__ ncs, double * __restrict__ sol, double mudyn, double mueq, double * sm, int * flag_out,
    double * P0out)
  {
    const int ndx = L.ndx, nu = L.nu;
    const int ldp = ndx + 1;
    double * P = sm + RS.oP; // [ndx][ndx+1]: cols 0..ndx-1 = P, col ndx = p
    double * Mb = sm + RS.oM;
    double * AB = sm + RS.oAB;
    double * PAB = sm + RS.oPAB;
    double * Sh = sm + RS.oSh;
    double * Rh = sm + RS.oRh;
    double * qh = sm + RS.oqh;
    double * rh = sm + RS.orh;
    double * w = sm + RS.ow;
    double * Y = sm + RS.oY;
    double * Wm = sm + RS.oW;
    double * Sg = sm + RS.oSg;
    double * G = sm + RS.oG;
    int * bad = (int *)(sm + RS.oflag);
    const double inv_mu = 1.0 / mueq;
    const int nab = ndx + nu;

    // ---- terminal node ----
    {
      const double * kn = knots + (size_t)N * L.size;
      const int nc = ncs[N];
      double * so = sol + (size_t)N * SL.size;
      SMPC_PAR(tid)
      {
        if (tid == 0)
          *bad = 0;
        for (int e = tid; e < nc * ndx; e += SMPC_NT)
          so[SL.oZ + e] = kn[L.oC + e] * inv_mu;
        for (int i = tid; i < nc; i += SMPC_NT)
          so[SL.oz + i] = kn[L.od + i] * inv_mu;
      }
      SMPC_END
      SMPC_PAR(tid)
      for (int e = tid; e < ndx * ldp; e += SMPC_NT)
      {
        const int a = e / ldp, b = e % ldp;
        if (b < ndx)
        {
          double acc = kn[L.oQ + a * ndx + b];
          for (int i = 0; i < nc; i++)
            acc = fma(kn[L.oC + i * ndx + a], so[SL.oZ + i * ndx + b], acc);
          P[e] = acc;
        }
        else
        {
          double acc = kn[L.oq + a];
          for (int i = 0; i < nc; i++)
            acc = fma(kn[L.oC + i * ndx + a], so[SL.oz + i], acc);
          P[e] = acc;
        }
      }
      SMPC_END
    }

    for (int t = N - 1; t >= 0; t--)
    {
      const double * kn = knots + (size_t)t * L.size;
      const int nc = ncs[t];
      double * so = sol + (size_t)t * SL.size;
      // 1. P~ | p~ in place
      if (mudyn > 0)
      {
        SMPC_PAR(tid)
        for (int e = tid; e < ndx * ndx; e += SMPC_NT)
        {
          const int a = e / ndx, b = e % ndx;
          Mb[e] = mudyn * P[a * ldp + b] + (a == b ? 1.0 : 0.0);
        }
        SMPC_END
        chol_lower_par(Mb, ndx, ndx, bad);
        solve_lower_par(Mb, ndx, ndx, P, ldp, ldp);
        solve_lower_T_par(Mb, ndx, ndx, P, ldp, ldp);
      }
      // stage data to shared memory; P~ p~ to HBM for the forward sweep
      SMPC_PAR(tid)
      {
        for (int e = tid; e < ndx * nab; e += SMPC_NT)
        {
          const int a = e / nab, b = e % nab;
          AB[e] = (b < ndx) ? kn[L.oA + a * ndx + b] : kn[L.oB + a * nu + (b - ndx)];
        }
        for (int e = tid; e < ndx * ldp; e += SMPC_NT)
        {
          const int a = e / ldp, b = e % ldp;
          if (b < ndx)
            so[SL.oPt + a * ndx + b] = P[e];
          else
            so[SL.opt + a] = P[e];
        }
      }
      SMPC_END
      // 2. w = P~ f + p~ ; PAB = P~ [A B]
      SMPC_PAR(tid)
      {
        for (int a = tid; a < ndx; a += SMPC_NT)
        {
          double acc = P[a * ldp + ndx];
          for (int b = 0; b < ndx; b++)
            acc = fma(P[a * ldp + b], kn[L.of + b], acc);
          w[a] = acc;
        }
        for (int e = tid; e < ndx * nab; e += SMPC_NT)
        {
          const int a = e / nab, b = e % nab;
          double acc = 0;
          for (int k = 0; k < ndx; k++)
            acc = fma(P[a * ldp + k], AB[k * nab + b], acc);
          PAB[e] = acc;
        }
      }
      SMPC_END
      // 3. Qh (in Mb), Sh, Rh, qh, rh
      SMPC_PAR(tid)
      {
        for (int e = tid; e < ndx * ndx; e += SMPC_NT)
        {
          const int a = e / ndx, b = e % ndx;
          double acc = kn[L.oQ + e];
          for (int k = 0; k < ndx; k++)
            acc = fma(AB[k * nab + a], PAB[k * nab + b], acc);
          Mb[e] = acc;
        }
        for (int e = tid; e < ndx * nu; e += SMPC_NT)
        {
          const int a = e / nu, b = e % nu;
          double acc = kn[L.oS + e];
          for (int k = 0; k < ndx; k++)
            acc = fma(AB[k * nab + a], PAB[k * nab + ndx + b], acc);
          Sh[e] = acc;
        }
        for (int e = tid; e < nu * nu; e += SMPC_NT)
        {
          const int a = e / nu, b = e % nu;
          double acc = kn[L.oR + e];
          for (int k = 0; k < ndx; k++)
            acc = fma(AB[k * nab + ndx + a], PAB[k * nab + ndx + b], acc);
          Rh[e] = acc;
        }
        for (int a = tid; a < ndx; a += SMPC_NT)
        {
          double acc = kn[L.oq + a];
          for (int k = 0; k < ndx; k++)
            acc = fma(AB[k * nab + a], w[k], acc);
          qh[a] = acc;
        }
        for (int a = tid; a < nu; a += SMPC_NT)
        {
          double acc = kn[L.or_ + a];
          for (int k = 0; k < ndx; k++)
            acc = fma(AB[k * nab + ndx + a], w[k], acc);
          rh[a] = acc;
        }
      }
      SMPC_END
      // 4. KKT
      chol_lower_par(Rh, nu, nu, bad);
      SMPC_PAR(tid)
      {
        for (int e = tid; e < nu * ldp; e += SMPC_NT)
        {
          const int a = e / ldp, c = e % ldp;
          Y[e] = (c == 0) ? rh[a] : Sh[(c - 1) * nu + a];
        }
        for (int e = tid; e < nu * nc; e += SMPC_NT)
        {
          const int a = e / nc, i = e % nc;
          Wm[e] = kn[L.oD + i * nu + a];
        }
      }
      SMPC_END
      solve_lower_par(Rh, nu, nu, Y, ldp, ldp);
      if (nc > 0)
      {
        solve_lower_par(Rh, nu, nu, Wm, nc, nc);
        SMPC_PAR(tid)
        {
          for (int e = tid; e < nc * nc; e += SMPC_NT)
          {
            const int i = e / nc, j = e % nc;
            // symmetric: evaluate the (max,min) ordered pair so (i,j) and (j,i) are bit-identical
            const int hi = i > j ? i : j, lo = i > j ? j : i;
            double acc = (i == j) ? mueq : 0.0;
            for (int k = 0; k < nu; k++)
              acc = fma(Wm[k * nc + hi], Wm[k * nc + lo], acc);
            Sg[e] = acc;
          }
          for (int e = tid; e < nc * ldp; e += SMPC_NT)
          {
            const int i = e / ldp, c = e % ldp;
            double acc = (c == 0) ? kn[L.od + i] : kn[L.oC + i * ndx + (c - 1)];
            for (int k = 0; k < nu; k++)
              acc = fma(-Wm[k * nc + i], Y[k * ldp + c], acc);
            G[e] = acc;
          }
        }
        SMPC_END
        chol_lower_par(Sg, nc, nc, bad);
        solve_lower_par(Sg, nc, nc, G, ldp, ldp);
        solve_lower_T_par(Sg, nc, nc, G, ldp, ldp);
        SMPC_PAR(tid)
        for (int e = tid; e < nu * ldp; e += SMPC_NT)
        {
          const int a = e / ldp, c = e % ldp;
          double acc = Y[e];
          for (int i = 0; i < nc; i++)
            acc = fma(Wm[a * nc + i], G[i * ldp + c], acc);
          Y[e] = acc;
        }
        SMPC_END
      }
      solve_lower_T_par(Rh, nu, nu, Y, ldp, ldp);
      // gains out: K = -Y[:,1:], k = -Y[:,0], Z = G[:,1:], z = G[:,0]  (Y negated in place for step 5)
      SMPC_PAR(tid)
      {
        for (int e = tid; e < nu * ldp; e += SMPC_NT)
        {
          const int a = e / ldp, c = e % ldp;
          const double val = -Y[e];
          Y[e] = val;
          if (c == 0)
            so[SL.ok + a] = val;
          else
            so[SL.oK + a * ndx + (c - 1)] = val;
        }
        for (int e = tid; e < nc * ldp; e += SMPC_NT)
        {
          const int i = e / ldp, c = e % ldp;
          if (c == 0)
            so[SL.oz + i] = G[e];
          else
            so[SL.oZ + i * ndx + (c - 1)] = G[e];
        }
      }
      SMPC_END
      // 5. P = Qh + Sh K + C^T Z ; p = qh + Sh k + C^T z
      SMPC_PAR(tid)
      for (int e = tid; e < ndx * ldp; e += SMPC_NT)
      {
        const int a = e / ldp, b = e % ldp;
        if (b < ndx)
        {
          double acc = Mb[a * ndx + b];
          for (int k = 0; k < nu; k++)
            acc = fma(Sh[a * nu + k], Y[k * ldp + 1 + b], acc);
          for (int i = 0; i < nc; i++)
            acc = fma(kn[L.oC + i * ndx + a], G[i * ldp + 1 + b], acc);
          P[e] = acc;
        }
        else
        {
          double acc = qh[a];
          for (int k = 0; k < nu; k++)
            acc = fma(Sh[a * nu + k], Y[k * ldp], acc);
          for (int i = 0; i < nc; i++)
            acc = fma(kn[L.oC + i * ndx + a], G[i * ldp], acc);
          P[e] = acc;
        }
      }
      SMPC_END
      SMPC_PAR(tid)
      for (int e = tid; e < ndx * ndx; e += SMPC_NT)
      {
        const int a = e / ndx, b = e % ndx;
        if (b < a)
        {
          const double m = 0.5 * (P[a * ldp + b] + P[b * ldp + a]);
          Mb[e] = m; // stash, write back after the barrier
        }
      }
      SMPC_END
      SMPC_PAR(tid)
      for (int e = tid; e < ndx * ndx; e += SMPC_NT)
      {
        const int a = e / ndx, b = e % ndx;
        if (b < a)
        {
          P[a * ldp + b] = Mb[e];
          P[b * ldp + a] = Mb[e];
        }
      }
      SMPC_END
    }
    SMPC_PAR(tid)
    {
      if (tid == 0 && flag_out)
        *flag_out = *bad;
      if (P0out)
        for (int e = tid; e < ndx * ldp; e += SMPC_NT)
          P0out[e] = P[e];
    }
    SMPC_END
  }

  // Forward sweep for one instance (dx0 = 0: force_initial_condition, reference src/mpc.cpp:53).
  // Outputs: dxs [N+1][ndx], dus [N][nu], vnew [N+1][ncmax] (total new multipliers), lnew [N+1][ndx],
  // and the merit directional derivative  sum_t gx_t.dx_t + gu_t.du_t - lplus_t.dx_t  (t >= 1 for lplus).
  // sm needs 3 ndx + nu + 2 doubles.
  SMPC_DEV void riccati_forward_inst(
    const KnotLayout & L, const SolLayout & SL, int N, const double * __restrict__ knots, const int * __restrict__ ncs,
    const double * __restrict__ sol, const double * __restrict__ P0, double mudyn, const double * __restrict__ gx,
    const double * __restrict__ gu, const double * __restrict__ lplus, double * __restrict__ dxs, double * __restrict__ dus,
    double * __restrict__ vnew, double * __restrict__ lnew, double * dphi_out, double * sm)
  {
    const int ndx = L.ndx, nu = L.nu, ncmax = L.ncmax;
    double * dx = sm;
    double * du = sm + ndx;
    double * y = du + nu;
    double * lam = y + ndx;
    double * part = lam + ndx; // partial dphi accumulators, one per thread slot < 64
    SMPC_PAR(tid)
    {
      for (int a = tid; a < ndx; a += SMPC_NT)
      {
        dx[a] = 0.0;
        dxs[a] = 0.0;
        lnew[a] = P0[a * (ndx + 1) + ndx]; // lam_0 = P_0 dx_0 + p_0 with dx_0 = 0
      }
      for (int a = tid; a < 64; a += SMPC_NT)
        part[a] = 0.0;
    }
    SMPC_END
    for (int t = 0; t < N; t++)
    {
      const double * kn = knots + (size_t)t * L.size;
      const double * so = sol + (size_t)t * SL.size;
      const int nc = ncs[t];
      SMPC_PAR(tid)
      {
        for (int a = tid; a < nu; a += SMPC_NT)
        {
          double acc = so[SL.ok + a];
          for (int b = 0; b < ndx; b++)
            acc = fma(so[SL.oK + a * ndx + b], dx[b], acc);
          du[a] = acc;
          dus[(size_t)t * nu + a] = acc;
        }
        for (int i = tid; i < ncmax; i += SMPC_NT)
        {
          double acc = 0;
          if (i < nc)
          {
            acc = so[SL.oz + i];
            for (int b = 0; b < ndx; b++)
              acc = fma(so[SL.oZ + i * ndx + b], dx[b], acc);
          }
          vnew[(size_t)t * ncmax + i] = acc;
        }
      }
      SMPC_END
      SMPC_PAR(tid)
      {
        for (int a = tid; a < ndx; a += SMPC_NT)
        {
          double acc = kn[L.of + a];
          for (int b = 0; b < ndx; b++)
            acc = fma(kn[L.oA + a * ndx + b], dx[b], acc);
          for (int b = 0; b < nu; b++)
            acc = fma(kn[L.oB + a * nu + b], du[b], acc);
          y[a] = acc;
          // directional derivative pieces of this node
          double dp = gx[(size_t)t * ndx + a] * dx[a];
          if (t > 0)
            dp -= lplus[(size_t)t * ndx + a] * dx[a];
          part[a & 63] += dp;
        }
      }
      SMPC_END
      SMPC_PAR(tid)
      {
        for (int a = tid; a < ndx; a += SMPC_NT)
        {
          double acc = so[SL.opt + a];
          for (int b = 0; b < ndx; b++)
            acc = fma(so[SL.oPt + a * ndx + b], y[b], acc);
          lam[a] = acc;
          lnew[(size_t)(t + 1) * ndx + a] = acc;
        }
        for (int a = tid; a < nu; a += SMPC_NT)
          part[a & 63] += gu[(size_t)t * nu + a] * du[a];
      }
      SMPC_END
      SMPC_PAR(tid)
      for (int a = tid; a < ndx; a += SMPC_NT)
      {
        const double v = y[a] - mudyn * lam[a];
        dx[a] = v;
        dxs[(size_t)(t + 1) * ndx + a] = v;
      }
      SMPC_END
    }
    // terminal multipliers and the last node's gradient term
    {
      const double * so = sol + (size_t)N * SL.size;
      const int nc = ncs[N];
      SMPC_PAR(tid)
      {
        for (int i = tid; i < ncmax; i += SMPC_NT)
        {
          double acc = 0;
          if (i < nc)
          {
            acc = so[SL.oz + i];
            for (int b = 0; b < ndx; b++)
              acc = fma(so[SL.oZ + i * ndx + b], dx[b], acc);
          }
          vnew[(size_t)N * ncmax + i] = acc;
        }
        for (int a = tid; a < ndx; a += SMPC_NT)
          part[a & 63] += (gx[(size_t)N * ndx + a] - lplus[(size_t)N * ndx + a]) * dx[a];
      }
      SMPC_END
    }
    SMPC_PAR(tid)
    if (tid == 0)
    {
      double sacc = 0;
      for (int a = 0; a < 64; a++)
        sacc += part[a];
      *dphi_out = sacc;
    }
    SMPC_END
  }

} // namespace smpc
