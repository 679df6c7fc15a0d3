// riccati_par.cuh — the PARALLEL (leg-condensed) proximal Riccati sweep: one CTA per LEG of the horizon.
//
// Replaces aligator::gar::ParallelRiccatiSolver ([UPSTREAM] aligator 0.16.0), which the reference selects with
// `num_threads > 1` (reference src/mpc.cpp:46-50, `linear_solver_choice = LQSolverChoice::PARALLEL`).
// It is the answer to SINGLE-INSTANCE latency: the serial sweep is N strictly sequential stages on one SM while 147 SMs idle.
//
// The horizon [0, N) is cut into J legs [a_j, a_{j+1}). With the co-state theta at the end of a leg as a parameter
// (cost-to-go P = 0, p = theta there) the leg's end state is affine in its start state and theta,
//     x_b = Psi x_a + psi - Lam theta ,      lambda_a = P_a x_a + p_a + Psi^T theta ,
// so the legs couple through a small block-tridiagonal system in (x_{a_j}, theta_j). aligator carries the theta columns
// through every stage factorisation; here the TUNED serial sweeps (riccati3.cuh / riccati4.cuh) stay untouched and run
// twice per leg, with two small kernels in between (canonical order: oracle/orc_solver.hpp riccati_backward_parallel):
//   pass 1  legs 1..J-1 in parallel: the ordinary sweep from a ZERO cost-to-go at the leg end (the true terminal node for
//           the last leg) -> gains K, k and P~, p~ per stage in `sol`, (P_a, p_a) per leg;
//   pass 2  legs 1..J-2 in parallel (leg_element_inst): (Psi, psi, Lam) from the stored gains, stage by stage backwards,
//             Pi~ = (I - mu P~) Pi+ ;  Y = B^T Pi~ ;  X = H Y  (H = (u,u) block of the stage's KKT inverse, rebuilt from
//             R^ = R + B^T P~ B and D) ;  Lam += Y^T X + mu Pi+^T Pi~ ;  psi += Pi+^T [(I - mu P~)(B k + f) - mu p~] ;
//             Pi = A^T Pi~ + K^T Y                      (Pi = Psi^T ; Pi = I, psi = 0, Lam = 0 at the leg end)
//   pass 3  one CTA per instance, serial over the J-2 inner legs (leg_coarse_inst): the TRUE cost-to-go at every leg start,
//             calP_j = P_j + Psi_j^T calP_{j+1} (I + Lam_j calP_{j+1})^-1 Psi_j     (+ the vector part),
//           evaluated as C (I + C^T Lam C)^-1 C^T with calP_{j+1} = C C^T: symmetric positive definite throughout;
//   pass 4  legs 0..J-2 in parallel: the ordinary sweep again, now from the true cost-to-go of the next leg.
// After pass 4 `sol` and P0 hold what the serial sweep computes, up to rounding in the hand-over (gains within 1e-9 of the
// serial sweep for 3..10 legs, tests/test_parallel_riccati.py), so the forward sweep and everything downstream are unchanged.
// Critical path: 2 x N/J tuned stages + N/J element stages + (J-2) condensation steps instead of N tuned stages.
#pragma once
#include "riccati.cuh"
#include "riccati_par_layout.h"

namespace smpc
{
  // ---- small dense kernels of the leg-element / condensation kernels. These kernels sit on the critical path of a
  // latency-bound tick, so their factorisations and triangular solves avoid CTA-wide barriers (the right-looking routines
  // of riccati.cuh take two per pivot: 350 barriers per stage made pass 2 slower than the sweep it is meant to shorten) ----
  // Cholesky of an n x n matrix (lower triangle, in place) by ONE warp, lane i owning rows i, i + 32, ...
  SMPC_DEV void chol_lower_warp(double * M, int n, int ld, int * bad)
  {
#ifdef __CUDACC__
    if (threadIdx.x < 32)
    {
      const int lane = threadIdx.x;
      for (int k = 0; k < n; k++)
      {
        if (lane == (k & 31))
        {
          double dk = M[k * ld + k];
          if (!(dk > 0.0))
          {
            *bad = 1;
            dk = 1.0;
          }
          M[k * ld + k] = sqrt(dk);
        }
        __syncwarp();
        const double inv = 1.0 / M[k * ld + k];
        for (int i = lane; i < n; i += 32)
          if (i > k)
            M[i * ld + k] *= inv;
        __syncwarp();
        for (int i = lane; i < n; i += 32)
          if (i > k)
          {
            const double lik = M[i * ld + k];
            for (int j = k + 1; j <= i; j++)
              M[i * ld + j] = fma(-lik, M[j * ld + k], M[i * ld + j]);
          }
        __syncwarp();
      }
    }
    __syncthreads();
#else
    for (int k = 0; k < n; k++)
    {
      double dk = M[k * ld + k];
      if (!(dk > 0.0))
      {
        *bad = 1;
        dk = 1.0;
      }
      M[k * ld + k] = sqrt(dk);
      const double inv = 1.0 / M[k * ld + k];
      for (int i = k + 1; i < n; i++)
        M[i * ld + k] *= inv;
      for (int i = k + 1; i < n; i++)
      {
        const double lik = M[i * ld + k];
        for (int j = k + 1; j <= i; j++)
          M[i * ld + j] = fma(-lik, M[j * ld + k], M[i * ld + j]);
      }
    }
#endif
  }
  // X <- L^-1 X and X <- L^-T X, one thread per column of X (its column is touched by no other thread: no barrier inside)
  SMPC_DEV void solve_lower_cols(const double * L, int n, int ldl, double * X, int m, int ldx)
  {
    SMPC_PAR(tid)
    for (int c = tid; c < m; c += SMPC_NT)
      for (int i = 0; i < n; i++)
      {
        double acc = X[i * ldx + c];
        for (int k = 0; k < i; k++)
          acc = fma(-L[i * ldl + k], X[k * ldx + c], acc);
        X[i * ldx + c] = acc / L[i * ldl + i];
      }
    SMPC_END
  }
  SMPC_DEV void solve_lower_T_cols(const double * L, int n, int ldl, double * X, int m, int ldx)
  {
    SMPC_PAR(tid)
    for (int c = tid; c < m; c += SMPC_NT)
      for (int i = n - 1; i >= 0; i--)
      {
        double acc = X[i * ldx + c];
        for (int k = i + 1; k < n; k++)
          acc = fma(-L[k * ldl + i], X[k * ldx + c], acc);
        X[i * ldx + c] = acc / L[i * ldl + i];
      }
    SMPC_END
  }

  // pass 2 for the leg [a, b) of one instance. knots / ncs / sol are the INSTANCE's arrays; out = leg_elem_size doubles.
  SMPC_DEV void leg_element_inst(
    const KnotLayout & L, const SolLayout & SL, const ParSmem & RS, int a, int b, const double * __restrict__ knots,
    const int * __restrict__ ncs, const double * __restrict__ sol, double mudyn, double mueq, double * sm, int * flag_out,
    double * __restrict__ out)
  {
    const int n = L.ndx, nu = L.nu;
    double * Pi = sm + RS.oPi;
    double * Pit = sm + RS.oPit;
    double * Lam = sm + RS.oLam;
    double * Y = sm + RS.oY;
    double * X = sm + RS.oX;
    double * PB = sm + RS.oX;
    double * Rh = sm + RS.oRh;
    double * W = sm + RS.oW;
    double * Sg = sm + RS.oSg;
    double * G = sm + RS.oG;
    double * v = sm + RS.ov;
    double * phi = sm + RS.ophi;
    double * psi = sm + RS.opsi;
    int * bad = (int *)(sm + RS.oflag);
    SMPC_PAR(tid)
    {
      if (tid == 0)
        bad[0] = bad[1] = 0;
      for (int e = tid; e < n * n; e += SMPC_NT)
      {
        Pi[e] = (e / n == e % n) ? 1.0 : 0.0;
        Lam[e] = 0.0;
      }
      for (int i = tid; i < n; i += SMPC_NT)
        psi[i] = 0.0;
    }
    SMPC_END
    for (int t = b - 1; t >= a; t--)
    {
      const double * kn = knots + (size_t)t * L.size;
      const double * so = sol + (size_t)t * SL.size;
      const double * Pt = so + SL.oPt;
      const double * Kk = so + SL.oK;
      const double * Ak = kn + L.oA;
      const double * Bk = kn + L.oB;
      const int nc = ncs[t];
      // Pi~ = Pi+ - mu P~ Pi+ ; v = B k + f
      SMPC_PAR(tid)
      {
        for (int e = tid; e < n * n; e += SMPC_NT)
        {
          const int i = e / n, j = e % n;
          double acc = 0.0;
          for (int k = 0; k < n; k++)
            acc = fma(Pt[i * n + k], Pi[k * n + j], acc);
          Pit[e] = fma(-mudyn, acc, Pi[e]);
        }
        for (int i = tid; i < n; i += SMPC_NT)
        {
          double acc = kn[L.of + i];
          for (int k = 0; k < nu; k++)
            acc = fma(Bk[i * nu + k], so[SL.ok + k], acc);
          v[i] = acc;
        }
        // P~ B (the products of this phase read Pi / global memory only)
        for (int e = tid; e < n * nu; e += SMPC_NT)
        {
          const int i = e / nu, j = e % nu;
          double acc = 0.0;
          for (int k = 0; k < n; k++)
            acc = fma(Pt[i * n + k], Bk[k * nu + j], acc);
          PB[e] = acc;
        }
      }
      SMPC_END
      // Y = B^T Pi~ ; R^ = R + B^T (P~ B) (lower triangle) ; phi = v - mu (P~ v + p~)
      SMPC_PAR(tid)
      {
        for (int e = tid; e < nu * n; e += SMPC_NT)
        {
          const int i = e / n, j = e % n;
          double acc = 0.0;
          for (int k = 0; k < n; k++)
            acc = fma(Bk[k * nu + i], Pit[k * n + j], acc);
          Y[e] = acc;
        }
        for (int e = tid; e < nu * nu; e += SMPC_NT)
        {
          const int i = e / nu, j = e % nu;
          if (j <= i)
          {
            double acc = kn[L.oR + e];
            for (int k = 0; k < n; k++)
              acc = fma(Bk[k * nu + i], PB[k * nu + j], acc);
            Rh[e] = acc;
          }
        }
        for (int i = tid; i < n; i += SMPC_NT)
        {
          double acc = so[SL.opt + i];
          for (int k = 0; k < n; k++)
            acc = fma(Pt[i * n + k], v[k], acc);
          phi[i] = fma(-mudyn, acc, v[i]);
        }
      }
      SMPC_END
      chol_lower_warp(Rh, nu, nu, bad);
      // X = L^-1 Y (PB is dead: X takes its storage) ; W = L^-1 D^T. bad[1]: some entry of D is non-zero (without force
      // cones no row of the kinodynamics problem depends on the control: W = 0, H = R^-1 and the Sigma block drops out)
      SMPC_PAR(tid)
      {
        for (int e = tid; e < nu * n; e += SMPC_NT)
          X[e] = Y[e];
        bool any = false;
        for (int e = tid; e < nu * nc; e += SMPC_NT)
        {
          const int i = e / nc, c = e % nc;
          const double dv = kn[L.oD + c * nu + i];
          W[e] = dv;
          any = any || dv != 0.0;
        }
        if (any)
          bad[1] = 1;
      }
      SMPC_END
      solve_lower_cols(Rh, nu, nu, X, n, n);
      const bool has_d = nc > 0 && bad[1] != 0; // monotonic over the leg: the structure is the same at every stage
      if (has_d)
      {
        solve_lower_cols(Rh, nu, nu, W, nc, nc);
        // Sigma = mu I + W^T W (lower) ; G = W^T X
        SMPC_PAR(tid)
        {
          for (int e = tid; e < nc * nc; e += SMPC_NT)
          {
            const int i = e / nc, j = e % nc;
            if (j <= i)
            {
              double acc = (i == j) ? mueq : 0.0;
              for (int k = 0; k < nu; k++)
                acc = fma(W[k * nc + i], W[k * nc + j], acc);
              Sg[e] = acc;
            }
          }
          for (int e = tid; e < nc * n; e += SMPC_NT)
          {
            const int i = e / n, j = e % n;
            double acc = 0.0;
            for (int k = 0; k < nu; k++)
              acc = fma(W[k * nc + i], X[k * n + j], acc);
            G[e] = acc;
          }
        }
        SMPC_END
        chol_lower_warp(Sg, nc, nc, bad);
        solve_lower_cols(Sg, nc, nc, G, n, n);
        solve_lower_T_cols(Sg, nc, nc, G, n, n);
        SMPC_PAR(tid)
        for (int e = tid; e < nu * n; e += SMPC_NT)
        {
          const int i = e / n, j = e % n;
          double acc = X[e];
          for (int c = 0; c < nc; c++)
            acc = fma(-W[i * nc + c], G[c * n + j], acc);
          X[e] = acc;
        }
        SMPC_END
      }
      solve_lower_T_cols(Rh, nu, nu, X, n, n);
      // Lam += Y^T X + mu Pi+^T Pi~ ; psi += Pi+^T phi
      SMPC_PAR(tid)
      {
        for (int e = tid; e < n * n; e += SMPC_NT)
        {
          const int i = e / n, j = e % n;
          double acc = Lam[e];
          for (int k = 0; k < nu; k++)
            acc = fma(Y[k * n + i], X[k * n + j], acc);
          double acc2 = 0.0;
          for (int k = 0; k < n; k++)
            acc2 = fma(Pi[k * n + i], Pit[k * n + j], acc2);
          Lam[e] = fma(mudyn, acc2, acc);
        }
        for (int i = tid; i < n; i += SMPC_NT)
        {
          double acc = psi[i];
          for (int k = 0; k < n; k++)
            acc = fma(Pi[k * n + i], phi[k], acc);
          psi[i] = acc;
        }
      }
      SMPC_END
      // Pi = A^T Pi~ + K^T Y
      SMPC_PAR(tid)
      for (int e = tid; e < n * n; e += SMPC_NT)
      {
        const int i = e / n, j = e % n;
        double acc = 0.0;
        for (int k = 0; k < n; k++)
          acc = fma(Ak[k * n + i], Pit[k * n + j], acc);
        for (int k = 0; k < nu; k++)
          acc = fma(Kk[k * n + i], Y[k * n + j], acc);
        Pi[e] = acc;
      }
      SMPC_END
    }
    SMPC_PAR(tid)
    {
      for (int e = tid; e < n * n; e += SMPC_NT)
      {
        const int i = e / n, j = e % n;
        out[e] = Pi[e];
        out[n * n + e] = 0.5 * (Lam[i * n + j] + Lam[j * n + i]);
      }
      for (int i = tid; i < n; i += SMPC_NT)
        out[2 * n * n + i] = psi[i];
      if (tid == 0 && flag_out)
        *flag_out = *bad;
    }
    SMPC_END
  }

  // pass 3 for one instance: legP0 [J][g(g+1)] = (P_a, p_a) of the zero-terminal sweeps (legs 1..J-1), legE [J][elem] the
  // elements of legs 1..J-2 -> legT [J][g(g+1)]: legT[j] = true cost-to-go at a_j (j = 1..J-1), the terminal of leg j-1
  SMPC_DEV void leg_coarse_inst(
    int n, int J, const CoarseSmem & RS, const double * __restrict__ legP0, const double * __restrict__ legE,
    double * __restrict__ legT, double * sm, int * flag_out)
  {
    const int ldp = n + 1, pp = leg_pp_size(n), es = leg_elem_size(n);
    double * Cc = sm + RS.oC;
    double * LC = sm + RS.oLC;
    double * S = sm + RS.oS;
    double * R = sm + RS.oR;
    double * U = sm + RS.oU;
    double * Pn = sm + RS.oPn; // [P | p] of the next boundary
    int * bad = (int *)(sm + RS.oflag);
    SMPC_PAR(tid)
    {
      if (tid == 0)
        *bad = 0;
      for (int e = tid; e < pp; e += SMPC_NT)
      {
        const double val = legP0[(size_t)(J - 1) * pp + e];
        Pn[e] = val;
        legT[(size_t)(J - 1) * pp + e] = val;
      }
    }
    SMPC_END
    for (int j = J - 2; j >= 1; j--)
    {
      const double * Pi = legE + (size_t)j * es;
      const double * Lam = Pi + n * n;
      const double * psi = Lam + n * n;
      const double * Pa = legP0 + (size_t)j * pp;
      // C = chol(Pn) (lower, upper part zeroed)
      SMPC_PAR(tid)
      for (int e = tid; e < n * n; e += SMPC_NT)
      {
        const int i = e / n, c = e % n;
        Cc[e] = (c <= i) ? Pn[i * ldp + c] : 0.0;
      }
      SMPC_END
      chol_lower_warp(Cc, n, n, bad);
      // LC = Lam C ; R = [Psi | psi - Lam pn] (Psi = Pi^T)
      SMPC_PAR(tid)
      {
        for (int e = tid; e < n * n; e += SMPC_NT)
        {
          const int i = e / n, c = e % n;
          double acc = 0.0;
          for (int k = c; k < n; k++) // C is lower triangular
            acc = fma(Lam[i * n + k], Cc[k * n + c], acc);
          LC[e] = acc;
        }
        for (int e = tid; e < n * ldp; e += SMPC_NT)
        {
          const int i = e / ldp, c = e % ldp;
          if (c < n)
            R[e] = Pi[c * n + i];
          else
          {
            double acc = psi[i];
            for (int k = 0; k < n; k++)
              acc = fma(-Lam[i * n + k], Pn[k * ldp + n], acc);
            R[e] = acc;
          }
        }
      }
      SMPC_END
      // S = I + C^T LC (lower) ; U = C^T R
      SMPC_PAR(tid)
      {
        for (int e = tid; e < n * n; e += SMPC_NT)
        {
          const int i = e / n, c = e % n;
          if (c <= i)
          {
            double acc1 = (i == c) ? 1.0 : 0.0, acc2 = acc1;
            for (int k = i; k < n; k++)
              acc1 = fma(Cc[k * n + i], LC[k * n + c], acc1);
            for (int k = c; k < n; k++)
              acc2 = fma(Cc[k * n + c], LC[k * n + i], acc2);
            S[e] = 0.5 * (acc1 + acc2);
          }
        }
        for (int e = tid; e < n * ldp; e += SMPC_NT)
        {
          const int i = e / ldp, c = e % ldp;
          double acc = 0.0;
          for (int k = i; k < n; k++)
            acc = fma(Cc[k * n + i], R[k * ldp + c], acc);
          U[e] = acc;
        }
      }
      SMPC_END
      chol_lower_warp(S, n, n, bad);
      solve_lower_cols(S, n, n, U, ldp, ldp);
      solve_lower_T_cols(S, n, n, U, ldp, ldp);
      // T = C U (+ pn on the vector column), into R
      SMPC_PAR(tid)
      for (int e = tid; e < n * ldp; e += SMPC_NT)
      {
        const int i = e / ldp, c = e % ldp;
        double acc = (c == n) ? Pn[i * ldp + n] : 0.0;
        for (int k = 0; k <= i; k++)
          acc = fma(Cc[i * n + k], U[k * ldp + c], acc);
        R[e] = acc;
      }
      SMPC_END
      // [P | p] = [Pa | pa] + Pi T, into U ; then symmetrised into Pn and legT[j]
      SMPC_PAR(tid)
      for (int e = tid; e < n * ldp; e += SMPC_NT)
      {
        const int i = e / ldp, c = e % ldp;
        double acc = Pa[e];
        for (int k = 0; k < n; k++)
          acc = fma(Pi[i * n + k], R[k * ldp + c], acc);
        U[e] = acc;
      }
      SMPC_END
      SMPC_PAR(tid)
      for (int e = tid; e < n * ldp; e += SMPC_NT)
      {
        const int i = e / ldp, c = e % ldp;
        const double val = (c < n && c != i) ? 0.5 * (U[i * ldp + c] + U[c * ldp + i]) : U[e];
        Pn[e] = val;
        legT[(size_t)j * pp + e] = val;
      }
      SMPC_END
    }
    SMPC_PAR(tid)
    if (tid == 0 && flag_out)
      *flag_out = *bad;
    SMPC_END
  }
} // namespace smpc
