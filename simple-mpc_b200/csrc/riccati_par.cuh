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
  // ---- small dense products of the leg-element / condensation kernels. These kernels sit on the critical path of a
  // latency-bound tick: a product written as one fused-multiply-add chain per output (operands from shared memory / L1) runs
  // at the load latency (measured: 99 us per element stage, 169 us per condensation step - slower than the sweep they are
  // meant to shorten). 2 x 2 register tiles with the k loop unrolled by four give every load two uses and every thread four
  // independent chains. ----
  // this thread's share of C = op(A) op(B): A(i,k) = pa[i * sai + k * sak], B(k,j) = pb[k * sbk + j * sbj]; epi(i, j, sum)
  template <class Epi>
  SMPC_DEV void gemm22(
    int tid, int nt, int M, int N, int K, const double * __restrict__ pa, int sai, int sak, const double * __restrict__ pb, int sbk,
    int sbj, Epi epi)
  {
    const int TN = (N + 1) / 2, tiles = ((M + 1) / 2) * TN;
    for (int e = tid; e < tiles; e += nt)
    {
      const int i0 = 2 * (e / TN), j0 = 2 * (e % TN);
      const int i1 = (i0 + 1 < M) ? i0 + 1 : i0, j1 = (j0 + 1 < N) ? j0 + 1 : j0;
      const double * a0 = pa + i0 * sai;
      const double * a1 = pa + i1 * sai;
      const double * b0 = pb + j0 * sbj;
      const double * b1 = pb + j1 * sbj;
      double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll 4
      for (int k = 0; k < K; k++)
      {
        const double x0 = a0[k * sak], x1 = a1[k * sak], y0 = b0[k * sbk], y1 = b1[k * sbk];
        c00 = fma(x0, y0, c00);
        c01 = fma(x0, y1, c01);
        c10 = fma(x1, y0, c10);
        c11 = fma(x1, y1, c11);
      }
      epi(i0, j0, c00);
      if (j1 != j0)
        epi(i0, j1, c01);
      if (i1 != i0)
      {
        epi(i1, j0, c10);
        if (j1 != j0)
          epi(i1, j1, c11);
      }
    }
  }

  // Right-looking LDL^T of a symmetric positive definite S (lower triangle, in place) whose elimination carries up to two
  // blocks of right-hand sides along: U <- L^-1 U with the UNIT lower factor, ONE CTA barrier per pivot (the scale-then-update
  // routines of riccati.cuh take two, and a separate forward substitution two more). On return the strict lower triangle
  // of S holds L D (column k = L(:,k) d_k), the diagonal d, and invd[k] = 1 / d_k.
  SMPC_DEV void ldl_fused(
    double * S, int n, int lds, double * invd, double * U1, int m1, int ld1, double * U2, int m2, int ld2, int * bad)
  {
    for (int k = 0; k < n; k++)
    {
      SMPC_PAR(tid)
      {
        double dk = S[k * lds + k];
        if (!(dk > 0.0))
        {
          if (tid == 0)
            *bad = 1;
          dk = 1.0;
        }
        const double id = 1.0 / dk;
        if (tid == 0)
          invd[k] = id;
        // a thread owns ONE column of the trailing block [S(:, k+1..) | U1 | U2] and every g-th row of it: the pivot-row entry
        // of its column is loaded once, and the only index arithmetic per element is an increment
        const int rows = n - k - 1, w = rows + m1 + m2;
        if (w > 0 && tid < (SMPC_NT / w) * w)
        {
          const int g = SMPC_NT / w, c = tid % w, r0 = tid / w;
          double * col;
          int ld, jrow = -1;
          if (c < rows)
          {
            col = S + (k + 1 + c), ld = lds, jrow = k + 1 + c; // column j of S: rows i >= j only
          }
          else if (c < rows + m1)
            col = U1 + (c - rows), ld = ld1;
          else
            col = U2 + (c - rows - m1), ld = ld2;
          const double piv = (jrow >= 0) ? S[jrow * lds + k] : col[k * ld];
          for (int i = k + 1 + r0; i < n; i += g)
            if (i >= jrow)
              col[i * ld] = fma(-(S[i * lds + k] * id), piv, col[i * ld]);
        }
      }
      SMPC_END
    }
  }
  // U <- L^-T U with the unit lower factor ldl_fused left in S: one barrier per pivot
  SMPC_DEV void ldl_backward(const double * S, int n, int lds, const double * invd, double * U, int m, int ld)
  {
    for (int k = n - 1; k >= 1; k--)
    {
      SMPC_PAR(tid)
      if (tid < (SMPC_NT / m) * m)
      {
        const int g = SMPC_NT / m, c = tid % m;
        const double piv = U[k * ld + c];
        for (int i = tid / m; i < k; i += g)
          U[i * ld + c] = fma(-(S[k * lds + i] * invd[i]), piv, U[i * ld + c]);
      }
      SMPC_END
    }
  }
  // rows of U scaled by invd[k] (power = 2) or sqrt(invd[k]) (power = 1)
  SMPC_DEV void scale_rows(double * U, int n, int m, int ld, const double * invd, int power)
  {
    SMPC_PAR(tid)
    for (int e = tid; e < n * m; e += SMPC_NT)
    {
      const int i = e / m, c = e - i * m;
      U[i * ld + c] *= (power == 2) ? invd[i] : sqrt(invd[i]);
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
    double * invr = sm + RS.oinvd;
    double * invs = invr + nu;
    double * sPt = sm + RS.osPt;
    double * sA = sm + RS.osA;
    double * sB = sm + RS.osB;
    double * sK = sm + RS.osK;
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
      const int nc = ncs[t];
      // stage the operands (coalesced, all loads independent): with the products reading global memory the kernel spent
      // 5 warp-cycles per issue waiting on the long scoreboard (profiles/r02y)
      SMPC_PAR(tid)
      {
        for (int e = tid; e < n * n; e += SMPC_NT)
        {
          sPt[e] = so[SL.oPt + e];
          sA[e] = kn[L.oA + e];
        }
        for (int e = tid; e < n * nu; e += SMPC_NT)
        {
          sB[e] = kn[L.oB + e];
          sK[e] = so[SL.oK + e];
        }
      }
      SMPC_END
      const double * Pt = sPt;
      const double * Kk = sK;
      const double * Ak = sA;
      const double * Bk = sB;
      // Pi~ = Pi+ - mu P~ Pi+ ; P~ B ; v = B k + f
      SMPC_PAR(tid)
      {
        gemm22(tid, SMPC_NT, n, n, n, Pt, n, 1, Pi, n, 1, [&](int i, int j, double acc) { Pit[i * n + j] = fma(-mudyn, acc, Pi[i * n + j]); });
        gemm22(tid, SMPC_NT, n, nu, n, Pt, n, 1, Bk, nu, 1, [&](int i, int j, double acc) { PB[i * nu + j] = acc; });
        for (int i = tid; i < n; i += SMPC_NT)
        {
          double acc = kn[L.of + i];
          for (int k = 0; k < nu; k++)
            acc = fma(Bk[i * nu + k], so[SL.ok + k], acc);
          v[i] = acc;
        }
      }
      SMPC_END
      // Y = B^T Pi~ ; R^ = R + B^T (P~ B) ; phi = v - mu (P~ v + p~)
      SMPC_PAR(tid)
      {
        gemm22(tid, SMPC_NT, nu, n, n, Bk, 1, nu, Pit, n, 1, [&](int i, int j, double acc) { Y[i * n + j] = acc; });
        gemm22(tid, SMPC_NT, nu, nu, n, Bk, 1, nu, PB, nu, 1, [&](int i, int j, double acc) { Rh[i * nu + j] = kn[L.oR + i * nu + j] + acc; });
        for (int i = tid; i < n; i += SMPC_NT)
        {
          double acc = so[SL.opt + i];
          for (int k = 0; k < n; k++)
            acc = fma(Pt[i * n + k], v[k], acc);
          phi[i] = fma(-mudyn, acc, v[i]);
        }
      }
      SMPC_END
      // X = Y (PB is dead: X takes its storage) ; W = D^T. bad[1]: some entry of D is non-zero (without force cones no row
      // of the kinodynamics problem depends on the control: W = 0, H = R^-1 and the Sigma block drops out)
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
      const bool has_d = nc > 0 && bad[1] != 0; // monotonic over the leg: the structure is the same at every stage
      // R^ = L D L^T ; X <- L^-1 Y, W <- L^-1 D^T in the same elimination ; H = L^-T [D^-1 - D^-1 W Sigma^-1 W^T D^-1] L^-1
      ldl_fused(Rh, nu, nu, invr, X, n, n, W, has_d ? nc : 0, nc, bad);
      if (has_d)
      {
        scale_rows(X, nu, n, n, invr, 1);
        scale_rows(W, nu, nc, nc, invr, 1);
        // Sigma = mu I + W^T W ; G = W^T X
        SMPC_PAR(tid)
        {
          gemm22(tid, SMPC_NT, nc, nc, nu, W, 1, nc, W, nc, 1, [&](int i, int j, double acc) { Sg[i * nc + j] = acc + ((i == j) ? mueq : 0.0); });
          gemm22(tid, SMPC_NT, nc, n, nu, W, 1, nc, X, n, 1, [&](int i, int j, double acc) { G[i * n + j] = acc; });
        }
        SMPC_END
        ldl_fused(Sg, nc, nc, invs, G, n, n, nullptr, 0, 0, bad);
        scale_rows(G, nc, n, n, invs, 2);
        ldl_backward(Sg, nc, nc, invs, G, n, n);
        SMPC_PAR(tid)
        gemm22(tid, SMPC_NT, nu, n, nc, W, nc, 1, G, n, 1, [&](int i, int j, double acc) { X[i * n + j] -= acc; });
        SMPC_END
        scale_rows(X, nu, n, n, invr, 1);
      }
      else
        scale_rows(X, nu, n, n, invr, 2);
      ldl_backward(Rh, nu, nu, invr, X, n, n);
      // Lam += Y^T X + mu Pi+^T Pi~ ; psi += Pi+^T phi
      SMPC_PAR(tid)
      {
        gemm22(tid, SMPC_NT, n, n, nu, Y, 1, n, X, n, 1, [&](int i, int j, double acc) { Lam[i * n + j] += acc; });
        for (int i = tid; i < n; i += SMPC_NT)
        {
          double acc = psi[i];
          for (int k = 0; k < n; k++)
            acc = fma(Pi[k * n + i], phi[k], acc);
          psi[i] = acc;
        }
      }
      SMPC_END
      SMPC_PAR(tid)
      gemm22(tid, SMPC_NT, n, n, n, Pi, 1, n, Pit, n, 1, [&](int i, int j, double acc) { Lam[i * n + j] = fma(mudyn, acc, Lam[i * n + j]); });
      SMPC_END
      // Pi = A^T Pi~ + K^T Y (Pi+ is dead)
      SMPC_PAR(tid)
      gemm22(tid, SMPC_NT, n, n, n, Ak, 1, n, Pit, n, 1, [&](int i, int j, double acc) { Pi[i * n + j] = acc; });
      SMPC_END
      SMPC_PAR(tid)
      gemm22(tid, SMPC_NT, n, n, nu, Kk, 1, n, Y, n, 1, [&](int i, int j, double acc) { Pi[i * n + j] += acc; });
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
    double * wv = sm + RS.owv; // inverse pivots
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
      ldl_fused(Cc, n, n, wv, nullptr, 0, 0, nullptr, 0, 0, bad);
      SMPC_PAR(tid)
      for (int e = tid; e < n * n; e += SMPC_NT)
      {
        const int i = e / n, c = e % n;
        if (c <= i)
          Cc[e] = (c == i) ? 1.0 / sqrt(wv[c]) : Cc[e] * sqrt(wv[c]); // C = L D^1/2
      }
      SMPC_END
      // LC = Lam C ; R = [Psi | psi - Lam pn] (Psi = Pi^T)
      SMPC_PAR(tid)
      {
        gemm22(tid, SMPC_NT, n, n, n, Lam, n, 1, Cc, n, 1, [&](int i, int c, double acc) { LC[i * n + c] = acc; });
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
      // S = I + C^T LC ; U = C^T R
      SMPC_PAR(tid)
      {
        gemm22(tid, SMPC_NT, n, n, n, Cc, 1, n, LC, n, 1, [&](int i, int c, double acc) { S[i * n + c] = acc + ((i == c) ? 1.0 : 0.0); });
        gemm22(tid, SMPC_NT, n, ldp, n, Cc, 1, n, R, ldp, 1, [&](int i, int c, double acc) { U[i * ldp + c] = acc; });
      }
      SMPC_END
      SMPC_PAR(tid)
      for (int e = tid; e < n * n; e += SMPC_NT)
      {
        const int i = e / n, c = e % n;
        if (c < i)
          S[e] = 0.5 * (S[e] + S[c * n + i]); // the factorisation reads the lower triangle only
      }
      SMPC_END
      ldl_fused(S, n, n, wv, U, ldp, ldp, nullptr, 0, 0, bad);
      scale_rows(U, n, ldp, ldp, wv, 2);
      ldl_backward(S, n, n, wv, U, ldp, ldp);
      // T = C U (+ pn on the vector column), into R
      SMPC_PAR(tid)
      gemm22(tid, SMPC_NT, n, ldp, n, Cc, n, 1, U, ldp, 1, [&](int i, int c, double acc) { R[i * ldp + c] = acc + ((c == n) ? Pn[i * ldp + n] : 0.0); });
      SMPC_END
      // [P | p] = [Pa | pa] + Pi T, into U ; then symmetrised into Pn and legT[j]
      SMPC_PAR(tid)
      gemm22(tid, SMPC_NT, n, ldp, n, Pi, n, 1, R, ldp, 1, [&](int i, int c, double acc) { U[i * ldp + c] = Pa[i * ldp + c] + acc; });
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
