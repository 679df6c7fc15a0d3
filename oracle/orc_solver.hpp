// ORACLE — test infrastructure only (see orc_math.hpp header). PARITY UNPINNED.
//
// orc_solver.hpp — the ProxDDP iteration that reference src/mpc.cpp:85,212 calls
// (aligator::SolverProxDDPTpl<double>::run, [UPSTREAM] aligator 0.16.0, pixi.lock:11) restated from
// the published algorithm (Jallet et al., "ProxDDP: Proximal Constrained Trajectory Optimization",
// 2023; Jallet et al., "Parallel and Proximal Constrained LQ methods for Real-Time Nonlinear MPC",
// RSS 2024) with the settings the reference selects (src/mpc.cpp:43-53): LINEAR rollout,
// force_initial_condition, serial proximal Riccati, Gauss-Newton Hessians, nonmonotone linesearch.
// Every constant that comes from upstream defaults is in smpc_solver_opts.
//
// The exact operation order of the Riccati recursion below is the CANONICAL order: the CUDA
// kernels follow it term by term (ascending-k fused multiply-adds, structural zeros skipped) so
// that the two agree to rounding.
#pragma once
#include "orc_ocp.hpp"
#include <algorithm>
#include <chrono>

namespace orc
{

  struct Knot
  {
    int nc = 0;
    Mat Q, S, R, A, B, C, D;
    Vec q, r, f, d;
  };
  struct StageSol
  {
    Mat K, Z, P, Pt; // du = k + K dx ; nu+ = z + Z dx ; cost-to-go ; P~ of the NEXT node
    Vec k, z, p, pt;
  };

  // Loop nests below are written (row, k, column): the innermost loop runs over CONTIGUOUS columns and every output
  // element still receives its fused multiply-adds in ascending k - the canonical per-element order the kernels follow -
  // so the compiler can vectorise them (liborc_fast.so: AVX2 + FMA) without reassociating anything: results are
  // bitwise those of the plain dot-product form.

  // in-place Cholesky (lower) of an n x n row-major SPD matrix; returns false if a pivot is <= 0.
  // Right-looking: entry (i,c) receives fma(-L(i,j), L(c,j), .) for ascending j, then the scaling by 1/L(c,c).
  inline bool chol_lower(Mat & M, int n)
  {
    Vec col(n > 0 ? n : 1);
    for (int j = 0; j < n; j++)
    {
      const double s = M(j, j);
      if (!(s > 0))
        return false;
      const double l = std::sqrt(s);
      M(j, j) = l;
      const double inv = 1.0 / l;
      for (int i = j + 1; i < n; i++)
      {
        M(i, j) = M(i, j) * inv;
        col[i] = M(i, j);
      }
      for (int i = j + 1; i < n; i++)
      {
        const double lij = col[i];
        double * Mi = &M(i, 0);
        for (int c = j + 1; c <= i; c++)
          Mi[c] = std::fma(-lij, col[c], Mi[c]);
      }
    }
    return true;
  }
  // X <- L^-1 X  (L lower n x n, X n x m)
  inline void solve_lower(const Mat & L, int n, Mat & X, int m)
  {
    for (int i = 0; i < n; i++)
    {
      double * Xi = &X(i, 0);
      for (int k = 0; k < i; k++)
      {
        const double lik = L(i, k);
        const double * Xk = &X(k, 0);
        for (int c = 0; c < m; c++)
          Xi[c] = std::fma(-lik, Xk[c], Xi[c]);
      }
      const double inv = 1.0 / L(i, i); // reciprocal pivot: the kernels keep 1/l, not l
      for (int c = 0; c < m; c++)
        Xi[c] = Xi[c] * inv;
    }
  }
  // X <- L^-T X
  inline void solve_lower_T(const Mat & L, int n, Mat & X, int m)
  {
    for (int i = n - 1; i >= 0; i--)
    {
      double * Xi = &X(i, 0);
      for (int k = n - 1; k > i; k--)
      {
        const double lki = L(k, i);
        const double * Xk = &X(k, 0);
        for (int c = 0; c < m; c++)
          Xi[c] = std::fma(-lki, Xk[c], Xi[c]);
      }
      const double inv = 1.0 / L(i, i); // reciprocal pivot: the kernels keep 1/l, not l
      for (int c = 0; c < m; c++)
        Xi[c] = Xi[c] * inv;
    }
  }
  // C(a, 0..nb) += sum_k A(k, a) B(k, 0..nb)   (A^T B, ascending k per element)
  inline void add_AtB(const Mat & A, const Mat & B, int nk, int na, int nb, Mat & Cm)
  {
    if (nk <= 0 || na <= 0 || nb <= 0)
      return;
    for (int a = 0; a < na; a++)
    {
      double * Ca = &Cm(a, 0);
      for (int k = 0; k < nk; k++)
      {
        const double aka = A(k, a);
        const double * Bk = &B.d[(size_t)k * B.c];
        for (int b = 0; b < nb; b++)
          Ca[b] = std::fma(aka, Bk[b], Ca[b]);
      }
    }
  }
  // C(a, 0..nb) += sum_k A(a, k) B(k, 0..nb)
  inline void add_AB(const Mat & A, const Mat & B, int na, int nk, int nb, Mat & Cm)
  {
    if (nk <= 0 || na <= 0 || nb <= 0)
      return;
    for (int a = 0; a < na; a++)
    {
      double * Ca = &Cm(a, 0);
      for (int k = 0; k < nk; k++)
      {
        const double aak = A(a, k);
        const double * Bk = &B.d[(size_t)k * B.c];
        for (int b = 0; b < nb; b++)
          Ca[b] = std::fma(aak, Bk[b], Ca[b]);
      }
    }
  }

  struct RiccatiStats
  {
    bool ok = true;
  };

  // Backward sweep of the proximal Riccati recursion (gar::ProximalRiccatiSolver::backward(mudyn, mueq)).
  // knots[0..N-1] are stages, knots[N] the terminal node (Q,q,C,d only).
  // Stages [t0, t1) only when t1 >= 0; a given cost-to-go (Pterm, pterm) at node t1 replaces the terminal-node computation
  // (the legs of riccati_backward_parallel).
  inline RiccatiStats riccati_backward(
    const std::vector<Knot> & knots, int ndx, int nu, double mudyn, double mueq, std::vector<StageSol> & sol, int t0 = 0, int t1 = -1,
    const Mat * Pterm = nullptr, const Vec * pterm = nullptr)
  {
    RiccatiStats st;
    const int N = (t1 >= 0) ? t1 : (int)knots.size() - 1;
    if ((int)sol.size() != (int)knots.size())
      sol.resize(knots.size());
    const double inv_mu = 1.0 / mueq;
    if (Pterm)
    {
      sol[N].P = *Pterm;
      sol[N].p = *pterm;
    }
    else
    // terminal: nu_N = (d + C dx)/mu ; P = Q + C^T Z ; p = q + C^T z
    {
      const Knot & kn = knots[N];
      StageSol & s = sol[N];
      const int nc = kn.nc;
      s.Z.resize(nc, ndx);
      s.z.assign(nc, 0);
      for (int i = 0; i < nc; i++)
      {
        s.z[i] = kn.d[i] * inv_mu;
        for (int j = 0; j < ndx; j++)
          s.Z(i, j) = kn.C(i, j) * inv_mu;
      }
      s.P = kn.Q;
      s.p = kn.q;
      add_AtB(kn.C, s.Z, nc, ndx, ndx, s.P);
      for (int a = 0; a < ndx; a++)
      {
        double acc = s.p[a];
        for (int i = 0; i < nc; i++)
          acc = std::fma(kn.C(i, a), s.z[i], acc);
        s.p[a] = acc;
      }
    }
    for (int t = N - 1; t >= t0; t--)
    {
      const Knot & kn = knots[t];
      const StageSol & nx = sol[t + 1];
      StageSol & s = sol[t];
      const int nc = kn.nc;
      // 1. P~ = (I + mudyn P+)^-1 P+ , p~ = (I + mudyn P+)^-1 p+
      s.Pt = nx.P;
      s.pt = nx.p;
      if (mudyn > 0)
      {
        Mat M(ndx, ndx);
        for (int a = 0; a < ndx; a++)
          for (int b = 0; b < ndx; b++)
            M(a, b) = mudyn * nx.P(a, b) + (a == b ? 1.0 : 0.0);
        if (!chol_lower(M, ndx))
          st.ok = false;
        Mat rhs(ndx, ndx + 1);
        for (int a = 0; a < ndx; a++)
        {
          for (int b = 0; b < ndx; b++)
            rhs(a, b) = nx.P(a, b);
          rhs(a, ndx) = nx.p[a];
        }
        solve_lower(M, ndx, rhs, ndx + 1);
        solve_lower_T(M, ndx, rhs, ndx + 1);
        for (int a = 0; a < ndx; a++)
        {
          for (int b = 0; b < ndx; b++)
            s.Pt(a, b) = rhs(a, b);
          s.pt[a] = rhs(a, ndx);
        }
      }
      // 2. w = P~ f + p~
      Vec w(ndx);
      for (int a = 0; a < ndx; a++)
      {
        double acc = s.pt[a];
        for (int b = 0; b < ndx; b++)
          acc = std::fma(s.Pt(a, b), kn.f[b], acc);
        w[a] = acc;
      }
      // 3. PA = P~ A, PB = P~ B
      Mat PA(ndx, ndx), PB(ndx, nu);
      add_AB(s.Pt, kn.A, ndx, ndx, ndx, PA);
      add_AB(s.Pt, kn.B, ndx, ndx, nu, PB);
      // 4. Qh = Q + A^T PA ; Sh = S + A^T PB ; Rh = R + B^T PB ; qh = q + A^T w ; rh = r + B^T w
      Mat Qh = kn.Q, Sh = kn.S, Rh = kn.R;
      Vec qh(ndx), rh(nu);
      add_AtB(kn.A, PA, ndx, ndx, ndx, Qh);
      add_AtB(kn.A, PB, ndx, ndx, nu, Sh);
      add_AtB(kn.B, PB, ndx, nu, nu, Rh);
      for (int a = 0; a < ndx; a++)
      {
        double acc = kn.q[a];
        for (int k = 0; k < ndx; k++)
          acc = std::fma(kn.A(k, a), w[k], acc);
        qh[a] = acc;
      }
      for (int a = 0; a < nu; a++)
      {
        double acc = kn.r[a];
        for (int k = 0; k < ndx; k++)
          acc = std::fma(kn.B(k, a), w[k], acc);
        rh[a] = acc;
      }
      // 5. quasi-definite KKT  [[Rh, D^T],[D, -mu I]] [u; nu] = -[[rh, Sh^T],[d, C]]  by block LDL^T:
      //    Rh = L L^T ; Y = L^-1 [rh Sh^T] ; W = L^-1 D^T ; Sigma = mu I + W^T W = Ls Ls^T ;
      //    nu = Sigma^-1 ([d C] - W^T Y) ; u = -L^-T (Y + W nu)
      if (!chol_lower(Rh, nu))
        st.ok = false;
      Mat Y(nu, ndx + 1);
      for (int a = 0; a < nu; a++)
      {
        Y(a, 0) = rh[a];
        for (int b = 0; b < ndx; b++)
          Y(a, 1 + b) = Sh(b, a);
      }
      solve_lower(Rh, nu, Y, ndx + 1);
      Mat G(nc, ndx + 1); // multiplier gains: col 0 = z, cols 1.. = Z
      if (nc > 0)
      {
        Mat W(nu, nc);
        for (int a = 0; a < nu; a++)
          for (int i = 0; i < nc; i++)
            W(a, i) = kn.D(i, a);
        solve_lower(Rh, nu, W, nc);
        Mat Sg(nc, nc);
        for (int i = 0; i < nc; i++)
        {
          double * Si = &Sg(i, 0);
          Si[i] = mueq;
          for (int k = 0; k < nu; k++)
          {
            const double wki = W(k, i);
            const double * Wk = &W(k, 0);
            for (int j = 0; j <= i; j++)
              Si[j] = std::fma(wki, Wk[j], Si[j]);
          }
        }
        for (int i = 0; i < nc; i++)
          for (int j = 0; j < i; j++)
            Sg(j, i) = Sg(i, j);
        if (!chol_lower(Sg, nc))
          st.ok = false;
        for (int i = 0; i < nc; i++)
        {
          double * Gi = &G(i, 0);
          Gi[0] = kn.d[i];
          for (int c = 0; c < ndx; c++)
            Gi[1 + c] = kn.C(i, c);
          for (int k = 0; k < nu; k++)
          {
            const double wki = -W(k, i);
            const double * Yk = &Y(k, 0);
            for (int c = 0; c < ndx + 1; c++)
              Gi[c] = std::fma(wki, Yk[c], Gi[c]);
          }
        }
        solve_lower(Sg, nc, G, ndx + 1);
        solve_lower_T(Sg, nc, G, ndx + 1);
        add_AB(W, G, nu, nc, ndx + 1, Y);
      }
      solve_lower_T(Rh, nu, Y, ndx + 1);
      s.K.resize(nu, ndx);
      s.k.assign(nu, 0);
      s.Z.resize(nc, ndx);
      s.z.assign(nc, 0);
      for (int a = 0; a < nu; a++)
      {
        s.k[a] = -Y(a, 0);
        for (int b = 0; b < ndx; b++)
          s.K(a, b) = -Y(a, 1 + b);
      }
      for (int i = 0; i < nc; i++)
      {
        s.z[i] = G(i, 0);
        for (int b = 0; b < ndx; b++)
          s.Z(i, b) = G(i, 1 + b);
      }
      // 6. P = Qh + Sh K + C^T Z (symmetrised) ; p = qh + Sh k + C^T z
      s.P = Qh;
      s.p.assign(ndx, 0);
      add_AB(Sh, s.K, ndx, nu, ndx, s.P);
      add_AtB(kn.C, s.Z, nc, ndx, ndx, s.P);
      for (int a = 0; a < ndx; a++)
      {
        double acc = qh[a];
        for (int k = 0; k < nu; k++)
          acc = std::fma(Sh(a, k), s.k[k], acc);
        for (int i = 0; i < nc; i++)
          acc = std::fma(kn.C(i, a), s.z[i], acc);
        s.p[a] = acc;
      }
      for (int a = 0; a < ndx; a++)
        for (int b = 0; b < a; b++)
        {
          double m = 0.5 * (s.P(a, b) + s.P(b, a));
          s.P(a, b) = m;
          s.P(b, a) = m;
        }
    }
    return st;
  }

  // ---- PARALLEL (leg-condensed) backward sweep: aligator's gar::ParallelRiccatiSolver, selected by the reference when
  // num_threads > 1 (reference src/mpc.cpp:46-50), restated for one CTA per leg (simple-mpc_b200/csrc/riccati_par.cuh).
  // The horizon is cut into J legs [a_j, a_{j+1}). With the co-state theta at the end of a leg as a parameter (cost-to-go
  // P = 0, p = theta there) the leg's end state is affine in its start state and theta:
  //     x_b = Psi x_a + psi - Lam theta ,   lambda_a = P_a x_a + p_a + Psi^T theta
  // (Psi = product of the closed-loop transition matrices, Lam >= 0 the reduced dual Hessian). Four passes:
  //   1. legs 1..J-1 in parallel: the ordinary backward sweep with zero cost-to-go at the leg end (true terminal node for the
  //      last leg) -> gains, P~, and (P_a, p_a);
  //   2. legs 1..J-2 in parallel: (Psi, psi, Lam) from the stored gains, stage by stage:
  //        Pi~ = (I - mu P~) Pi+ ; Y = B^T Pi~ ; X = H Y with H = the (u,u) block of the stage's KKT inverse ;
  //        Lam += Y^T X + mu Pi+^T Pi~ ; psi += Pi+^T [(I - mu P~)(B k + f) - mu p~] ; Pi = A^T Pi~ + K^T Y   (Pi = Psi^T)
  //   3. serial over the J-2 inner boundaries: the TRUE cost-to-go at every leg start,
  //        calP_j = P_j + Psi_j^T calP_{j+1} (I + Lam_j calP_{j+1})^-1 Psi_j   (and the vector part);
  //   4. legs 0..J-2 in parallel: the ordinary backward sweep again, now from the true cost-to-go.
  // After pass 4 `sol` holds what the serial sweep computes, up to rounding in the hand-over of calP.
  inline void leg_bounds(int N, int J, std::vector<int> & a)
  {
    a.resize(J + 1);
    for (int j = 0; j <= J; j++)
      a[j] = (int)(((long long)j * N) / J);
  }
  // LU with partial pivoting, in place; solves M X = R for the m columns of R (row-major), returns false if singular
  inline bool lu_solve(Mat & M, int n, Mat & R, int m)
  {
    for (int k = 0; k < n; k++)
    {
      int piv = k;
      double best = std::fabs(M(k, k));
      for (int i = k + 1; i < n; i++)
        if (std::fabs(M(i, k)) > best)
          best = std::fabs(M(i, k)), piv = i;
      if (best == 0.0)
        return false;
      if (piv != k)
      {
        for (int j = 0; j < n; j++)
          std::swap(M(k, j), M(piv, j));
        for (int j = 0; j < m; j++)
          std::swap(R(k, j), R(piv, j));
      }
      const double inv = 1.0 / M(k, k);
      for (int i = k + 1; i < n; i++)
      {
        const double l = M(i, k) * inv;
        M(i, k) = l;
        for (int j = k + 1; j < n; j++)
          M(i, j) = std::fma(-l, M(k, j), M(i, j));
        for (int j = 0; j < m; j++)
          R(i, j) = std::fma(-l, R(k, j), R(i, j));
      }
    }
    for (int k = n - 1; k >= 0; k--)
    {
      const double inv = 1.0 / M(k, k);
      for (int j = 0; j < m; j++)
      {
        double acc = R(k, j);
        for (int i = k + 1; i < n; i++)
          acc = std::fma(-M(k, i), R(i, j), acc);
        R(k, j) = acc * inv;
      }
    }
    return true;
  }
  struct LegElement
  {
    Mat Pi, Lam; // Pi = Psi^T
    Vec psi;
  };
  // pass 2 for the leg [a, b): reads the gains the zero-terminal sweep left in sol
  inline bool leg_element(
    const std::vector<Knot> & knots, const std::vector<StageSol> & sol, int ndx, int nu, double mudyn, double mueq, int a, int b,
    LegElement & E)
  {
    bool ok = true;
    E.Pi.resize(ndx, ndx);
    E.Lam.resize(ndx, ndx);
    E.psi.assign(ndx, 0.0);
    for (int i = 0; i < ndx; i++)
      E.Pi(i, i) = 1.0;
    Mat Pit(ndx, ndx), Y(nu, ndx), X(nu, ndx), PB(ndx, nu), Rh(nu, nu), Pin(ndx, ndx);
    Vec v(ndx), phi(ndx);
    for (int t = b - 1; t >= a; t--)
    {
      const Knot & kn = knots[t];
      const StageSol & s = sol[t];
      const int nc = kn.nc;
      // Pi~ = (I + mu P+)^-1 Pi+ = Pi+ - mu P~ Pi+ (measured: as accurate here as solving with the factor of I + mu P+)
      for (int i = 0; i < ndx; i++)
        for (int j = 0; j < ndx; j++)
        {
          double acc = 0.0;
          for (int k = 0; k < ndx; k++)
            acc = std::fma(s.Pt(i, k), E.Pi(k, j), acc);
          Pit(i, j) = std::fma(-mudyn, acc, E.Pi(i, j));
        }
      // Y = B^T Pi~
      for (int i = 0; i < nu; i++)
        for (int j = 0; j < ndx; j++)
        {
          double acc = 0.0;
          for (int k = 0; k < ndx; k++)
            acc = std::fma(kn.B(k, i), Pit(k, j), acc);
          Y(i, j) = acc;
        }
      // R^ = R + B^T P~ B = L L^T
      PB.zero();
      add_AB(s.Pt, kn.B, ndx, ndx, nu, PB);
      Rh = kn.R;
      add_AtB(kn.B, PB, ndx, nu, nu, Rh);
      if (!chol_lower(Rh, nu))
        ok = false;
      // X = H Y, H = L^-T (I - W (mu I + W^T W)^-1 W^T) L^-1, W = L^-1 D^T
      X = Y;
      solve_lower(Rh, nu, X, ndx);
      if (nc > 0)
      {
        Mat W(nu, nc), Sg(nc, nc), G(nc, ndx);
        for (int i = 0; i < nu; i++)
          for (int c = 0; c < nc; c++)
            W(i, c) = kn.D(c, i);
        solve_lower(Rh, nu, W, nc);
        for (int i = 0; i < nc; i++)
          for (int j = 0; j < nc; j++)
          {
            double acc = (i == j) ? mueq : 0.0;
            for (int k = 0; k < nu; k++)
              acc = std::fma(W(k, i), W(k, j), acc);
            Sg(i, j) = acc;
          }
        if (!chol_lower(Sg, nc))
          ok = false;
        for (int i = 0; i < nc; i++)
          for (int j = 0; j < ndx; j++)
          {
            double acc = 0.0;
            for (int k = 0; k < nu; k++)
              acc = std::fma(W(k, i), X(k, j), acc);
            G(i, j) = acc;
          }
        solve_lower(Sg, nc, G, ndx);
        solve_lower_T(Sg, nc, G, ndx);
        for (int i = 0; i < nu; i++)
          for (int j = 0; j < ndx; j++)
          {
            double acc = X(i, j);
            for (int c = 0; c < nc; c++)
              acc = std::fma(-W(i, c), G(c, j), acc);
            X(i, j) = acc;
          }
      }
      solve_lower_T(Rh, nu, X, ndx);
      // Lam += Y^T X + mu Pi+^T Pi~
      for (int i = 0; i < ndx; i++)
        for (int j = 0; j < ndx; j++)
        {
          double acc = E.Lam(i, j);
          for (int k = 0; k < nu; k++)
            acc = std::fma(Y(k, i), X(k, j), acc);
          double acc2 = 0.0;
          for (int k = 0; k < ndx; k++)
            acc2 = std::fma(E.Pi(k, i), Pit(k, j), acc2);
          E.Lam(i, j) = std::fma(mudyn, acc2, acc);
        }
      // psi += Pi+^T phi, phi = (I - mu P~)(B k + f) - mu p~
      for (int i = 0; i < ndx; i++)
      {
        double acc = kn.f[i];
        for (int k = 0; k < nu; k++)
          acc = std::fma(kn.B(i, k), s.k[k], acc);
        v[i] = acc;
      }
      for (int i = 0; i < ndx; i++)
      {
        double acc = s.pt[i];
        for (int k = 0; k < ndx; k++)
          acc = std::fma(s.Pt(i, k), v[k], acc);
        phi[i] = std::fma(-mudyn, acc, v[i]);
      }
      for (int i = 0; i < ndx; i++)
      {
        double acc = E.psi[i];
        for (int k = 0; k < ndx; k++)
          acc = std::fma(E.Pi(k, i), phi[k], acc);
        E.psi[i] = acc;
      }
      // Pi = A^T Pi~ + K^T Y
      for (int i = 0; i < ndx; i++)
        for (int j = 0; j < ndx; j++)
        {
          double acc = 0.0;
          for (int k = 0; k < ndx; k++)
            acc = std::fma(kn.A(k, i), Pit(k, j), acc);
          for (int k = 0; k < nu; k++)
            acc = std::fma(s.K(k, i), Y(k, j), acc);
          Pin(i, j) = acc;
        }
      E.Pi = Pin;
    }
    for (int i = 0; i < ndx; i++)
      for (int j = 0; j < i; j++)
      {
        const double m = 0.5 * (E.Lam(i, j) + E.Lam(j, i));
        E.Lam(i, j) = m;
        E.Lam(j, i) = m;
      }
    return ok;
  }
  // pass 3, one boundary: cost-to-go (Pn, pn) at the end of a leg with element E and start values (Pa, pa) -> (P, p) at its start
  inline bool leg_condense(const LegElement & E, const Mat & Pa, const Vec & pa, const Mat & Pn, const Vec & pn, int ndx, Mat & P, Vec & p)
  {
    // Pn (I + Lam Pn)^-1 = C (I + C^T Lam C)^-1 C^T with Pn = C C^T: symmetric positive definite all the way (measured: gains
    // within 1e-9 of the serial sweep for 3 to 10 legs; the LU form below, kept for a cost-to-go that is only semi-definite,
    // loses a digit or two to the 1/mu stiffness of the penalised directions)
    {
      // S = I + C^T Lam C ; T = C S^-1 C^T [Psi | psi - Lam pn] (+ pn)
      Mat Cc = Pn;
      if (chol_lower(Cc, ndx))
      {
        for (int i = 0; i < ndx; i++)
          for (int j = i + 1; j < ndx; j++)
            Cc(i, j) = 0.0;
        Mat LC(ndx, ndx), S(ndx, ndx), R(ndx, ndx + 1), U(ndx, ndx + 1), T(ndx, ndx + 1);
        add_AB(E.Lam, Cc, ndx, ndx, ndx, LC);
        for (int i = 0; i < ndx; i++)
          S(i, i) = 1.0;
        add_AtB(Cc, LC, ndx, ndx, ndx, S);
        for (int i = 0; i < ndx; i++)
          for (int j = 0; j < i; j++)
            S(i, j) = S(j, i) = 0.5 * (S(i, j) + S(j, i));
        chol_lower(S, ndx);
        // (the algebraically equal vector part C S^-1 (C^T psi + C^-1 pn) was measured too: C^-1 pn loses more to the
        // conditioning of C than psi - Lam pn loses to cancellation)
        for (int i = 0; i < ndx; i++)
        {
          for (int j = 0; j < ndx; j++)
            R(i, j) = E.Pi(j, i);
          double acc = E.psi[i];
          for (int k = 0; k < ndx; k++)
            acc = std::fma(-E.Lam(i, k), pn[k], acc);
          R(i, ndx) = acc;
        }
        add_AtB(Cc, R, ndx, ndx, ndx + 1, U);
        solve_lower(S, ndx, U, ndx + 1);
        solve_lower_T(S, ndx, U, ndx + 1);
        add_AB(Cc, U, ndx, ndx, ndx + 1, T);
        for (int i = 0; i < ndx; i++)
          T(i, ndx) += pn[i];
        P.resize(ndx, ndx);
        p.assign(ndx, 0.0);
        for (int i = 0; i < ndx; i++)
          for (int j = 0; j <= ndx; j++)
          {
            double acc = (j == ndx) ? pa[i] : Pa(i, j);
            for (int k = 0; k < ndx; k++)
              acc = std::fma(E.Pi(i, k), T(k, j), acc);
            if (j == ndx)
              p[i] = acc;
            else
              P(i, j) = acc;
          }
        for (int i = 0; i < ndx; i++)
          for (int j = 0; j < i; j++)
          {
            const double m = 0.5 * (P(i, j) + P(j, i));
            P(i, j) = m;
            P(j, i) = m;
          }
        return true;
      }
    }
    // M = I + Lam Pn ; [Xi | xi] = M^-1 [Psi | psi - Lam pn]
    Mat M(ndx, ndx), R(ndx, ndx + 1);
    for (int i = 0; i < ndx; i++)
      for (int j = 0; j < ndx; j++)
      {
        double acc = (i == j) ? 1.0 : 0.0;
        for (int k = 0; k < ndx; k++)
          acc = std::fma(E.Lam(i, k), Pn(k, j), acc);
        M(i, j) = acc;
      }
    for (int i = 0; i < ndx; i++)
    {
      for (int j = 0; j < ndx; j++)
        R(i, j) = E.Pi(j, i);
      double acc = E.psi[i];
      for (int k = 0; k < ndx; k++)
        acc = std::fma(-E.Lam(i, k), pn[k], acc);
      R(i, ndx) = acc;
    }
    if (!lu_solve(M, ndx, R, ndx + 1))
      return false;
    // T = Pn [Xi | xi] (+ pn on the vector column) ; [P | p] = [Pa | pa] + Psi^T T
    Mat T(ndx, ndx + 1);
    for (int i = 0; i < ndx; i++)
      for (int j = 0; j <= ndx; j++)
      {
        double acc = (j == ndx) ? pn[i] : 0.0;
        for (int k = 0; k < ndx; k++)
          acc = std::fma(Pn(i, k), R(k, j), acc);
        T(i, j) = acc;
      }
    P.resize(ndx, ndx);
    p.assign(ndx, 0.0);
    for (int i = 0; i < ndx; i++)
      for (int j = 0; j <= ndx; j++)
      {
        double acc = (j == ndx) ? pa[i] : Pa(i, j);
        for (int k = 0; k < ndx; k++)
          acc = std::fma(E.Pi(i, k), T(k, j), acc);
        if (j == ndx)
          p[i] = acc;
        else
          P(i, j) = acc;
      }
    for (int i = 0; i < ndx; i++)
      for (int j = 0; j < i; j++)
      {
        const double m = 0.5 * (P(i, j) + P(j, i));
        P(i, j) = m;
        P(j, i) = m;
      }
    return true;
  }
  inline int riccati_legs_used(int N, int J) { return (J >= 2 && N >= 2 * J) ? J : 1; }
  inline RiccatiStats riccati_backward_parallel(
    const std::vector<Knot> & knots, int ndx, int nu, double mudyn, double mueq, std::vector<StageSol> & sol, int J)
  {
    const int N = (int)knots.size() - 1;
    if (riccati_legs_used(N, J) < 2)
      return riccati_backward(knots, ndx, nu, mudyn, mueq, sol);
    RiccatiStats st;
    std::vector<int> a;
    leg_bounds(N, J, a);
    sol.resize(N + 1);
    Mat Zero(ndx, ndx);
    Vec zero(ndx, 0.0);
    std::vector<Mat> P0(J), PT(J);
    std::vector<Vec> p0(J), pT(J);
    for (int j = 1; j < J; j++) // pass 1
    {
      const bool last = j == J - 1;
      if (!riccati_backward(knots, ndx, nu, mudyn, mueq, sol, a[j], a[j + 1], last ? nullptr : &Zero, last ? nullptr : &zero).ok)
        st.ok = false;
      P0[j] = sol[a[j]].P;
      p0[j] = sol[a[j]].p;
    }
    std::vector<LegElement> E(J);
    for (int j = 1; j < J - 1; j++) // pass 2
      if (!leg_element(knots, sol, ndx, nu, mudyn, mueq, a[j], a[j + 1], E[j]))
        st.ok = false;
    PT[J - 1] = P0[J - 1]; // pass 3
    pT[J - 1] = p0[J - 1];
    for (int j = J - 2; j >= 1; j--)
      if (!leg_condense(E[j], P0[j], p0[j], PT[j + 1], pT[j + 1], ndx, PT[j], pT[j]))
        st.ok = false;
    for (int j = J - 2; j >= 0; j--) // pass 4
      if (!riccati_backward(knots, ndx, nu, mudyn, mueq, sol, a[j], a[j + 1], &PT[j + 1], &pT[j + 1]).ok)
        st.ok = false;
    return st;
  }

  // Forward sweep: dx0 given (0 with force_initial_condition). Outputs steps dx, du and the TOTAL new
  // multipliers nu+ (per stage, plus terminal) and lambda+ (lam[0] initial condition, lam[t+1] dynamics t).
  inline void riccati_forward(
    const std::vector<Knot> & knots, const std::vector<StageSol> & sol, int ndx, int nu, double mudyn, std::vector<Vec> & dxs,
    std::vector<Vec> & dus, std::vector<Vec> & vnew, std::vector<Vec> & lnew)
  {
    const int N = (int)knots.size() - 1;
    dxs.assign(N + 1, Vec(ndx, 0.0));
    dus.assign(N, Vec(nu, 0.0));
    vnew.resize(N + 1);
    lnew.assign(N + 1, Vec(ndx, 0.0));
    // lambda_0 = P_0 dx_0 + p_0
    for (int a = 0; a < ndx; a++)
    {
      double acc = sol[0].p[a];
      for (int b = 0; b < ndx; b++)
        acc = std::fma(sol[0].P(a, b), dxs[0][b], acc);
      lnew[0][a] = acc;
    }
    for (int t = 0; t < N; t++)
    {
      const Knot & kn = knots[t];
      const StageSol & s = sol[t];
      const Vec & dx = dxs[t];
      for (int a = 0; a < nu; a++)
      {
        double acc = s.k[a];
        for (int b = 0; b < ndx; b++)
          acc = std::fma(s.K(a, b), dx[b], acc);
        dus[t][a] = acc;
      }
      vnew[t].assign(kn.nc, 0.0);
      for (int i = 0; i < kn.nc; i++)
      {
        double acc = s.z[i];
        for (int b = 0; b < ndx; b++)
          acc = std::fma(s.Z(i, b), dx[b], acc);
        vnew[t][i] = acc;
      }
      Vec y(ndx);
      for (int a = 0; a < ndx; a++)
      {
        double acc = kn.f[a];
        for (int b = 0; b < ndx; b++)
          acc = std::fma(kn.A(a, b), dx[b], acc);
        for (int b = 0; b < nu; b++)
          acc = std::fma(kn.B(a, b), dus[t][b], acc);
        y[a] = acc;
      }
      for (int a = 0; a < ndx; a++)
      {
        double acc = s.pt[a];
        for (int b = 0; b < ndx; b++)
          acc = std::fma(s.Pt(a, b), y[b], acc);
        lnew[t + 1][a] = acc;
        dxs[t + 1][a] = y[a] - mudyn * acc;
      }
    }
    const Knot & kn = knots[N];
    vnew[N].assign(kn.nc, 0.0);
    for (int i = 0; i < kn.nc; i++)
    {
      double acc = sol[N].z[i];
      for (int b = 0; b < ndx; b++)
        acc = std::fma(sol[N].Z(i, b), dxs[N][b], acc);
      vnew[N][i] = acc;
    }
  }

  // normal-cone projection of the shifted constraint value z (per row); returns projection, sets active
  inline double ncone_proj(int type, double z, double lo, double hi, bool & active)
  {
    if (type == CSTR_EQ)
    {
      active = true;
      return z;
    }
    if (type == CSTR_NEG)
    {
      active = z > 0;
      return active ? z : 0.0;
    }
    // box
    if (z < lo)
    {
      active = true;
      return z - lo;
    }
    if (z > hi)
    {
      active = true;
      return z - hi;
    }
    active = false;
    return 0.0;
  }

  inline double inf_norm(const Vec & v)
  {
    double m = 0;
    for (double e : v)
      m = std::max(m, std::fabs(e));
    return m;
  }

  struct Solver
  {
    Problem P;
    smpc_solver_opts opt;
    int N, nx, ndx, nu, ncmax;
    // problem data
    std::vector<smpc_slot> slots;
    smpc_terminal term;
    // iterate (persistent between solves: aligator results_ / workspace_)
    std::vector<Vec> xs, us, vs, lams;
    std::vector<Vec> prev_vs, prev_lams;
    double mu, preg, inner_tol, prim_tol;
    double ls_mov_avg = 0, ls_avg_weight = 0;
    // last linearisation / solution
    std::vector<StageEval> ev;
    std::vector<Knot> knots;
    std::vector<StageSol> sol;
    std::vector<Vec> dxs, dus, vnew, lnew;
    std::vector<Vec> xdots; // getStateDerivative
    std::vector<Vec> fcs;   // getContactForces (full dynamics)
    smpc_solve_info info;
    // wall-clock split of run() for the timed CPU baseline: [0] linearisation, [1] Riccati backward, [2] forward sweep,
    // [3] linesearch trials (values only), seconds accumulated since construction
    double timers[4] = {0, 0, 0, 0};
    int riccati_legs = 1; // > 1: PARALLEL (leg-condensed) backward sweep (smpc_solver_opts::num_threads)
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

    Solver(const smpc_model_desc & md, const smpc_ocp_desc & od, const smpc_solver_opts & so) : P(md, od), opt(so)
    {
      riccati_legs = so.num_threads > 1 ? so.num_threads : 1;
      N = P.N;
      nx = P.nx;
      ndx = P.ndx;
      nu = P.nu;
      ncmax = std::max(P.max_nc(), P.term_nc());
      slots.resize(N);
      for (auto & s : slots)
        memset(&s, 0, sizeof(s));
      memset(&term, 0, sizeof(term));
      xs.assign(N + 1, Vec(nx, 0.0));
      us.assign(N, Vec(nu, 0.0));
      vs.assign(N + 1, Vec(ncmax, 0.0));
      lams.assign(N + 1, Vec(ndx, 0.0));
      prev_vs = vs;
      prev_lams = lams;
      xdots.assign(N, Vec(ndx, 0.0));
      fcs.assign(N, Vec());
      mu = opt.mu_init;
      preg = opt.reg_init;
      memset(&info, 0, sizeof(info));
    }

    // replaceStageCircular + cycleProblem: drop stage 0, append a new tail stage; duals follow.
    void cycle(const smpc_slot & tail)
    {
      std::rotate(slots.begin(), slots.begin() + 1, slots.end());
      slots[N - 1] = tail;
      // stage multipliers rotate with their stage; the new tail stage starts from zero multipliers
      std::rotate(vs.begin(), vs.begin() + 1, vs.begin() + N);
      std::fill(vs[N - 1].begin(), vs[N - 1].end(), 0.0);
      // co-states: lams[0] (initial condition) stays, lams[1..N] rotate left (wrap-around)
      std::rotate(lams.begin() + 1, lams.begin() + 2, lams.end());
    }

    // values (and derivatives) of every node at (xs_, us_); fills ev
    // knot_threads > 1: OpenMP team over the knots of ONE instance, what aligator does with num_threads
    // (reference src/mpc.cpp:46-52, examples/go2_kinodynamics.py:101); the sum over knots stays in ascending order
    int knot_threads = 1;
    double evaluate(const std::vector<Vec> & xs_, const std::vector<Vec> & us_, bool derivs, std::vector<StageEval> & e)
    {
      e.resize(N + 1);
      std::string err;
#pragma omp parallel for schedule(dynamic, 1) num_threads(knot_threads) if (knot_threads > 1)
      for (int t = 0; t <= N; t++)
      {
        try
        {
          if (t < N)
            stage_eval(P, slots[t], xs_[t].data(), us_[t].data(), derivs, e[t]);
          else
            terminal_eval(P, term, xs_[N].data(), derivs, e[N]);
        }
        catch (const std::exception & ex)
        {
#pragma omp critical
          err = ex.what();
        }
      }
      if (!err.empty())
        throw std::runtime_error(err);
      double cost = 0;
      for (int t = 0; t <= N; t++)
        cost += e[t].cost;
      return cost;
    }

    // AL merit (LinesearchMode::PRIMAL):  cost + sum mu/2 |lam_e + e/mu|^2 + sum mu/2 |Pi_N(nu_e + c/mu)|^2
    // also returns per-node lam_plus, nu_plus (scaled), slack e, and primal infeasibility.
    struct Mults
    {
      std::vector<Vec> e;       // dynamics slack f(x,u) (-) x+
      std::vector<Vec> lplus;   // lam_e + e/mu           (index t+1)
      std::vector<Vec> vproj;   // Pi_N(c + mu nu_e)      (unscaled)
      std::vector<std::vector<char>> active;
      double merit = 0, prim_infeas = 0;
    };
    void multipliers(const std::vector<Vec> & xs_, const std::vector<StageEval> & e, double cost, Mults & m)
    {
      const double mud = mu, mue = mu;
      m.e.assign(N, Vec(ndx, 0.0));
      m.lplus.assign(N + 1, Vec(ndx, 0.0));
      m.vproj.resize(N + 1);
      m.active.resize(N + 1);
      double pen = 0, pinf = 0;
      for (int t = 0; t < N; t++)
      {
        state_difference(P, xs_[t + 1].data(), e[t].xnext.data(), m.e[t].data());
        pinf = std::max(pinf, inf_norm(m.e[t]));
        double sq = 0;
        for (int a = 0; a < ndx; a++)
        {
          double lp = prev_lams[t + 1][a] + m.e[t][a] / mud;
          m.lplus[t + 1][a] = lp;
          sq += lp * lp;
        }
        pen += 0.5 * mud * sq;
      }
      for (int t = 0; t <= N; t++)
      {
        const int nc = e[t].nc;
        m.vproj[t].assign(nc, 0.0);
        m.active[t].assign(nc, 0);
        double sq = 0;
        for (int i = 0; i < nc; i++)
        {
          double z = e[t].c[i] + mue * prev_vs[t][i];
          bool act;
          double pr = ncone_proj(e[t].ctype[i], z, e[t].lo[i], e[t].hi[i], act);
          m.vproj[t][i] = pr;
          m.active[t][i] = act;
          double vp = pr / mue;
          sq += vp * vp;
          // c - Pi_C(c + mu nu_e) = pr - mu nu_e
          pinf = std::max(pinf, std::fabs(pr - mue * prev_vs[t][i]));
        }
        pen += 0.5 * mue * sq;
      }
      m.merit = cost + pen;
      m.prim_infeas = pinf;
    }

    // LQ subproblem at the current iterate
    void build_lq(const Mults & m)
    {
      knots.resize(N + 1);
      for (int t = 0; t <= N; t++)
      {
        const StageEval & e = ev[t];
        Knot & kn = knots[t];
        const int nc = e.nc;
        kn.nc = nc;
        kn.Q = e.lxx;
        kn.q = e.lx;
        for (int a = 0; a < ndx; a++)
          kn.Q(a, a) += preg;
        kn.C.resize(nc, ndx);
        kn.d.assign(nc, 0.0);
        for (int i = 0; i < nc; i++)
        {
          kn.d[i] = m.vproj[t][i];
          if (m.active[t][i])
            for (int a = 0; a < ndx; a++)
              kn.C(i, a) = e.Cx(i, a);
        }
        if (t == N)
          break;
        kn.S = e.lxu;
        kn.R = e.luu;
        kn.r = e.lu;
        for (int a = 0; a < nu; a++)
          kn.R(a, a) += preg;
        kn.A = e.A;
        kn.B = e.B;
        kn.f.assign(ndx, 0.0);
        for (int a = 0; a < ndx; a++)
          kn.f[a] = m.e[t][a] + mu * prev_lams[t + 1][a];
        kn.D.resize(nc, nu);
        for (int i = 0; i < nc; i++)
          if (m.active[t][i])
            for (int a = 0; a < nu; a++)
              kn.D(i, a) = e.Cu(i, a);
      }
    }

    // Lagrangian gradient norms with multipliers (l, v); optionally the directional derivative along (dxs, dus)
    double lagrangian_grad(const std::vector<Vec> & l, const std::vector<Vec> & v, double * dphi)
    {
      double dinf = 0, dd = 0;
      for (int t = 0; t <= N; t++)
      {
        const StageEval & e = ev[t];
        Vec Lx = e.lx;
        for (int a = 0; a < ndx; a++)
        {
          double acc = Lx[a];
          if (t < N)
            for (int b = 0; b < ndx; b++)
              acc += e.A(b, a) * l[t + 1][b];
          for (int i = 0; i < e.nc; i++)
            acc += e.Cx(i, a) * v[t][i];
          acc -= l[t][a];
          Lx[a] = acc;
        }
        if (t > 0) // dx_0 = 0 is imposed, its stationarity row defines lam_0
          dinf = std::max(dinf, inf_norm(Lx));
        if (dphi)
          for (int a = 0; a < ndx; a++)
            dd += Lx[a] * dxs[t][a];
        if (t < N)
        {
          Vec Lu = e.lu;
          for (int a = 0; a < nu; a++)
          {
            double acc = Lu[a];
            for (int b = 0; b < ndx; b++)
              acc += e.B(b, a) * l[t + 1][b];
            for (int i = 0; i < e.nc; i++)
              acc += e.Cu(i, a) * v[t][i];
            Lu[a] = acc;
          }
          dinf = std::max(dinf, inf_norm(Lu));
          if (dphi)
            for (int a = 0; a < nu; a++)
              dd += Lu[a] * dus[t][a];
        }
      }
      if (dphi)
        *dphi = dd;
      return dinf;
    }

    // One call of SolverProxDDP::run (warm start in xs/us, duals persistent).
    bool run(const double * x0, int max_iters)
    {
      xs[0].assign(x0, x0 + nx); // force_initial_condition_ (reference src/mpc.cpp:53)
      // the nonmonotone reference value belongs to one run: a stale average from the previous tick's
      // (different) problem would reject every step whenever the merit jumps between ticks
      ls_mov_avg = 0;
      ls_avg_weight = 0;
      mu = opt.mu_init;
      prev_vs = vs;
      prev_lams = lams;
      auto tols_on_failure = [&]() {
        prim_tol = opt.prim_tol0 * std::pow(mu, opt.prim_alpha);
        inner_tol = opt.inner_tol0 * std::pow(mu, opt.dual_alpha);
      };
      tols_on_failure();
      inner_tol = std::max(inner_tol, opt.tol);
      prim_tol = std::max(prim_tol, opt.tol);
      info.num_iters = 0;
      info.al_iters = 0;
      info.conv = 0;
      info.ls_trials = 0;
      info.alpha = 0;
      info.flags = 0;
      bool conv = false;
      while (info.al_iters < opt.max_al_iters && info.num_iters < max_iters)
      {
        bool inner_conv = inner_loop(max_iters);
        if (!inner_conv)
        {
          info.al_iters++;
          break;
        }
        if (info.conv) // outer criterion met inside the inner loop
        {
          conv = true;
          break;
        }
        if (info.prim_infeas <= prim_tol)
        {
          prim_tol = std::max(prim_tol * std::pow(mu, opt.prim_beta), opt.tol);
          inner_tol = std::max(inner_tol * std::pow(mu, opt.dual_beta), opt.tol);
          prev_vs = vs; // MultiplierUpdateMode::NEWTON
          prev_lams = lams;
          if (std::max(info.dual_infeas, info.prim_infeas) <= opt.tol)
          {
            conv = true;
            break;
          }
        }
        else
        {
          mu = std::max(mu * opt.mu_update_factor, opt.mu_lower_bound);
          tols_on_failure();
          inner_tol = std::max(inner_tol, opt.tol);
          prim_tol = std::max(prim_tol, opt.tol);
        }
        info.al_iters++;
      }
      info.conv = conv;
      info.mu = mu;
      info.preg = preg;
      return conv;
    }

    // returns true when the inner (fixed mu, fixed multiplier estimates) problem is solved
    bool inner_loop(int max_iters)
    {
      for (; info.num_iters < max_iters;)
      {
        // linearise at the current iterate (values + derivatives), multipliers, merit
        double tm0 = now();
        double cost = evaluate(xs, us, true, ev);
        timers[0] += now() - tm0;
        for (int t = 0; t < N; t++)
        {
          xdots[t] = ev[t].xdot;
          fcs[t] = ev[t].fc;
        }
        Mults m;
        multipliers(xs, ev, cost, m);
        const double phi0 = m.merit;
        info.traj_cost = cost;
        info.merit = phi0;
        build_lq(m);
        info.prim_infeas = m.prim_infeas;
        info.dual_infeas = lagrangian_grad(lams, vs, nullptr);
        // inner criterion: stationarity + dual residuals mu (mult_plus - mult)
        double inner_crit = info.dual_infeas;
        for (int t = 0; t < N; t++)
          for (int a = 0; a < ndx; a++)
            inner_crit = std::max(inner_crit, std::fabs(mu * (m.lplus[t + 1][a] - lams[t + 1][a])));
        for (int t = 0; t <= N; t++)
          for (int i = 0; i < ev[t].nc; i++)
            inner_crit = std::max(inner_crit, std::fabs(m.vproj[t][i] - mu * vs[t][i]));
        if (std::max(info.prim_infeas, info.dual_infeas) <= opt.tol)
        {
          info.conv = 1;
          return true;
        }
        // Newton step of the proximal subproblem
        tm0 = now();
        if (!(riccati_legs > 1 ? riccati_backward_parallel(knots, ndx, nu, mu, mu, sol, riccati_legs) : riccati_backward(knots, ndx, nu, mu, mu, sol)).ok)
          info.flags |= 1;
        timers[1] += now() - tm0;
        tm0 = now();
        riccati_forward(knots, sol, ndx, nu, mu, dxs, dus, vnew, lnew);
        timers[2] += now() - tm0;
        if (inner_crit <= inner_tol)
          return true;
        // directional derivative of the merit: gradient with the first-order multipliers lam_plus, nu_plus
        std::vector<Vec> vplus(N + 1);
        for (int t = 0; t <= N; t++)
        {
          vplus[t].assign(ev[t].nc, 0.0);
          for (int i = 0; i < ev[t].nc; i++)
            vplus[t][i] = m.vproj[t][i] / mu;
        }
        std::vector<Vec> lplus = m.lplus;
        lplus[0].assign(ndx, 0.0);
        double dphi0 = 0;
        lagrangian_grad(lplus, vplus, &dphi0);
        // nonmonotone (Zhang-Hager) backtracking linesearch on the merit, LINEAR rollout
        ls_mov_avg = opt.ls_avg_eta * ls_avg_weight * ls_mov_avg + phi0;
        ls_avg_weight = opt.ls_avg_eta * ls_avg_weight + 1.0;
        ls_mov_avg /= ls_avg_weight;
        double alpha = 1.0;
        std::vector<Vec> txs(N + 1, Vec(nx)), tus(N, Vec(nu));
        std::vector<StageEval> tev;
        Mults tm;
        double tcost = 0;
        int trials = 0;
        auto try_step = [&](double a) {
          for (int t = 0; t <= N; t++)
          {
            Vec sdx(ndx);
            for (int i = 0; i < ndx; i++)
              sdx[i] = a * dxs[t][i];
            state_integrate(P, xs[t].data(), sdx.data(), txs[t].data());
          }
          for (int t = 0; t < N; t++)
            for (int i = 0; i < nu; i++)
              tus[t][i] = us[t][i] + a * dus[t][i];
          tcost = evaluate(txs, tus, false, tev);
          multipliers(txs, tev, tcost, tm);
          trials++;
          return tm.merit;
        };
        double phi_new;
        tm0 = now();
        while (true)
        {
          phi_new = try_step(alpha);
          if (phi_new <= ls_mov_avg + opt.ls_armijo_c1 * alpha * dphi0)
            break;
          if (alpha <= opt.ls_alpha_min || trials >= opt.ls_max_steps)
            break;
          alpha = std::max(alpha * opt.ls_contraction, opt.ls_alpha_min);
        }
        timers[3] += now() - tm0;
        // accept
        for (int t = 0; t <= N; t++)
        {
          xs[t] = txs[t];
          for (int i = 0; i < ev[t].nc; i++)
            vs[t][i] += alpha * (vnew[t][i] - vs[t][i]);
          for (int a = 0; a < ndx; a++)
            lams[t][a] += alpha * (lnew[t][a] - lams[t][a]);
        }
        for (int t = 0; t < N; t++)
        {
          us[t] = tus[t];
          xdots[t] = tev[t].xdot;
          fcs[t] = tev[t].fc;
        }
        info.traj_cost = tcost;
        info.merit = phi_new;
        info.alpha = alpha;
        info.ls_trials = trials;
        if (alpha <= opt.ls_alpha_min)
          preg = std::min(preg * opt.reg_inc_k, opt.reg_max);
        else
          preg = std::max(preg * opt.reg_dec_k, opt.reg_min);
        info.num_iters++;
      }
      return false;
    }
  };

} // namespace orc
