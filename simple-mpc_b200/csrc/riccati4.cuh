// riccati4.cuh — proximal Riccati backward sweep, any even (NDX, NU, NC): the shapes of every multibody configuration
// (Go2 kinodynamics 36/24/24|36, Go2 full dynamics 36/12/24|60, Talos full dynamics 56/22/78|90, ...).
//
// Same recursion and the same per-element operation order as oracle/orc_solver.hpp (see riccati.cuh for the algebra and
// riccati3.cuh for the staging scheme), so the sweep stays BITWISE equal to the oracle on identical knots
// (tests/test_emul_vs_oracle.py). What is new against riccati3.cuh, each item answering profiles/r02c_ric3_phase_cycles.txt
// (three Cholesky factorisations = 43 % of the kernel at ~700 cycles per column with three CTAs per SM):
//   * chol2: TWO columns per barrier. The strictly-lower part is dealt as 2 x 2 register tiles; in the step for columns
//     (j, j+1) every thread rebuilds L(., j) and L(., j+1) of its tile's two rows and two columns from the published raw
//     columns (L(i,j+1) = fma(-L(i,j), L(j+1,j), m(i,j+1)) / L(j+1,j+1), the very operation the sequential algorithm performs),
//     i.e. 4 vector loads + 12 multiplies for 8 fused multiply-adds instead of 16 loads + 16 multiplies; tiles whose
//     columns are factorised are skipped. The pivot warp owns (d(2u), m(2u+1,2u), d(2u+1)) per lane, so the whole
//     serial chain of a step (1/sqrt, scale, fma, 1/sqrt) stays inside one lane: no shuffle, no extra barrier.
//   * shapes are template parameters but nothing assumes "one tile per thread": every product is a job list of 4 x 4
//     register tiles strided over the CTA; ragged edges (NU = 22, NC = 78) read past the block into the next block of the
//     same buffer (finite or not, the values are discarded) and predicate the stores.
//   * the knot arrives in three bulk copies ([Q S R q r], [A B f], later [C D d]) to three freely placed regions, so the
//     dead [A B f | P~] span can host Sigma for large NC (Talos: 78 x 78).
//   * M = I + mu P for the next stage is written by the symmetrisation phase (one phase and one pass over P less).
#pragma once
#include "riccati3.cuh"

namespace smpc
{

  constexpr int r4_even(int x) { return (x + 1) & ~1; }
  constexpr int r4_max(int a, int b) { return a > b ? a : b; }
  constexpr int r4_min(int a, int b) { return a < b ? a : b; }

  template <int NDX, int NU, int NC>
  struct Ric4
  {
    static_assert(NDX % 4 == 0 && NU % 2 == 0 && NC % 2 == 0 && NC > 0, "riccati4: NDX multiple of 4, NU and NC even");
    static constexpr int LDP = NDX + 1, NAB = NDX + NU, NRY = NDX + 1 + NC;
    // dense knot layout: must equal make_knot_layout(NDX, NU, NC) (checked by the host at create)
    static constexpr int kQ = 0;
    static constexpr int kS = kQ + NDX * NDX;
    static constexpr int kR = kS + NDX * NU;
    static constexpr int kq = kR + NU * NU;
    static constexpr int kr = kq + NDX;
    static constexpr int kA = kr + NU;
    static constexpr int kB = kA + NDX * NDX;
    static constexpr int kf = kB + NDX * NU;
    static constexpr int kC = kf + NDX;
    static constexpr int kD = kC + NC * NDX;
    static constexpr int kd = kD + NC * NU;
    static constexpr int KSIZE = (kd + NC + 1) & ~1;
    static_assert(kA % 2 == 0 && kC % 2 == 0, "bulk copies need 16-byte aligned pieces");
    static constexpr int szH = kA, szABf = kC - kA, szCDd = KSIZE - kC;
    static constexpr int szPP = r4_even(NDX * LDP);
    static constexpr int szX = r4_even(r4_max(NDX * NDX, NU * NRY)) + 8; // M | PAB rows [0, PR0) | YW ; + read slack
    static constexpr int PR0 = r4_min(NDX, 4 * ((r4_even(r4_max(NDX * NDX, NU * NRY)) / NAB) / 4)); // PAB rows kept in X
    static_assert((NDX - PR0) * NAB <= szCDd, "the tail rows of PAB must fit the C D d region");
    static constexpr int szG = r4_even(NC * LDP), szSg = r4_even(NC * NC);
    // where G = [z | Z] and Sigma live (see header): inside dead regions when they fit
    static constexpr bool BOTH_IN_ABF = szG + szSg <= szABf;
    static constexpr bool SG_IN_PP = !BOTH_IN_ABF && szG <= szABf && szSg <= szPP;
    static constexpr bool SG_IN_U = !BOTH_IN_ABF && !SG_IN_PP && szSg <= szABf + szPP; // spans [A B f | P~]; G on its own
    static constexpr bool G_OWN = !(BOTH_IN_ABF || SG_IN_PP);
    static constexpr bool SG_OWN = !(BOTH_IN_ABF || SG_IN_PP || SG_IN_U);
    // regions, in this order (A B f and P~ adjacent for SG_IN_U)
    static constexpr int oH = 0;
    static constexpr int oCDd = oH + szH;
    static constexpr int oABf = oCDd + szCDd;
    static constexpr int oPP = oABf + szABf;
    static constexpr int oX = oPP + szPP;
    static constexpr int oGown = oX + szX;
    static constexpr int oSgown = oGown + (G_OWN ? szG : 0);
    static constexpr int ovec = oSgown + (SG_OWN ? szSg : 0);
    static constexpr int oG = G_OWN ? oGown : oABf;
    static constexpr int oSg = BOTH_IN_ABF ? (oABf + szG) : (SG_IN_PP ? oPP : (SG_IN_U ? oABf : oSgown));
    static constexpr int oPAB1 = oCDd; // PAB rows [PR0, NDX)
    static constexpr int ow = ovec;
    static constexpr int oinvp = ow + NDX;         // pivot reciprocals of I + mu P ...
    static constexpr int oinvr = oinvp;            // ... share storage with those of R^ and Sigma
    static constexpr int oinvs = oinvr + NU;
    static constexpr int szinv = r4_max(NU + NC, NDX);
    static constexpr int NMAX = r4_max(NDX, r4_max(NU, NC));
    static constexpr int ocol = r4_even(oinvr + szinv); // published column pairs, double buffered: 2 x (2 NMAX + 4)
    static constexpr int obar = r4_even(ocol + 2 * (2 * NMAX + 4));
    static constexpr int oflag = obar + 2;
    static constexpr int SIZE = oflag + 2;
  };

  // one thread: arm the barrier for `n1 + n2` doubles, then two bulk copies (n2 may be 0)
  SMPC_DEV void bulk_load2(double * d1, const double * s1, int n1, double * d2, const double * s2, int n2, void * bar)
  {
#ifdef __CUDACC__
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    const unsigned bytes1 = (unsigned)n1 * 8u, bytes2 = (unsigned)n2 * 8u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes1 + bytes2) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(d1)),
                 "l"(s1), "r"(bytes1), "r"(b)
                 : "memory");
    if (n2 > 0)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(d2)),
                   "l"(s2), "r"(bytes2), "r"(b)
                   : "memory");
#else
    for (int i = 0; i < n1; i++)
      d1[i] = s1[i];
    for (int i = 0; i < n2; i++)
      d2[i] = s2[i];
    (void)bar;
#endif
  }
  // X <- L^-1 X (FWD) / X <- L^-T X (BWD), one thread per column with the column in registers (riccati3.cuh sweep_cols),
  // for any column count up to the CTA size
  template <int N, int LD, int MC, int LDX, bool FWD, bool BWD>
  SMPC_DEV void sweep_cols4(const double * __restrict__ M, const double * __restrict__ invd, double * __restrict__ X)
  {
    static_assert(LD % 2 == 0, "factor rows must keep 16-byte alignment");
    SMPC_PAR(tid)
    for (int c = tid; c < MC; c += SMPC_NT)
    {
      double x[N];
#pragma unroll
      for (int i = 0; i < N; i++)
        x[i] = X[i * LDX + c];
      if (FWD)
        SweepFwd<N, LD, 0>::run(M, invd, x);
      if (BWD)
        SweepBwd<N, LD, N - 1>::run(M, invd, x);
#pragma unroll
      for (int i = 0; i < N; i++)
        X[i * LDX + c] = x[i];
    }
    SMPC_END
  }

  // acc[4][4] (+)= sum_k lhs(k, r) * rhs(k, q): lhs(k, r) = lp[k * ldl + r], rhs(k, q) = rp[k * ldr + q], both read as pairs
  // (16-byte aligned: even offsets, even strides). Reads of rows / columns past a ragged edge stay inside the buffers.
  SMPC_DEV void mm44_tn(const double * __restrict__ lp, int ldl, const double * __restrict__ rp, int ldr, int k0, int k1, double (&c)[4][4])
  {
#pragma unroll 4
    for (int k = k0; k < k1; k++)
    {
      double av[4], bv[4];
      ld2(lp + k * ldl, av[0], av[1]);
      ld2(lp + k * ldl + 2, av[2], av[3]);
      ld2(rp + k * ldr, bv[0], bv[1]);
      ld2(rp + k * ldr + 2, bv[2], bv[3]);
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int q = 0; q < 4; q++)
          c[r][q] = fma(av[r], bv[q], c[r][q]);
    }
  }

  // ---- padded mode (GN < NDX): the problem's state dimension GN is not a multiple of 4 (the centroidal problem: 9) ----
  // The kernel computes on NDX = GN rounded up to 4 with the extra rows / columns of every matrix held at exact zeros in
  // shared memory (P, A, Q: zero; I + mu P: identity there), so every real entry receives the operations of the unpadded
  // recursion plus products with exact zeros: bitwise the same values. Global memory keeps the unpadded layouts (L, SL built
  // with ndx = GN); the knot is expanded into the padded regions by ordinary loads instead of bulk copies (a stage is ~10 KB).
  template <int GN, int NDX, int NU, int NC>
  SMPC_DEV void ric4_expand_H_ABf(const KnotLayout & L, const double * __restrict__ kn, double * H, double * ABf, int tid, int nt)
  {
    typedef Ric4<NDX, NU, NC> R;
    constexpr int NAB = NDX + NU;
    for (int e = tid; e < NDX * NDX; e += nt)
    {
      const int a = e / NDX, b = e - a * NDX;
      const bool in = a < GN && b < GN;
      H[R::kQ + e] = in ? kn[L.oQ + a * GN + b] : 0.0;
      ABf[e] = in ? kn[L.oA + a * GN + b] : 0.0;
    }
    for (int e = tid; e < NDX * NU; e += nt)
    {
      const int a = e / NU;
      H[R::kS + e] = a < GN ? kn[L.oS + e] : 0.0;
      ABf[NDX * NDX + e] = a < GN ? kn[L.oB + e] : 0.0;
    }
    for (int e = tid; e < NU * NU; e += nt)
      H[R::kR + e] = kn[L.oR + e];
    for (int a = tid; a < NDX; a += nt)
    {
      H[R::kq + a] = a < GN ? kn[L.oq + a] : 0.0;
      ABf[NDX * NAB + a] = a < GN ? kn[L.of + a] : 0.0;
    }
    for (int j = tid; j < NU; j += nt)
      H[R::kr + j] = kn[L.or_ + j];
  }
  template <int GN, int NDX, int NU, int NC>
  SMPC_DEV void ric4_expand_CDd(const KnotLayout & L, const double * __restrict__ kn, double * CDd, int tid, int nt)
  {
    constexpr int NAB = NDX + NU;
    for (int e = tid; e < NC * NDX; e += nt)
    {
      const int i = e / NDX, b = e - i * NDX;
      CDd[e] = b < GN ? kn[L.oC + i * GN + b] : 0.0;
    }
    for (int e = tid; e < NC * NU; e += nt)
      CDd[NC * NDX + e] = kn[L.oD + e];
    for (int i = tid; i < NC; i += nt)
      CDd[NC * NAB + i] = kn[L.od + i];
  }

  template <int NDX, int NU, int NC, int TPB, int GN = NDX>
  SMPC_DEV void riccati4_backward_inst(
    const KnotLayout & L, const SolLayout & SL, int N, const double * __restrict__ knots, double * __restrict__ sol,
    double mudyn, double mueq, double * sm, int * flag_out, double * P0out, const double * Pterm = nullptr)
  {
    // Pterm != nullptr (a leg of the PARALLEL sweep, riccati_par.cuh): the cost-to-go [P | p] at node N is given (layout of
    // P0out) and node N's own knot is not touched - it belongs to the next leg
    typedef Ric4<NDX, NU, NC> R;
    constexpr int LDP = R::LDP, NAB = R::NAB, NRY = R::NRY, PR0 = R::PR0;
    double * H = sm + R::oH;
    double * ABf = sm + R::oABf;
    double * CDd = sm + R::oCDd;
    double * PP = sm + R::oPP; // [NDX][LDP]: col 0 = p, cols 1.. = P
    double * Mb = sm + R::oX;
    double * PAB = sm + R::oX;     // rows [0, PR0)
    double * PAB1 = sm + R::oPAB1; // rows [PR0, NDX)
    double * YW = sm + R::oX;      // Mb, PAB and YW are the same buffer at different times
    double * G = sm + R::oG;
    double * Sg = sm + R::oSg;
    double * w = sm + R::ow;
    double * invp = sm + R::oinvp;
    double * invr = sm + R::oinvr;
    double * invs = sm + R::oinvs;
    double * colb = sm + R::ocol;
    void * bar = (void *)(sm + R::obar);
    int * bad = (int *)(sm + R::oflag);
    double * Qh = H + R::kQ; // accumulate in place
    double * Sh = H + R::kS;
    double * Rh = H + R::kR;
    double * qh = H + R::kq;
    double * rh = H + R::kr;
    const double * Ak = ABf;
    const double * Bk = ABf + NDX * NDX;
    const double * fk = ABf + NDX * NAB;
    const double * Ck = CDd;
    const double * Dk = CDd + NC * NDX;
    const double * dk = CDd + NC * NAB;
    const double inv_mu = 1.0 / mueq;
    unsigned parity = 0;
    static_assert(GN <= NDX && NDX - GN < 4, "GN is the problem's state dimension, NDX the same rounded up to a multiple of 4");
    constexpr bool PAD = GN != NDX;

    // ---- terminal node: Z = C/mu, z = d/mu ; P = Q + C^T Z ; p = q + C^T z ; M = I + mudyn P ----
    {
      const double * kn = knots + (size_t)N * L.size;
      double * so = sol + (size_t)N * SL.size;
      if (Pterm)
      {
        SMPC_PAR(tid)
        {
          if (tid == 0)
          {
            *bad = 0;
            bulk_init(bar);
          }
          for (int e = tid; e < NDX * LDP; e += SMPC_NT)
          {
            const int a = e / LDP, bb = e % LDP;
            PP[e] = (!PAD || (a < GN && bb <= GN)) ? Pterm[a * (GN + 1) + (bb == 0 ? GN : bb - 1)] : 0.0;
          }
        }
        SMPC_END
      }
      else
      {
        SMPC_PAR(tid)
        {
          if (tid == 0)
          {
            *bad = 0;
            bulk_init(bar);
          }
          for (int e = tid; e < NC * GN; e += SMPC_NT)
            so[SL.oZ + e] = kn[L.oC + e] * inv_mu;
          for (int i = tid; i < NC; i += SMPC_NT)
            so[SL.oz + i] = kn[L.od + i] * inv_mu;
        }
        SMPC_END
        SMPC_PAR(tid)
        for (int e = tid; e < NDX * LDP; e += SMPC_NT)
        {
          const int a = e / LDP, bb = e % LDP;
          double acc = 0.0;
          if (PAD && (a >= GN || bb > GN))
            acc = 0.0; // padding rows / columns
          else if (bb == 0)
          {
            acc = kn[L.oq + a];
            for (int i = 0; i < NC; i++)
              acc = fma(kn[L.oC + i * GN + a], so[SL.oz + i], acc);
          }
          else
          {
            acc = kn[L.oQ + a * GN + (bb - 1)];
            for (int i = 0; i < NC; i++)
              acc = fma(kn[L.oC + i * GN + a], so[SL.oZ + i * GN + (bb - 1)], acc);
          }
          PP[e] = acc;
        }
        SMPC_END
      }
      SMPC_PAR(tid)
      if (mudyn > 0)
        for (int e = tid; e < NDX * NDX; e += SMPC_NT)
        {
          const int a = e / NDX, b = e % NDX;
          if (b <= a)
            Mb[e] = mul_then_add(mudyn, PP[a * LDP + 1 + b], (a == b ? 1.0 : 0.0));
        }
      SMPC_END
    }

    RIC_T0
    for (int t = N - 1; t >= 0; t--)
    {
      const double * kn = knots + (size_t)t * L.size;
      double * so = sol + (size_t)t * SL.size;
      // ---- 0. start the copies of this stage's [Q S R q r] and [A B f]; C D d follow after step 3 ----
      SMPC_PAR(tid)
      if (PAD)
        ric4_expand_H_ABf<GN, NDX, NU, NC>(L, kn, H, ABf, tid, SMPC_NT);
      else if (tid == 0)
        bulk_load2(H, kn, R::szH, ABf, kn + R::kA, R::szABf, bar);
      SMPC_END
      RIC_T(0)
      // ---- 1. [p~ | P~] = (I + mudyn P)^-1 [p | P], in place ----
      if (mudyn > 0)
      {
        chol2<NDX, NDX, TPB>(Mb, colb, invp, bad);
        RIC_T(1)
        sweep_cols4<NDX, NDX, LDP, LDP, true, false>(Mb, invp, PP);
        sweep_cols4<NDX, NDX, LDP, LDP, false, true>(Mb, invp, PP);
        RIC_T(2)
      }
      // ---- 2. knot landed ; P~ p~ out (for the forward sweep) ; w = P~ f + p~ ; PAB = P~ [A B] ----
      SMPC_PAR(tid)
      {
        if (!PAD)
          bulk_wait(bar, parity);
        for (int a = tid; a < GN; a += SMPC_NT)
          so[SL.opt + a] = PP[a * LDP];
        for (int e = tid; e < GN * GN; e += SMPC_NT)
        {
          const int a = e / GN, b = e - a * GN;
          so[SL.oPt + e] = PP[a * LDP + 1 + b];
        }
        // 4 x 4 tiles of PAB: rows a0.., columns over A (NDX / 4 tiles) then over B (ceil(NU / 4) tiles, the last may be 2 wide)
        constexpr int TR = NDX / 4, TCA = NDX / 4, TCB = (NU + 3) / 4, TC = TCA + TCB, NJ = TR * TC;
        for (int job = tid; job < NJ + NDX; job += SMPC_NT)
        {
          if (job < NJ)
          {
            const int a0 = 4 * (job / TC), tc = job % TC;
            const bool inA = tc < TCA;
            const int c0 = inA ? 4 * tc : 4 * (tc - TCA);
            const double * src = inA ? Ak + c0 : Bk + c0;
            const int lds = inA ? NDX : NU;
            const int ncols = inA ? 4 : ((NU - c0) < 4 ? (NU - c0) : 4);
            double c[4][4];
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
              for (int q = 0; q < 4; q++)
                c[r][q] = 0.0;
#pragma unroll 4
            for (int k = 0; k < NDX; k++)
            {
              double bv[4];
              ld2(src + k * lds, bv[0], bv[1]);
              ld2(src + k * lds + 2, bv[2], bv[3]);
#pragma unroll
              for (int r = 0; r < 4; r++)
              {
                const double pk = PP[(a0 + r) * LDP + 1 + k];
#pragma unroll
                for (int q = 0; q < 4; q++)
                  c[r][q] = fma(pk, bv[q], c[r][q]);
              }
            }
            const int cb = inA ? c0 : NDX + c0;
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
              double * dst = ((a0 + r) < PR0) ? PAB + (a0 + r) * NAB : PAB1 + (a0 + r - PR0) * NAB;
              st2(dst + cb, c[r][0], c[r][1]);
              if (ncols == 4)
                st2(dst + cb + 2, c[r][2], c[r][3]);
            }
          }
          else
          {
            const int a = job - NJ;
            double acc = PP[a * LDP];
            for (int b = 0; b < NDX; b++)
              acc = fma(PP[a * LDP + 1 + b], fk[b], acc);
            w[a] = acc;
          }
        }
      }
      SMPC_END
      parity ^= 1u;
      RIC_T(3)
      // ---- 3. [Qh | Sh] = [Q | S] + A^T PAB ; Rh = R + B^T PB ; qh = q + A^T w ; rh = r + B^T w (in place) ----
      SMPC_PAR(tid)
      {
        constexpr int TRA = NDX / 4, TCQ = NDX / 4, TCU = (NU + 3) / 4, TRU = (NU + 3) / 4;
        constexpr int NJQ = TRA * TCQ, NJS = TRA * TCU, NJR = TRU * TCU, NJ = NJQ + NJS + NJR;
        for (int job = tid; job < NJ + NAB; job += SMPC_NT)
        {
          if (job < NJ)
          {
            // left operand: columns ar.. of A (Qh, Sh rows) or of B (Rh rows); right operand: columns b0.. of PAB
            const double * lhs;
            int ldl, b0, ldo, nrows, ncols;
            double * out;
            if (job < NJQ)
            {
              const int ar = 4 * (job / TCQ), cq = 4 * (job % TCQ);
              lhs = Ak + ar, ldl = NDX, b0 = cq, out = Qh + ar * NDX + cq, ldo = NDX, nrows = 4, ncols = 4;
            }
            else if (job < NJQ + NJS)
            {
              const int jj = job - NJQ;
              const int ar = 4 * (jj / TCU), cu = 4 * (jj % TCU);
              lhs = Ak + ar, ldl = NDX, b0 = NDX + cu, out = Sh + ar * NU + cu, ldo = NU, nrows = 4, ncols = (NU - cu) < 4 ? (NU - cu) : 4;
            }
            else
            {
              const int jj = job - NJQ - NJS;
              const int ar = 4 * (jj / TCU), cu = 4 * (jj % TCU);
              lhs = Bk + ar, ldl = NU, b0 = NDX + cu, out = Rh + ar * NU + cu, ldo = NU;
              nrows = (NU - ar) < 4 ? (NU - ar) : 4, ncols = (NU - cu) < 4 ? (NU - cu) : 4;
            }
            double c[4][4];
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
              // rows / columns past a ragged edge start from whatever follows in the buffer: never stored
              ld2(out + r * ldo, c[r][0], c[r][1]);
              ld2(out + r * ldo + 2, c[r][2], c[r][3]);
            }
            // rows [0, PR0) of PAB live in X, the rest in the C D d region: same k order, two loops
            mm44_tn(lhs, ldl, PAB + b0, NAB, 0, PR0, c);
            mm44_tn(lhs, ldl, PAB1 - PR0 * NAB + b0, NAB, PR0, NDX, c);
#pragma unroll
            for (int r = 0; r < 4; r++)
              if (r < nrows)
              {
                st2(out + r * ldo, c[r][0], c[r][1]);
                if (ncols == 4)
                  st2(out + r * ldo + 2, c[r][2], c[r][3]);
              }
          }
          else
          {
            const int tv = job - NJ;
            const bool isq = tv < NDX;
            const double * colp = isq ? Ak + tv : Bk + (tv - NDX);
            const int ld = isq ? NDX : NU;
            double * dst = isq ? qh + tv : rh + (tv - NDX);
            double acc = *dst;
            for (int k = 0; k < NDX; k++)
              acc = fma(colp[k * ld], w[k], acc);
            *dst = acc;
          }
        }
      }
      SMPC_END
      RIC_T(4)
      // ---- 4. KKT by block LDL^T ----
      // PAB is dead: fetch C D d into its tail rows' storage while R^ is factorised
      SMPC_PAR(tid)
      if (PAD)
        ric4_expand_CDd<GN, NDX, NU, NC>(L, kn, CDd, tid, SMPC_NT);
      else if (tid == 0)
        bulk_load2(CDd, kn + R::kC, R::szCDd, nullptr, nullptr, 0, bar);
      SMPC_END
      chol2<NU, NU, TPB>(Rh, colb, invr, bad);
      RIC_T(5)
      SMPC_PAR(tid)
      {
        if (!PAD)
          bulk_wait(bar, parity);
        for (int e = tid; e < NU * NRY; e += SMPC_NT)
        {
          const int a = e / NRY, c = e - a * NRY;
          double v;
          if (c == 0)
            v = rh[a];
          else if (c <= NDX)
            v = Sh[(c - 1) * NU + a];
          else
            v = Dk[(c - NDX - 1) * NU + a];
          YW[e] = v;
        }
      }
      SMPC_END
      parity ^= 1u;
      sweep_cols4<NU, NU, NRY, NRY, true, false>(Rh, invr, YW);
      RIC_T(6)
      {
        // Sigma = mu I + W^T W (lower) and G = [d | C] - W^T Y : 4 x 4 register tiles (ragged last block row when NC % 4 = 2)
        SMPC_PAR(tid)
        {
          constexpr int NB = (NC + 3) / 4, NCG = NDX / 4, GT = NB * NCG, GL = NB, ST = NB * (NB + 1) / 2, NJOBS = GT + GL + ST;
          for (int job = tid; job < NJOBS; job += SMPC_NT)
          {
            if (job < GT)
            {
              // columns 1 .. NDX of G hold Z (from C); tile columns c0 .. c0+3 of C
              const int i0 = 4 * (job / NCG), c0 = 4 * (job % NCG);
              double acc[4][4];
#pragma unroll
              for (int r = 0; r < 4; r++)
              {
                const int ir = (i0 + r < NC) ? i0 + r : NC - 1;
                ld2(Ck + ir * NDX + c0, acc[r][0], acc[r][1]);
                ld2(Ck + ir * NDX + c0 + 2, acc[r][2], acc[r][3]);
              }
#pragma unroll 4
              for (int k = 0; k < NU; k++)
              {
                double wv[4], yv[4];
#pragma unroll
                for (int r = 0; r < 4; r++)
                  wv[r] = YW[k * NRY + LDP + i0 + r];
#pragma unroll
                for (int q = 0; q < 4; q++)
                  yv[q] = YW[k * NRY + 1 + c0 + q];
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                  for (int q = 0; q < 4; q++)
                    acc[r][q] = fma(-wv[r], yv[q], acc[r][q]);
              }
#pragma unroll
              for (int r = 0; r < 4; r++)
                if (i0 + r < NC)
#pragma unroll
                  for (int q = 0; q < 4; q++)
                    G[(i0 + r) * LDP + 1 + c0 + q] = acc[r][q];
            }
            else if (job < GT + GL)
            {
              // column 0 of G holds z (from d)
              const int i0 = 4 * (job - GT);
              double acc[4];
#pragma unroll
              for (int r = 0; r < 4; r++)
                acc[r] = dk[(i0 + r < NC) ? i0 + r : NC - 1];
#pragma unroll 4
              for (int k = 0; k < NU; k++)
              {
                const double yv = YW[k * NRY];
#pragma unroll
                for (int r = 0; r < 4; r++)
                  acc[r] = fma(-YW[k * NRY + LDP + i0 + r], yv, acc[r]);
              }
#pragma unroll
              for (int r = 0; r < 4; r++)
                if (i0 + r < NC)
                  G[(i0 + r) * LDP] = acc[r];
            }
            else
            {
              int tt = job - GT - GL, bi = 0;
              while (tt > bi)
              {
                tt -= bi + 1;
                bi++;
              }
              const int i0 = 4 * bi, j0 = 4 * tt;
              double acc[4][4];
#pragma unroll
              for (int r = 0; r < 4; r++)
#pragma unroll
                for (int q = 0; q < 4; q++)
                  acc[r][q] = (i0 + r == j0 + q) ? mueq : 0.0;
#pragma unroll 4
              for (int k = 0; k < NU; k++)
              {
                double wi[4], wj[4];
#pragma unroll
                for (int r = 0; r < 4; r++)
                {
                  wi[r] = YW[k * NRY + LDP + i0 + r];
                  wj[r] = YW[k * NRY + LDP + j0 + r];
                }
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                  for (int q = 0; q < 4; q++)
                    acc[r][q] = fma(wi[r], wj[q], acc[r][q]);
              }
#pragma unroll
              for (int r = 0; r < 4; r++)
#pragma unroll
                for (int q = 0; q < 4; q++)
                  if (j0 + q <= i0 + r && i0 + r < NC)
                    Sg[(i0 + r) * NC + j0 + q] = acc[r][q];
            }
          }
        }
        SMPC_END
        RIC_T(7)
        chol2<NC, NC, TPB>(Sg, colb, invs, bad);
        RIC_T(8)
        sweep_cols4<NC, NC, LDP, LDP, true, false>(Sg, invs, G);
        sweep_cols4<NC, NC, LDP, LDP, false, true>(Sg, invs, G);
        RIC_T(9)
        // Y <- Y + W G : 4 x 4 register tiles over columns 1 .. NDX, column 0 by four-row strips
        SMPC_PAR(tid)
        {
          constexpr int NBU = (NU + 3) / 4, NCG = NDX / 4, YT = NBU * NCG, NJOBS = YT + NBU;
          for (int job = tid; job < NJOBS; job += SMPC_NT)
          {
            if (job < YT)
            {
              const int a0 = 4 * (job / NCG), c0 = 4 * (job % NCG);
              int ar[4];
#pragma unroll
              for (int r = 0; r < 4; r++)
                ar[r] = (a0 + r < NU) ? a0 + r : NU - 1;
              double acc[4][4];
#pragma unroll
              for (int r = 0; r < 4; r++)
#pragma unroll
                for (int q = 0; q < 4; q++)
                  acc[r][q] = YW[ar[r] * NRY + 1 + c0 + q];
#pragma unroll 4
              for (int i = 0; i < NC; i++)
              {
                double wv[4], gv[4];
#pragma unroll
                for (int r = 0; r < 4; r++)
                  wv[r] = YW[ar[r] * NRY + LDP + i];
#pragma unroll
                for (int q = 0; q < 4; q++)
                  gv[q] = G[i * LDP + 1 + c0 + q];
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                  for (int q = 0; q < 4; q++)
                    acc[r][q] = fma(wv[r], gv[q], acc[r][q]);
              }
#pragma unroll
              for (int r = 0; r < 4; r++)
                if (a0 + r < NU)
#pragma unroll
                  for (int q = 0; q < 4; q++)
                    YW[(a0 + r) * NRY + 1 + c0 + q] = acc[r][q];
            }
            else
            {
              const int a0 = 4 * (job - YT);
              int ar[4];
#pragma unroll
              for (int r = 0; r < 4; r++)
                ar[r] = (a0 + r < NU) ? a0 + r : NU - 1;
              double acc[4];
#pragma unroll
              for (int r = 0; r < 4; r++)
                acc[r] = YW[ar[r] * NRY];
#pragma unroll 4
              for (int i = 0; i < NC; i++)
              {
                const double gv = G[i * LDP];
#pragma unroll
                for (int r = 0; r < 4; r++)
                  acc[r] = fma(YW[ar[r] * NRY + LDP + i], gv, acc[r]);
              }
#pragma unroll
              for (int r = 0; r < 4; r++)
                if (a0 + r < NU)
                  YW[(a0 + r) * NRY] = acc[r];
            }
          }
        }
        SMPC_END
        RIC_T(10)
      }
      sweep_cols4<NU, NU, LDP, NRY, false, true>(Rh, invr, YW);
      RIC_T(11)
      // ---- gains: [k | K] = -Y, [z | Z] = G ----
      SMPC_PAR(tid)
      {
        for (int a = tid; a < NU; a += SMPC_NT)
        {
          const double val = -YW[a * NRY];
          YW[a * NRY] = val;
          so[SL.ok + a] = val;
        }
        for (int e = tid; e < NU * NDX; e += SMPC_NT)
        {
          const int a = e / NDX, c = e - a * NDX;
          const double val = -YW[a * NRY + 1 + c];
          YW[a * NRY + 1 + c] = val;
          if (!PAD || c < GN)
            so[SL.oK + a * GN + c] = val;
        }
        for (int i = tid; i < NC; i += SMPC_NT)
          so[SL.oz + i] = G[i * LDP];
        for (int e = tid; e < NC * NDX; e += SMPC_NT)
        {
          const int i = e / NDX, c = e - i * NDX;
          if (!PAD || c < GN)
            so[SL.oZ + i * GN + c] = G[i * LDP + 1 + c];
        }
      }
      SMPC_END
      RIC_T(12)
      // ---- 5. [p | P] = [qh | Qh] + Sh [k | K] + C^T [z | Z] : 4 x 4 register tiles, p by two-row strips ----
      SMPC_PAR(tid)
      {
        constexpr int TR = NDX / 4, TC = NDX / 4, NJ = TR * TC, NP = NDX / 2;
        for (int job = tid; job < NJ + NP; job += SMPC_NT)
        {
          if (job < NJ)
          {
            const int a0 = 4 * (job / TC), b0 = 4 * (job % TC);
            double c[4][4];
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
              ld2(Qh + (a0 + r) * NDX + b0, c[r][0], c[r][1]);
              ld2(Qh + (a0 + r) * NDX + b0 + 2, c[r][2], c[r][3]);
            }
#pragma unroll 4
            for (int k = 0; k < NU; k++)
            {
              double kv[4];
#pragma unroll
              for (int q = 0; q < 4; q++)
                kv[q] = YW[k * NRY + 1 + b0 + q];
#pragma unroll
              for (int r = 0; r < 4; r++)
              {
                const double sv = Sh[(a0 + r) * NU + k];
#pragma unroll
                for (int q = 0; q < 4; q++)
                  c[r][q] = fma(sv, kv[q], c[r][q]);
              }
            }
#pragma unroll 4
            for (int i = 0; i < NC; i++)
            {
              double cv[4], zv[4];
              ld2(Ck + i * NDX + a0, cv[0], cv[1]);
              ld2(Ck + i * NDX + a0 + 2, cv[2], cv[3]);
#pragma unroll
              for (int q = 0; q < 4; q++)
                zv[q] = G[i * LDP + 1 + b0 + q];
#pragma unroll
              for (int r = 0; r < 4; r++)
#pragma unroll
                for (int q = 0; q < 4; q++)
                  c[r][q] = fma(cv[r], zv[q], c[r][q]);
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
              for (int q = 0; q < 4; q++)
                PP[(a0 + r) * LDP + 1 + b0 + q] = c[r][q];
          }
          else
          {
            // p: two rows per job, independent chains
            const int a1 = 2 * (job - NJ);
            double pv[2] = {qh[a1], qh[a1 + 1]};
            for (int k = 0; k < NU; k++)
            {
              const double kv = YW[k * NRY];
              pv[0] = fma(Sh[a1 * NU + k], kv, pv[0]);
              pv[1] = fma(Sh[(a1 + 1) * NU + k], kv, pv[1]);
            }
            for (int i = 0; i < NC; i++)
            {
              const double zv = G[i * LDP];
              pv[0] = fma(Ck[i * NDX + a1], zv, pv[0]);
              pv[1] = fma(Ck[i * NDX + a1 + 1], zv, pv[1]);
            }
            PP[a1 * LDP] = pv[0];
            PP[(a1 + 1) * LDP] = pv[1];
          }
        }
      }
      SMPC_END
      RIC_T(13)
      // symmetrise P; the same pass leaves M = I + mudyn P (lower triangle) for the next stage in X (YW is dead)
      SMPC_PAR(tid)
      {
        // one thread per entry of the lower triangle (rows r and NDX-1-r share a strip of NDX + 1 entries): each thread reads and
        // rewrites its own pair P(a,b), P(b,a) only, so one phase suffices
        static_assert(NDX % 2 == 0, "row pairing");
        for (int e = tid; e < (NDX / 2) * (NDX + 1); e += SMPC_NT)
        {
          const int r = e / (NDX + 1), c = e - r * (NDX + 1);
          const int a = (c <= r) ? r : NDX - 1 - r, b = (c <= r) ? c : c - r - 1;
          if (b < a)
          {
            const double v = 0.5 * (PP[a * LDP + 1 + b] + PP[b * LDP + 1 + a]);
            Mb[a * NDX + b] = mudyn * v;
            PP[a * LDP + 1 + b] = v;
            PP[b * LDP + 1 + a] = v;
          }
          else
            Mb[a * NDX + a] = mul_then_add(mudyn, PP[a * LDP + 1 + a], 1.0);
        }
      }
      SMPC_END
      RIC_T(14)
    }
    SMPC_PAR(tid)
    {
      if (tid == 0 && flag_out)
        *flag_out = *bad;
      if (P0out) // layout expected by riccati_forward_inst: [P | p], (GN + 1) columns
        for (int e = tid; e < NDX * LDP; e += SMPC_NT)
        {
          const int a = e / LDP, bb = e % LDP;
          if (!PAD || (a < GN && bb <= GN))
            P0out[a * (GN + 1) + (bb == 0 ? GN : bb - 1)] = PP[e];
        }
    }
    SMPC_END
  }

} // namespace smpc
