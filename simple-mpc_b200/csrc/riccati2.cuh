// riccati2.cuh — the proximal Riccati backward sweep of riccati.cuh, specialised at compile time on
// (NDX, NU, NC) and restructured for the SM:
//   * factorisations: ONE THREAD PER MATRIX ROW, the row held in registers (Cholesky-Crout). A column step
//     is the two fused dot products  t = M(i,j) - sum_k L(i,k) L(j,k)  (own row x pivot row, the pivot row
//     read from shared memory as a broadcast) and the pivot  s = M(j,j) - sum_k L(j,k)^2  recomputed by
//     every thread, so a column costs one barrier and no exchange;
//   * triangular sweeps: ONE THREAD PER RIGHT-HAND-SIDE COLUMN, the column held in registers, forward and
//     backward substitution back to back with no barrier at all (the factor is read as broadcasts);
//   * reciprocal pivots, no fp64 division inside the sweeps;
//   * compile-time extents: index arithmetic is multiply/shift, the register arrays are fully unrolled;
//   * 1x4 register tiles in the large products (P~[A B], [A B]^T (P~[A B]));
//   * constraint rows padded to NC with exact zeros (the knot kernel writes them), which leaves every
//     partial sum bit-identical to the oracle's nc-row loops.
// About a dozen barriers per stage instead of ~450 in the generic kernel. Per-element operation order is
// still the canonical order of oracle/orc_solver.hpp (ascending-k FMAs; descending for L^-T), so the
// sweep stays bitwise equal to the oracle on identical knots (tests/test_emul_vs_oracle.py).
//
// Shared memory per CTA (doubles): PP [NDX][NDX+1] = [p | P] · M (Cholesky of I + mu P, later Rh and
// Sigma) · AB [NDX][NDX+NU] (later [Y | W]) · PAB (later G) · QS = [Qh | Sh] · small vectors.
#pragma once
#include "par.cuh"
#include "dev_types.h"

namespace smpc
{

  template <int NDX, int NU, int NC>
  struct Ric2
  {
    static constexpr int LDP = NDX + 1;
    static constexpr int NAB = NDX + NU;
    static constexpr int NRY = NDX + 1 + NC;
    static constexpr int oPP = 0;
    static constexpr int oM = oPP + NDX * LDP;
    static constexpr int szM = (NDX * NDX > NU * NU + NC * NC) ? NDX * NDX : NU * NU + NC * NC;
    static constexpr int oAB = oM + szM;
    static constexpr int oPAB = oAB + NDX * NAB;
    static constexpr int oQS = oPAB + NDX * NAB;
    static constexpr int ow = oQS + NDX * NAB;
    static constexpr int oqr = ow + NDX;   // [qh | rh]
    static constexpr int oinvp = oqr + NAB; // reciprocal pivots of I + mu P
    static constexpr int oinvr = oinvp + NDX; // ... of Rh
    static constexpr int oinvs = oinvr + NU;  // ... of Sigma
    static constexpr int ofv = oinvs + NC;    // f of the stage
    static constexpr int oflag = (ofv + NDX + 1) & ~1;
    static_assert((oM % 2) == 0 && (szM % 2) == 0 && (NAB % 2) == 0, "128-bit shared loads need even offsets");
    static constexpr int SIZE = oflag + 2;
    static_assert(NU * NRY <= NDX * NAB, "[Y|W] must fit in the AB buffer");
    static_assert(NC * LDP <= NDX * NAB, "G must fit in the PAB buffer");
  };

  // two consecutive doubles from a 16-byte aligned shared address in one LDS.128
  SMPC_DEV void ld2(const double * p, double & a, double & b)
  {
#ifdef __CUDACC__
    const double2 v = *reinterpret_cast<const double2 *>(p);
    a = v.x;
    b = v.y;
#else
    a = p[0];
    b = p[1];
#endif
  }

  // Cholesky of the N x N matrix held in the LOWER triangle of M (ld LD): on exit L(i,j), i > j, replaces
  // M(i,j); the diagonal keeps the un-rooted pivots and invd[j] = 1 / L(j,j). Thread i owns row i.
  template <int N, int LD>
  SMPC_DEV void chol_rows(double * M, double * invd, int * bad)
  {
    SMPC_THREAD_ARRAY(row, N)
    SMPC_PAR(tid)
    {
      double * row = SMPC_THREAD_PTR(row, N, tid);
      if (tid < N)
      {
#pragma unroll
        for (int k = 0; k < N; k++)
          row[k] = (k <= tid) ? M[tid * LD + k] : 0.0;
      }
    }
    SMPC_END
#pragma unroll
    for (int j = 0; j < N; j++)
    {
      SMPC_PAR(tid)
      {
        double * row = SMPC_THREAD_PTR(row, N, tid);
        if (tid >= j && tid < N)
        {
          double s = M[j * LD + j];
          double t = row[j];
#pragma unroll
          for (int k = 0; k < j; k++)
          {
            const double ljk = M[j * LD + k];
            s = fma(-ljk, ljk, s);
            t = fma(-row[k], ljk, t);
          }
          const double l = sqrt(s);
          const double inv = 1.0 / l;
          if (tid == j)
          {
            invd[j] = inv;
            if (!(s > 0))
              *bad = 1;
          }
          else
          {
            const double lij = t * inv;
            row[j] = lij;
            M[tid * LD + j] = lij;
          }
        }
      }
      SMPC_END
    }
  }

  // X <- L^-1 X (FWD), X <- L^-T X (BWD) or both, for the N x MC block X (ld LDX); L from chol_rows.
  // Thread c owns column c. No barrier inside; the caller's phase boundary follows.
  template <int N, int LD, int MC, int LDX, bool FWD, bool BWD>
  SMPC_DEV void solve_cols(const double * __restrict__ M, const double * __restrict__ invd, double * __restrict__ X)
  {
    static_assert(MC <= 128, "one thread per column: MC must not exceed the CTA size");
    SMPC_PAR(tid)
    const int c = tid; // (a strided loop over c keeps x[] out of registers)
    if (c < MC)
    {
      double x[N];
#pragma unroll
      for (int i = 0; i < N; i++)
        x[i] = X[i * LDX + c];
      if (FWD)
      {
#pragma unroll
        for (int k = 0; k < N; k++)
        {
          x[k] = x[k] * invd[k];
#pragma unroll
          for (int i = k + 1; i < N; i++)
            x[i] = fma(-M[i * LD + k], x[k], x[i]);
          SMPC_SCHED_FENCE // keeps ptxas from hoisting all N^2/2 factor loads to the top and spilling
        }
      }
      if (BWD)
      {
#pragma unroll
        for (int k = N - 1; k >= 0; k--)
        {
          x[k] = x[k] * invd[k];
#pragma unroll
          for (int i = 0; i < k; i++)
            x[i] = fma(-M[k * LD + i], x[k], x[i]);
          SMPC_SCHED_FENCE
        }
      }
#pragma unroll
      for (int i = 0; i < N; i++)
        X[i * LDX + c] = x[i];
    }
    SMPC_END
  }

  template <int NDX, int NU, int NC>
  SMPC_DEV void riccati2_backward_inst(
    const KnotLayout & L, const SolLayout & SL, int N, const double * __restrict__ knots, double * __restrict__ sol,
    double mudyn, double mueq, double * sm, int * flag_out, double * P0out)
  {
    typedef Ric2<NDX, NU, NC> R;
    constexpr int LDP = R::LDP, NAB = R::NAB, NRY = R::NRY;
    double * PP = sm + R::oPP; // [NDX][LDP]: col 0 = p, cols 1.. = P
    double * Mb = sm + R::oM;
    double * AB = sm + R::oAB;
    double * PAB = sm + R::oPAB;
    double * QS = sm + R::oQS;
    double * w = sm + R::ow;
    double * qr = sm + R::oqr;
    double * invp = sm + R::oinvp;
    double * invr = sm + R::oinvr;
    double * invs = sm + R::oinvs;
    double * fv = sm + R::ofv;
    int * bad = (int *)(sm + R::oflag);
    double * Rh = Mb;           // [NU][NU] after step 1
    double * Sg = Mb + NU * NU; // [NC][NC]
    double * YW = AB;           // [NU][NRY] after step 3
    double * G = PAB;           // [NC][LDP] after step 3
    const double inv_mu = 1.0 / mueq;

    // ---- terminal node: Z = C/mu, z = d/mu ; P = Q + C^T Z ; p = q + C^T z ----
    {
      const double * kn = knots + (size_t)N * L.size;
      double * so = sol + (size_t)N * SL.size;
      SMPC_PAR(tid)
      {
        if (tid == 0)
          *bad = 0;
        for (int e = tid; e < NC * NDX; e += SMPC_NT)
          so[SL.oZ + e] = kn[L.oC + e] * inv_mu;
        for (int i = tid; i < NC; i += SMPC_NT)
          so[SL.oz + i] = kn[L.od + i] * inv_mu;
      }
      SMPC_END
      SMPC_PAR(tid)
      for (int e = tid; e < NDX * LDP; e += SMPC_NT)
      {
        const int a = e / LDP, bb = e % LDP;
        double acc;
        if (bb == 0)
        {
          acc = kn[L.oq + a];
          for (int i = 0; i < NC; i++)
            acc = fma(kn[L.oC + i * NDX + a], so[SL.oz + i], acc);
        }
        else
        {
          acc = kn[L.oQ + a * NDX + (bb - 1)];
          for (int i = 0; i < NC; i++)
            acc = fma(kn[L.oC + i * NDX + a], so[SL.oZ + i * NDX + (bb - 1)], acc);
        }
        PP[e] = acc;
      }
      SMPC_END
    }

    for (int t = N - 1; t >= 0; t--)
    {
      const double * kn = knots + (size_t)t * L.size;
      double * so = sol + (size_t)t * SL.size;
      // ---- 1. [p~ | P~] = (I + mudyn P)^-1 [p | P], in place ----
      if (mudyn > 0)
      {
        SMPC_PAR(tid)
        {
#ifdef __CUDACC__
          // pull this stage's knot (A B Q S R q r f C D d) towards L2 while the factorisation below runs;
          // the "stage data in" phase then finds it there instead of paying HBM latency
          for (int ln = tid; ln * 16 < L.size; ln += SMPC_NT)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(kn + ln * 16));
#endif
          for (int e = tid; e < NDX * NDX; e += SMPC_NT)
          {
            const int a = e / NDX, b = e % NDX;
            if (b <= a)
              Mb[e] = mudyn * PP[a * LDP + 1 + b] + (a == b ? 1.0 : 0.0);
          }
        }
        SMPC_END
        chol_rows<NDX, NDX>(Mb, invp, bad);
        // two phases: one 2 x N^2/2-FMA basic block makes ptxas give up on register allocation
        solve_cols<NDX, NDX, LDP, LDP, true, false>(Mb, invp, PP);
        solve_cols<NDX, NDX, LDP, LDP, false, true>(Mb, invp, PP);
      }
      // ---- stage data in; P~ p~ out (for the forward sweep) ----
      // every HBM operand of steps 2-3 is fetched here, all loads independent (deep memory-level
      // parallelism), so no FMA chain later waits on a global load: [A B] -> AB, [Q S] -> QS (accumulated
      // in place), R -> Rh, [q r] -> qr, f -> fv
      SMPC_PAR(tid)
      {
        for (int e = tid; e < NDX * NAB; e += SMPC_NT)
        {
          const int a = e / NAB, b = e % NAB;
          AB[e] = (b < NDX) ? kn[L.oA + a * NDX + b] : kn[L.oB + a * NU + (b - NDX)];
          QS[e] = (b < NDX) ? kn[L.oQ + a * NDX + b] : kn[L.oS + a * NU + (b - NDX)];
        }
        for (int e = tid; e < NU * NU; e += SMPC_NT)
          Rh[e] = kn[L.oR + e];
        for (int a = tid; a < NAB; a += SMPC_NT)
          qr[a] = (a < NDX) ? kn[L.oq + a] : kn[L.or_ + (a - NDX)];
        for (int a = tid; a < NDX; a += SMPC_NT)
          fv[a] = kn[L.of + a];
        for (int e = tid; e < NDX * LDP; e += SMPC_NT)
        {
          const int a = e / LDP, bb = e % LDP;
          if (bb == 0)
            so[SL.opt + a] = PP[e];
          else
            so[SL.oPt + a * NDX + (bb - 1)] = PP[e];
        }
      }
      SMPC_END
      // ---- 2. w = P~ f + p~ ; PAB = P~ [A B]  (4x6 register tiles, 128-bit shared loads) ----
      SMPC_PAR(tid)
      {
        constexpr int TR = NDX / 4, TC = NAB / 6;
        static_assert(NDX % 4 == 0 && NAB % 6 == 0, "tile shape 4x6 must divide NDX x (NDX+NU)");
        static_assert(TR * TC + NDX <= 128, "one tile per thread plus the w rows");
        if (tid < TR * TC)
        {
          const int a0 = 4 * (tid / TC), b0 = 6 * (tid % TC);
          double c[4][6];
#pragma unroll
          for (int r = 0; r < 4; r++)
#pragma unroll
            for (int q = 0; q < 6; q++)
              c[r][q] = 0.0;
#pragma unroll 4
          for (int k = 0; k < NDX; k++)
          {
            double bv[6];
            ld2(AB + k * NAB + b0, bv[0], bv[1]);
            ld2(AB + k * NAB + b0 + 2, bv[2], bv[3]);
            ld2(AB + k * NAB + b0 + 4, bv[4], bv[5]);
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
              const double pk = PP[(a0 + r) * LDP + 1 + k];
#pragma unroll
              for (int q = 0; q < 6; q++)
                c[r][q] = fma(pk, bv[q], c[r][q]);
            }
          }
#pragma unroll
          for (int r = 0; r < 4; r++)
#pragma unroll
            for (int q = 0; q < 6; q++)
              PAB[(a0 + r) * NAB + b0 + q] = c[r][q];
        }
        else if (tid - TR * TC < NDX)
        {
          const int a = tid - TR * TC;
          double acc = PP[a * LDP];
          for (int b = 0; b < NDX; b++)
            acc = fma(PP[a * LDP + 1 + b], fv[b], acc);
          w[a] = acc;
        }
      }
      SMPC_END
      // ---- 3. [Qh | Sh] = [Q | S] + A^T PAB ; Rh = R + B^T PB ; qh = q + A^T w ; rh = r + B^T w ----
      // all three accumulate in place on the values staged above (same order as the oracle: the cost block
      // first, then ascending-k FMAs); [A B]^T PAB is one 60 x 60 product of which only [Qh | Sh] and Rh are kept
      SMPC_PAR(tid)
      {
        constexpr int TR = NDX / 4, TC = NAB / 6, TRU = NU / 4, TCU = NU / 6;
        static_assert(NU % 4 == 0 && NU % 6 == 0, "tile shape 4x6 must divide NU x NU");
        static_assert(TR * TC + TRU * TCU <= 128, "one tile per thread");
        if (tid < TR * TC + TRU * TCU)
        {
          const bool isR = tid >= TR * TC;
          const int tt = isR ? tid - TR * TC : tid;
          // output rows a0.. of [A B]^T (columns of AB), output columns b0.. of PAB
          const int a0 = isR ? NDX + 4 * (tt / TCU) : 4 * (tt / TC);
          const int b0 = isR ? NDX + 6 * (tt % TCU) : 6 * (tt % TC);
          double * out = isR ? Rh + (a0 - NDX) * NU + (b0 - NDX) : QS + a0 * NAB + b0;
          const int ldo = isR ? NU : NAB;
          double c[4][6];
#pragma unroll
          for (int r = 0; r < 4; r++)
#pragma unroll
            for (int q = 0; q < 6; q++)
              c[r][q] = out[r * ldo + q];
#pragma unroll 4
          for (int k = 0; k < NDX; k++)
          {
            double av[4], bv[6];
            ld2(AB + k * NAB + a0, av[0], av[1]);
            ld2(AB + k * NAB + a0 + 2, av[2], av[3]);
            ld2(PAB + k * NAB + b0, bv[0], bv[1]);
            ld2(PAB + k * NAB + b0 + 2, bv[2], bv[3]);
            ld2(PAB + k * NAB + b0 + 4, bv[4], bv[5]);
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
              for (int q = 0; q < 6; q++)
                c[r][q] = fma(av[r], bv[q], c[r][q]);
          }
#pragma unroll
          for (int r = 0; r < 4; r++)
#pragma unroll
            for (int q = 0; q < 6; q++)
              out[r * ldo + q] = c[r][q];
        }
        if (tid < NAB)
        {
          double acc = qr[tid];
          for (int k = 0; k < NDX; k++)
            acc = fma(AB[k * NAB + tid], w[k], acc);
          qr[tid] = acc;
        }
      }
      SMPC_END
      // ---- 4. KKT by block LDL^T ----
      chol_rows<NU, NU>(Rh, invr, bad);
      SMPC_PAR(tid)
      for (int e = tid; e < NU * NRY; e += SMPC_NT)
      {
        const int a = e / NRY, c = e % NRY;
        double v;
        if (c == 0)
          v = qr[NDX + a];
        else if (c <= NDX)
          v = QS[(c - 1) * NAB + NDX + a];
        else
          v = kn[L.oD + (c - NDX - 1) * NU + a];
        YW[e] = v;
      }
      SMPC_END
      solve_cols<NU, NU, NRY, NRY, true, false>(Rh, invr, YW);
      if (NC > 0)
      {
        SMPC_PAR(tid)
        {
          for (int e = tid; e < NC * NC; e += SMPC_NT)
          {
            const int i = e / NC, j = e % NC;
            if (j <= i)
            {
              double acc = (i == j) ? mueq : 0.0;
              for (int k = 0; k < NU; k++)
                acc = fma(YW[k * NRY + LDP + i], YW[k * NRY + LDP + j], acc);
              Sg[e] = acc;
            }
          }
          for (int e = tid; e < NC * LDP; e += SMPC_NT)
          {
            const int i = e / LDP, c = e % LDP;
            double acc = (c == 0) ? kn[L.od + i] : kn[L.oC + i * NDX + (c - 1)];
            for (int k = 0; k < NU; k++)
              acc = fma(-YW[k * NRY + LDP + i], YW[k * NRY + c], acc);
            G[e] = acc;
          }
        }
        SMPC_END
        chol_rows<NC, NC>(Sg, invs, bad);
        solve_cols<NC, NC, LDP, LDP, true, false>(Sg, invs, G);
        solve_cols<NC, NC, LDP, LDP, false, true>(Sg, invs, G);
        SMPC_PAR(tid)
        for (int e = tid; e < NU * LDP; e += SMPC_NT)
        {
          const int a = e / LDP, c = e % LDP;
          double acc = YW[a * NRY + c];
          for (int i = 0; i < NC; i++)
            acc = fma(YW[a * NRY + LDP + i], G[i * LDP + c], acc);
          YW[a * NRY + c] = acc;
        }
        SMPC_END
      }
      solve_cols<NU, NU, LDP, NRY, false, true>(Rh, invr, YW);
      // gains: [k | K] = -Y, [z | Z] = G
      SMPC_PAR(tid)
      {
        for (int e = tid; e < NU * LDP; e += SMPC_NT)
        {
          const int a = e / LDP, c = e % LDP;
          const double val = -YW[a * NRY + c];
          YW[a * NRY + c] = val;
          if (c == 0)
            so[SL.ok + a] = val;
          else
            so[SL.oK + a * NDX + (c - 1)] = val;
        }
        for (int e = tid; e < NC * LDP; e += SMPC_NT)
        {
          const int i = e / LDP, c = e % LDP;
          if (c == 0)
            so[SL.oz + i] = G[e];
          else
            so[SL.oZ + i * NDX + (c - 1)] = G[e];
        }
      }
      SMPC_END
      // ---- 5. [p | P] = [qh | Qh] + Sh [k | K] + C^T [z | Z] ----
      SMPC_PAR(tid)
      for (int e = tid; e < NDX * LDP; e += SMPC_NT)
      {
        const int a = e / LDP, bb = e % LDP;
        double acc = (bb == 0) ? qr[a] : QS[a * NAB + (bb - 1)];
#pragma unroll 4
        for (int k = 0; k < NU; k++)
          acc = fma(QS[a * NAB + NDX + k], YW[k * NRY + bb], acc);
#pragma unroll 4
        for (int i = 0; i < NC; i++)
          acc = fma(kn[L.oC + i * NDX + a], G[i * LDP + bb], acc);
        PP[e] = acc;
      }
      SMPC_END
      // symmetrise P (two phases: read both halves, then write)
      SMPC_PAR(tid)
      for (int e = tid; e < NDX * NDX; e += SMPC_NT)
      {
        const int a = e / NDX, b = e % NDX;
        if (b < a)
          Mb[e] = 0.5 * (PP[a * LDP + 1 + b] + PP[b * LDP + 1 + a]);
      }
      SMPC_END
      SMPC_PAR(tid)
      for (int e = tid; e < NDX * NDX; e += SMPC_NT)
      {
        const int a = e / NDX, b = e % NDX;
        if (b < a)
        {
          PP[a * LDP + 1 + b] = Mb[e];
          PP[b * LDP + 1 + a] = Mb[e];
        }
      }
      SMPC_END
    }
    SMPC_PAR(tid)
    {
      if (tid == 0 && flag_out)
        *flag_out = *bad;
      if (P0out) // layout expected by riccati_forward_inst: [P | p]
        for (int e = tid; e < NDX * LDP; e += SMPC_NT)
        {
          const int a = e / LDP, bb = e % LDP;
          P0out[a * LDP + (bb == 0 ? NDX : bb - 1)] = PP[e];
        }
    }
    SMPC_END
  }

} // namespace smpc
