// chol2.cuh — right-looking Cholesky of a small dense matrix held in shared memory, two columns per barrier, trailing
// matrix in registers as 2 x 2 tiles, one warp dedicated to the pivots. Used by riccati3.cuh and riccati4.cuh for the
// three factorisations of a Riccati stage (I + mu P, R^, Sigma). Bitwise equal to the sequential right-looking algorithm
// of oracle/orc_solver.hpp (chol_lower): every entry receives the same fused multiply-adds in the same order.
//
// Why (profiles/r02c_ric3_phase_cycles.txt): the one-column-per-barrier factorisation cost ~700 cycles per column with
// three CTAs per SM (43 % of the Riccati kernel). In the step for columns (j, j+1) every thread rebuilds L(., j) and
// L(., j+1) of its tile's two rows and two columns from the published raw columns (L(i,j+1) = fma(-L(i,j), L(j+1,j),
// m(i,j+1)) / L(j+1,j+1), the very operation the sequential algorithm performs): 4 vector loads + 12 multiplies for 8
// fused multiply-adds instead of 16 loads + 16 multiplies, and tiles whose columns are factorised are skipped. The pivot
// warp owns (d(2u), m(2u+1,2u), d(2u+1)) per lane, so the serial chain of a step (1/sqrt, scale, fma, 1/sqrt) stays
// inside one lane: no shuffle, no extra barrier. Measured on B200 (profiles/r02e): 51.5 k -> 36.9 k cycles per stage.
#pragma once
#include "par.cuh"

namespace smpc
{

  // Reciprocal square root of a Cholesky pivot. The pivot chain is the serial critical path of the sweep (one per column,
  // 84 columns per stage): ncu attributes 17 % of the kernel's stall samples to the barrier the other warps wait at while
  // the pivot lane runs sqrt-then-divide (2 MUFU + 12 dependent FP64 operations, 147 cycles measured,
  // profiles/r02g_microbench_fp64.txt). The device build therefore uses CUDA's rsqrt() (MUFU.RSQ64H + one cubic Newton
  // step, 61 cycles, <= 1 ulp): measured on B200 at B = 4096 26.2 -> 24.4 ms per launch, single-instance tick 2.53 -> 2.33 ms
  // (profiles/r02i). Consequence: on the device the factor differs from the oracle's correctly rounded 1.0 / sqrt(d) in the last
  // bit, i.e. by the same 1-ulp perturbation the two independent knot evaluations already differ by; the GPU parity tolerances are
  // unchanged (tests/test_gpu_parity.py). -DSMPC_EXACT_PIVOTS restores 1.0 / sqrt(d); the CPU emulation harness always uses
  // it, so tests/test_emul_vs_oracle.py keeps checking the sweep's operation order bit for bit.
  SMPC_DEV double pivot_rsqrt(double d)
  {
#if defined(__CUDACC__) && !defined(SMPC_EXACT_PIVOTS)
    return rsqrt(d);
#else
    return 1.0 / sqrt(d);
#endif
  }

  // a * b + c with two roundings, as the oracle computes I + mu P (the compiler must not contract it into one fma)
  SMPC_DEV double mul_then_add(double a, double b, double c)
  {
#ifdef __CUDACC__
    return __dadd_rn(__dmul_rn(a, b), c);
#else
    return a * b + c; // the host harness is built with -ffp-contract=off
#endif
  }
  SMPC_DEV void st2(double * p, double a, double b)
  {
#ifdef __CUDACC__
    *reinterpret_cast<double2 *>(p) = make_double2(a, b);
#else
    p[0] = a;
    p[1] = b;
#endif
  }

  // ---- right-looking Cholesky, two columns per barrier (see the header) ------------------------------------------
  // In: the LOWER triangle (with diagonal) of the N x N matrix M (ld LD). Out: L(i,j), i > j, in M[i][j] AND in the
  // transposed slot M[j][i]; invd[j] = 1 / L(j,j). col: 2 x (2 N + 4) doubles.
  template <int N, int LD, int TPB>
  SMPC_DEV void chol2(double * M, double * col, double * invd, int * bad)
  {
    static_assert(N % 2 == 0 && N >= 2, "chol2 works on column pairs");
    constexpr int TW = TPB - 32;                 // threads that own off-diagonal 2 x 2 tiles
    constexpr int NT = N / 2;                    // tile rows / columns
    constexpr int NTL = NT * (NT - 1) / 2;       // strictly-lower tiles
    constexpr int ET = (NTL + TW - 1) / TW > 0 ? (NTL + TW - 1) / TW : 1; // tiles per thread
    constexpr int SS = (NT + 31) / 32;           // diagonal pairs (d(2u), m(2u+1,2u), d(2u+1)) per pivot lane
    constexpr int CS = 2 * N + 4;                // one published column pair: [i] -> (m(i,j), m(i,j+1)); then 1/L(j,j), 1/L(j+1,j+1), L(j+1,j)
    constexpr int NR = (4 * ET > 3 * SS) ? 4 * ET : 3 * SS;
    SMPC_THREAD_ARRAY(m, NR)
    SMPC_THREAD_IARRAY(ik, ET)
    SMPC_PAR(tid)
    {
      double * m = SMPC_THREAD_PTR(m, NR, tid);
      int * ik = SMPC_THREAD_PTR(ik, ET, tid);
      if (tid < TW)
      {
#pragma unroll
        for (int q = 0; q < ET; q++)
        {
          const int e = tid + TW * q;
          int K = 0, I = 0;
          if (e < NTL)
          {
            // column-major over the strictly-lower tiles: column K holds rows K+1 .. NT-1; a thread's q-th tile sits in a later
            // column than its (q-1)-th one, so every thread keeps live tiles until late in the factorisation
            int rem = e;
            while (rem >= NT - 1 - K)
            {
              rem -= NT - 1 - K;
              K++;
            }
            I = K + 1 + rem;
#pragma unroll
            for (int r = 0; r < 2; r++)
            {
              double a, b;
              ld2(M + (2 * I + r) * LD + 2 * K, a, b);
              m[4 * q + 2 * r] = a;
              m[4 * q + 2 * r + 1] = b;
              if (K == 0)
                st2(col + 2 * (2 * I + r), a, b);
            }
          }
          else
            m[4 * q] = m[4 * q + 1] = m[4 * q + 2] = m[4 * q + 3] = 0.0;
          ik[q] = (e < NTL) ? ((I << 16) | (K + 1)) : 0; // low half: K + 1 (0 = no tile)
        }
      }
      else
      {
        const int lane = tid - TW;
#pragma unroll
        for (int q = 0; q < SS; q++)
        {
          const int u = lane + 32 * q;
          const bool on = u < NT;
          m[3 * q] = on ? M[(2 * u) * LD + 2 * u] : 1.0;
          m[3 * q + 1] = on ? M[(2 * u + 1) * LD + 2 * u] : 0.0;
          m[3 * q + 2] = on ? M[(2 * u + 1) * LD + 2 * u + 1] : 1.0;
        }
        ik[0] = 0; // pivot lanes: "met a non-positive pivot"
        if (lane == 0)
        {
          const double d0 = m[0];
          const double inv0 = pivot_rsqrt(d0);
          const double l10 = m[1] * inv0;
          const double d1 = fma(-l10, l10, m[2]);
          const double inv1 = pivot_rsqrt(d1);
          col[2 * N] = inv0;
          col[2 * N + 1] = inv1;
          col[2 * N + 2] = l10;
          invd[0] = inv0;
          invd[1] = inv1;
          M[LD] = l10; // L(1,0): nobody reads M during the factorisation, the tiles live in registers
          M[1] = l10;
          if (!(d0 > 0) || !(d1 > 0))
            ik[0] = 1;
        }
      }
    }
    SMPC_END
#pragma unroll 1
    for (int s = 0; s < NT - 1; s++) // the last tile column holds only its diagonal pair, finished by step NT - 2
    {
      SMPC_PAR(tid)
      {
        double * m = SMPC_THREAD_PTR(m, NR, tid);
        int * ik = SMPC_THREAD_PTR(ik, ET, tid);
        const double * cur = col + (s & 1) * CS;
        double * nxt = col + ((s + 1) & 1) * CS;
        const double invj = cur[2 * N], invj1 = cur[2 * N + 1], lj1j = cur[2 * N + 2];
        if (tid < TW)
        {
#pragma unroll
          for (int q = 0; q < ET; q++)
          {
            const int K = (ik[q] & 0xffff) - 1, I = ik[q] >> 16;
            if (K > s)
            {
              double li[2], li1[2], lk[2], lk1[2];
#pragma unroll
              for (int r = 0; r < 2; r++)
              {
                double a, b;
                ld2(cur + 2 * (2 * I + r), a, b);
                li[r] = a * invj;
                li1[r] = fma(-li[r], lj1j, b) * invj1;
                ld2(cur + 2 * (2 * K + r), a, b);
                lk[r] = a * invj;
                lk1[r] = fma(-lk[r], lj1j, b) * invj1;
              }
#pragma unroll
              for (int r = 0; r < 2; r++)
#pragma unroll
                for (int c = 0; c < 2; c++)
                {
                  double v = fma(-li[r], lk[c], m[4 * q + 2 * r + c]);
                  m[4 * q + 2 * r + c] = fma(-li1[r], lk1[c], v);
                }
              if (K == s + 1)
              {
                st2(nxt + 2 * (2 * I), m[4 * q], m[4 * q + 1]);
                st2(nxt + 2 * (2 * I + 1), m[4 * q + 2], m[4 * q + 3]);
              }
            }
          }
        }
        else
        {
          const int lane = tid - TW;
#pragma unroll
          for (int q = 0; q < SS; q++)
          {
            const int u = lane + 32 * q;
            if (u > s && u < NT)
            {
              double a, b;
              ld2(cur + 2 * (2 * u), a, b);
              const double l0 = a * invj;
              const double l01 = fma(-l0, lj1j, b) * invj1;
              ld2(cur + 2 * (2 * u + 1), a, b);
              const double l1 = a * invj;
              const double l11 = fma(-l1, lj1j, b) * invj1;
              double d0 = fma(-l0, l0, m[3 * q]);
              d0 = fma(-l01, l01, d0);
              double sb = fma(-l1, l0, m[3 * q + 1]);
              sb = fma(-l11, l01, sb);
              double d1 = fma(-l1, l1, m[3 * q + 2]);
              d1 = fma(-l11, l11, d1);
              m[3 * q] = d0;
              m[3 * q + 1] = sb;
              m[3 * q + 2] = d1;
              if (u == s + 1)
              {
                // the serial chain of the step, inside one lane
                const double inva = pivot_rsqrt(d0);
                const double lba = sb * inva;
                const double d1f = fma(-lba, lba, d1);
                const double invb = pivot_rsqrt(d1f);
                nxt[2 * N] = inva;
                nxt[2 * N + 1] = invb;
                nxt[2 * N + 2] = lba;
                invd[2 * u] = inva;
                invd[2 * u + 1] = invb;
                M[(2 * u + 1) * LD + 2 * u] = lba;
                M[(2 * u) * LD + 2 * u + 1] = lba;
                if (!(d0 > 0) || !(d1f > 0))
                  ik[0] = 1;
              }
            }
          }
        }
      }
      SMPC_END
    }
    SMPC_PAR(tid)
    {
      const double * m = SMPC_THREAD_PTR(m, NR, tid);
      const int * ik = SMPC_THREAD_PTR(ik, ET, tid);
      if (tid < TW)
      {
#pragma unroll
        for (int q = 0; q < ET; q++)
        {
          const int K = (ik[q] & 0xffff) - 1, I = ik[q] >> 16;
          if (K >= 0)
          {
            // the tile's second column still lacks the update by its first one (the panel is never its own trailing matrix)
            const double i0 = invd[2 * K], i1 = invd[2 * K + 1], lsub = M[(2 * K + 1) * LD + 2 * K];
#pragma unroll
            for (int r = 0; r < 2; r++)
            {
              const double a = m[4 * q + 2 * r] * i0;
              const double b = fma(-a, lsub, m[4 * q + 2 * r + 1]) * i1;
              st2(M + (2 * I + r) * LD + 2 * K, a, b);
              M[(2 * K) * LD + 2 * I + r] = a;
              M[(2 * K + 1) * LD + 2 * I + r] = b;
            }
          }
        }
      }
      else
      {
        if (ik[0])
          *bad = 1;
      }
    }
    SMPC_END
  }

} // namespace smpc
