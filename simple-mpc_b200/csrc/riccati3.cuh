// riccati3.cuh — proximal Riccati backward sweep, compile-time shapes, TMA-staged knots.
//
// Same recursion and same per-element operation order as oracle/orc_solver.hpp (see riccati.cuh for the
// algebra), so it stays bitwise equal to the oracle on identical knots. What it does beyond the generic sweep of riccati.cuh,
// each item answering a line of profiles/r01a_ncu_summary.md:
//   * stage data arrive by ONE bulk asynchronous copy (TMA, cp.async.bulk + mbarrier) of the whole dense
//     knot (Q S R q r A B f C D d are contiguous in HBM) into shared memory, issued at the top of the stage
//     and waited for after the factorisation of I + mu P, which does not need it. Nothing in the stage reads
//     global memory inside a k-loop any more (C, D, d used to be read from HBM in the P update);
//   * Q^, S^, R^, q^, r^ accumulate IN PLACE in the staged knot (no second copy of [Q S]);
//   * factorisations are right-looking over ALL threads with the trailing matrix in registers, TWO columns per
//     barrier (chol2.cuh), the pivot chain inside one lane of a dedicated warp, and the column loop a runtime loop
//     (a fully unrolled factorisation made the kernel 600 KB of SASS and the sweep instruction-fetch bound);
//     the scaled column is also written to the transposed (upper) slot, so the forward sweep reads L by rows
//     (128-bit shared loads) exactly like the backward sweep;
//   * every product that used strided one-output-per-thread loops (Sigma, G, Y + W G, P update) is register
//     tiled or processes four outputs at a time, so no thread sits in one dependent FMA chain.
//
// Shared memory per CTA (doubles): KN knot copy · PP [NDX][NDX+1] = [p | P] · X = {M (Cholesky of I + mu P) | PAB =
// P~[A B] | YW [NU][NDX+1+NC] = [Y | W]} time-shared · G [NC][NDX+1] and Sigma inside the dead A / B blocks of the
// knot copy · small vectors · the mbarrier: 76.6 KB for (36, 24, 24), three CTAs per SM.
#pragma once
#include "par.cuh"
#include "dev_types.h"
#include "chol2.cuh"

// -DSMPC_RIC_PROFILE: thread 0 of every CTA adds the clock64() span of each phase of a stage to g_ric_prof[phase]
// (tools/ric_phases.py prints the split); never defined in the shipped build.
#if defined(__CUDACC__) && defined(SMPC_RIC_PROFILE)
__device__ unsigned long long g_ric_prof[16];
#define RIC_T(k)                                                                                                       \
  if (threadIdx.x == 0)                                                                                                \
  {                                                                                                                    \
    const long long now_ = clock64();                                                                                  \
    atomicAdd(&g_ric_prof[k], (unsigned long long)(now_ - ric_t0_));                                                   \
    ric_t0_ = now_;                                                                                                    \
  }
#define RIC_T0 long long ric_t0_ = clock64();
#else
#define RIC_T(k)
#define RIC_T0
#endif

namespace smpc
{

  template <int NDX, int NU, int NC>
  struct Ric3
  {
    static constexpr int LDP = NDX + 1;
    static constexpr int NAB = NDX + NU;
    static constexpr int NRY = NDX + 1 + NC;
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
    // shared memory. Buffers whose lifetimes inside a stage do not overlap share storage, so that THREE CTAs fit one
    // SM (profiles/r01e: the sweep is latency-bound on its serial pivot chains and scales with resident CTAs):
    //   X  : M = chol(I + mu P) in steps 0-1 | rows [0, PR0) of PAB = P~[A B] in steps 2-3 | YW = [Y | W] in steps
    //        4-5 | symmetrise scratch
    //   the C D d block of the knot copy: rows [PR0, NDX) of PAB in steps 2-3; C D d arrive by a SECOND bulk copy
    //        issued after step 3 and awaited after the factorisation of R^ (which hides it)
    //   G  : in the A block of the knot copy (A, B, f are dead after step 3) when it fits, else its own buffer
    //   Sg : Sigma, in the B block of the knot copy when G took the A block, else in the A block
    static constexpr int oKN = 0;
    static constexpr int oPP = oKN + KSIZE;
    static constexpr int oX = (oPP + NDX * LDP + 1) & ~1;
    static constexpr int PR0 = 4 * ((2 * NDX / 3) / 4); // PAB rows kept in X
    static constexpr int szYW = NU * NRY;
    static constexpr int szX1 = (NDX * NDX > szYW) ? NDX * NDX : szYW;
    static constexpr int szX0 = (PR0 * NAB > szX1) ? PR0 * NAB : szX1;
    static constexpr int szX = (szX0 + 1) & ~1;
    static constexpr int oM = oX, oPAB = oX, oYW = oX;
    static constexpr int oPAB1 = oKN + kC; // PAB rows [PR0, NDX)
    static_assert((NDX - PR0) * NAB <= KSIZE - kC, "the tail rows of PAB must fit the C D d block");
    static_assert(PR0 % 4 == 0 && (NAB % 2) == 0, "tiles never straddle the PAB split; rows stay 16-byte aligned");
    static constexpr bool G_IN_A = NC * LDP <= NDX * NDX;
    static_assert(G_IN_A ? (NC * NC <= NDX * NU) : (NC * NC <= NDX * NDX), "Sigma must fit the dead A / B block of the knot");
    static constexpr int oGown = oX + szX;
    static constexpr int oG = G_IN_A ? (oKN + kA) : oGown;
    static constexpr int oSg = G_IN_A ? (oKN + kB) : (oKN + kA);
    static constexpr int ow = G_IN_A ? oGown : ((oGown + NC * LDP + 1) & ~1);
    // the pivot reciprocals of I + mu P (steps 0-1) share storage with those of R^ and Sigma (steps 4-5)
    static constexpr int oinvr = ow + NDX;
    static constexpr int oinvs = oinvr + NU;
    static constexpr int oinvp = oinvr;
    static constexpr int szinv = (NU + NC > NDX) ? NU + NC : NDX;
    static constexpr int ocol = (oinvr + szinv + 1) & ~1; // published column pairs of chol2, double buffered
    static constexpr int NMAX = (NDX > NC) ? (NDX > NU ? NDX : NU) : (NC > NU ? NC : NU);
    static constexpr int obar = (ocol + 2 * (2 * NMAX + 4) + 1) & ~1; // 8-byte mbarrier
    static constexpr int oflag = obar + 2;
    static constexpr int SIZE = oflag + 2;
    static_assert((kA % 2) == 0 && (kB % 2) == 0 && (kC % 2) == 0 && (kQ % 2) == 0 && (kS % 2) == 0 && (kR % 2) == 0,
                  "128-bit shared loads need even block offsets");
    static_assert(NDX % 4 == 0 && NDX % 6 == 0 && NU % 4 == 0 && NU % 6 == 0, "4x6 tiles must divide the blocks");
  };

  // ---- bulk asynchronous global -> shared copy (TMA) -------------------------------------------------
  // Device: cp.async.bulk completing on an mbarrier; CPU emulation harness: memcpy.
  SMPC_DEV void bulk_init(void * bar)
  {
#ifdef __CUDACC__
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
    (void)bar;
#endif
  }
  // one thread: order earlier generic-proxy accesses of `dst` before the async-proxy write, arm, copy
  SMPC_DEV void bulk_load(double * dst, const double * src, int ndoubles, void * bar)
  {
#ifdef __CUDACC__
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    const unsigned bytes = (unsigned)ndoubles * 8u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src),
                 "r"(bytes), "r"(b)
                 : "memory");
#else
    for (int i = 0; i < ndoubles; i++)
      dst[i] = src[i];
    (void)bar;
#endif
  }
  SMPC_DEV void bulk_wait(void * bar, unsigned parity)
  {
#ifdef __CUDACC__
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(b),
      "r"(parity)
      : "memory");
#else
    (void)bar;
    (void)parity;
#endif
  }

  // x[i] -= row[i] * xk for i in [I0, I1): row is a shared-memory row of the factor (16-byte aligned when its
  // element index is even); pairs of consecutive entries are fetched with one 128-bit load
  template <int I0, int I1, int BASE>
  SMPC_DEV void axpy_row(const double * __restrict__ Mrow, double xk, double * x)
  {
    // BASE = parity information: element index of Mrow[0] modulo 2 (compile time)
    constexpr int first_even = ((BASE + I0) % 2 == 0) ? I0 : I0 + 1;
    if (first_even > I0 && I0 < I1)
      x[I0] = fma(-Mrow[I0], xk, x[I0]);
#pragma unroll
    for (int i = first_even; i + 1 < I1; i += 2)
    {
      double a, b;
      ld2(Mrow + i, a, b);
      x[i] = fma(-a, xk, x[i]);
      x[i + 1] = fma(-b, xk, x[i + 1]);
    }
    constexpr int npairs = (I1 - first_even > 0) ? (I1 - first_even) / 2 : 0;
    constexpr int last = first_even + 2 * npairs;
    if (last < I1 && last >= first_even)
      x[last] = fma(-Mrow[last], xk, x[last]);
  }

  template <int N, int LD, int K>
  struct SweepFwd
  {
    static SMPC_DEV void run(const double * __restrict__ M, const double * __restrict__ invd, double * x)
    {
      x[K] = x[K] * invd[K];
      axpy_row<K + 1, N, (K * LD) % 2>(M + K * LD, x[K], x); // upper slots of row K hold column K of L
      SMPC_SCHED_FENCE
      SweepFwd<N, LD, K + 1>::run(M, invd, x);
    }
  };
  template <int N, int LD>
  struct SweepFwd<N, LD, N>
  {
    static SMPC_DEV void run(const double *, const double *, double *) {}
  };
  template <int N, int LD, int K>
  struct SweepBwd
  {
    static SMPC_DEV void run(const double * __restrict__ M, const double * __restrict__ invd, double * x)
    {
      x[K] = x[K] * invd[K];
      axpy_row<0, K, (K * LD) % 2>(M + K * LD, x[K], x); // lower slots of row K hold row K of L
      SMPC_SCHED_FENCE
      SweepBwd<N, LD, K - 1>::run(M, invd, x);
    }
  };
  template <int N, int LD>
  struct SweepBwd<N, LD, -1>
  {
    static SMPC_DEV void run(const double *, const double *, double *) {}
  };

  // X <- L^-1 X (FWD) / X <- L^-T X (BWD) for the N x MC block X (ld LDX); factor from chol_rl (LD even).
  // One thread per column, the column in registers, no barrier inside.
  template <int N, int LD, int MC, int LDX, bool FWD, bool BWD>
  SMPC_DEV void sweep_cols(const double * __restrict__ M, const double * __restrict__ invd, double * __restrict__ X)
  {
    static_assert(MC <= 128, "one thread per column");
    static_assert(LD % 2 == 0, "factor rows must keep 16-byte alignment");
    SMPC_PAR(tid)
    const int c = tid;
    if (c < MC)
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

  // Structure of the kinodynamics stage dynamics (kino_knot.cuh, semi-implicit Euler with the joint accelerations as controls):
  // of the 2 NV rows of [A B] only the base rows [0, 6) and [NV, NV + 6) are dense; row k in [6, NV) holds A(k,k) = 1,
  // A(k,NV+k) = dt, B(k,NF+k-6) = dt^2 and row NV + r, r >= 6, holds A(NV+r,NV+r) = 1, B(NV+r,NF+r-6) = dt. Seen by columns, column c
  // of [A B] meets at most one sparse row in [6, NV) and one in [NV + 6, 2 NV). The two products that touch [A B] skip the
  // structural zeros: every output still receives its non-zero terms in ascending k, i.e. the same fused multiply-adds
  // in the same order as the dense product (adding an exact zero product changes nothing), at a third of the work.
  template <int NV, int NF>
  struct KinoRows
  {
    // sparse row of [A B] met by column c in [6, NV) (-1: none)
    static SMPC_DEV int s1(int c)
    {
      if (c < 2 * NV)
        return (c >= 6 && c < NV) ? c : ((c >= NV + 6) ? c - NV : -1);
      const int m = c - 2 * NV - NF;
      return m >= 0 ? 6 + m : -1;
    }
    // ... in [NV + 6, 2 NV)
    static SMPC_DEV int s2(int c)
    {
      if (c < 2 * NV)
        return (c >= NV + 6) ? c : -1;
      const int m = c - 2 * NV - NF;
      return m >= 0 ? NV + 6 + m : -1;
    }
  };

  // DZ: no constraint row of any stage depends on the control (D = 0: joint limits, contact-velocity and landing-height rows;
  // the kinodynamics problem without force cones, i.e. the reference's Go2 example). Then W = R^-1 D^T = 0, Sigma = mu I and
  // G = Sigma^-1 [d | C]: the W columns of the forward sweep, the Sigma build and factorisation, the sweeps of G and Y += W G
  // drop out. What is left performs exactly the operations the general path would (its other terms are products with exact
  // zeros), so the result is bitwise the same.
  template <int NDX, int NU, int NC, int TPB, int SNV = 0, bool DZ = false>
  SMPC_DEV void riccati3_backward_inst(
    const KnotLayout & L, const SolLayout & SL, int N, const double * __restrict__ knots, double * __restrict__ sol,
    double mudyn, double mueq, double * sm, int * flag_out, double * P0out, const double * Pterm = nullptr)
  {
    // Pterm != nullptr (a leg of the PARALLEL sweep, riccati_par.cuh): the cost-to-go [P | p] at node N is given (layout of
    // P0out) and node N's own knot is not touched - it belongs to the next leg
    typedef Ric3<NDX, NU, NC> R;
    constexpr int LDP = R::LDP, NAB = R::NAB, NRY = R::NRY;
    constexpr int TM = (TPB >= 256) ? 2 : 4; // register tile rows: smaller tiles when there are more threads to feed
    static_assert(TPB == 128 || TPB == 256, "tile maps are written for 128 or 256 threads");
    double * KN = sm + R::oKN;
    double * PP = sm + R::oPP; // [NDX][LDP]: col 0 = p, cols 1.. = P
    double * Mb = sm + R::oM;
    double * PAB = sm + R::oPAB;   // rows [0, PR0)
    double * PAB1 = sm + R::oPAB1; // rows [PR0, NDX)
    constexpr int PR0 = R::PR0;
    double * YW = sm + R::oYW; // Mb, PAB and YW are the same buffer at different times (see Ric3)
    double * G = sm + R::oG;
    double * w = sm + R::ow;
    double * invp = sm + R::oinvp;
    double * invr = sm + R::oinvr;
    double * invs = sm + R::oinvs;
    double * colb = sm + R::ocol;
    void * bar = (void *)(sm + R::obar);
    int * bad = (int *)(sm + R::oflag);
    double * Qh = KN + R::kQ; // accumulate in place
    double * Sh = KN + R::kS;
    double * Rh = KN + R::kR;
    double * qh = KN + R::kq;
    double * rh = KN + R::kr;
    const double * Ak = KN + R::kA;
    const double * Bk = KN + R::kB;
    const double * fk = KN + R::kf;
    const double * Ck = KN + R::kC;
    const double * Dk = KN + R::kD;
    const double * dk = KN + R::kd;
    double * Sg = sm + R::oSg;
    const double inv_mu = 1.0 / mueq;
    unsigned parity = 0;

    // ---- terminal node: Z = C/mu, z = d/mu ; P = Q + C^T Z ; p = q + C^T z ----
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
            PP[e] = Pterm[a * LDP + (bb == 0 ? NDX : bb - 1)];
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
      // ---- 0. start the copy of this stage's knot (M = I + mudyn P was left in X by the previous symmetrisation) ----
      SMPC_PAR(tid)
      if (tid == 0)
        bulk_load(KN, kn, R::kC, bar); // Q S R q r A B f; C D d follow after step 3
      SMPC_END
      // ---- 1. [p~ | P~] = (I + mudyn P)^-1 [p | P], in place ----
      if (mudyn > 0)
      {
        RIC_T(0)
        chol2<NDX, NDX, TPB>(Mb, colb, invp, bad);
        RIC_T(1)
        sweep_cols<NDX, NDX, LDP, LDP, true, false>(Mb, invp, PP);
        sweep_cols<NDX, NDX, LDP, LDP, false, true>(Mb, invp, PP);
        RIC_T(2)
      }
      // ---- 2. knot landed ; P~ p~ out (for the forward sweep) ; w = P~ f + p~ ; PAB = P~ [A B] ----
      SMPC_PAR(tid)
      {
        bulk_wait(bar, parity);
        for (int e = tid; e < NDX * LDP; e += SMPC_NT)
        {
          const int a = e / LDP, bb = e % LDP;
          if (bb == 0)
            so[SL.opt + a] = PP[e];
          else
            so[SL.oPt + a * NDX + (bb - 1)] = PP[e];
        }
        constexpr int TR = NDX / TM, TCA = NDX / 6, TC = NAB / 6;
        static_assert(TR * TC + NDX <= TPB, "one tile per thread plus the w rows");
        if (tid < TR * TC)
        {
          const int a0 = TM * (tid / TC), tc = tid % TC;
          const bool inA = tc < TCA;
          const double * src = inA ? Ak + 6 * tc : Bk + 6 * (tc - TCA);
          const int lds = inA ? NDX : NU;
          double * pab_dst = (a0 < PR0) ? PAB + a0 * NAB : PAB1 + (a0 - PR0) * NAB;
          double c[TM][6];
#pragma unroll
          for (int r = 0; r < TM; r++)
#pragma unroll
            for (int q = 0; q < 6; q++)
              c[r][q] = 0.0;
          if (SNV == 0)
          {
#pragma unroll 4
            for (int k = 0; k < NDX; k++)
            {
              double bv[6];
              ld2(src + k * lds, bv[0], bv[1]);
              ld2(src + k * lds + 2, bv[2], bv[3]);
              ld2(src + k * lds + 4, bv[4], bv[5]);
#pragma unroll
              for (int r = 0; r < TM; r++)
              {
                const double pk = PP[(a0 + r) * LDP + 1 + k];
#pragma unroll
                for (int q = 0; q < 6; q++)
                  c[r][q] = fma(pk, bv[q], c[r][q]);
              }
            }
          }
          else
          {
            // kinodynamics structure (KinoRows): dense rows [0, 6), the tile's sparse rows in [6, NV), dense rows [NV, NV + 6),
            // sparse rows in [NV + 6, 2 NV). A 6-column tile never straddles a structure block (NV, NF multiples of 6): its
            // sparse rows are six consecutive ones, met on the diagonal of a 6 x 6 sub-block.
            typedef KinoRows<(SNV > 0 ? SNV : 6), NU - (SNV > 0 ? SNV : 6) + 6> KR;
            static_assert(SNV == 0 || (SNV % 6 == 0 && (NU - SNV + 6) % 6 == 0 && 2 * SNV == NDX), "structure blocks are whole 6-column tiles");
            const int cg = 6 * tc; // first column of the tile in [A B]
#pragma unroll 1
            for (int half = 0; half < 2; half++)
            {
              const int kd = half ? SNV : 0;
#pragma unroll 2
              for (int k = kd; k < kd + 6; k++)
              {
                double bv[6];
                ld2(src + k * lds, bv[0], bv[1]);
                ld2(src + k * lds + 2, bv[2], bv[3]);
                ld2(src + k * lds + 4, bv[4], bv[5]);
#pragma unroll
                for (int r = 0; r < TM; r++)
                {
                  const double pk = PP[(a0 + r) * LDP + 1 + k];
#pragma unroll
                  for (int q = 0; q < 6; q++)
                    c[r][q] = fma(pk, bv[q], c[r][q]);
                }
              }
              const int ks = half ? KR::s2(cg) : KR::s1(cg);
              if (ks >= 0)
              {
#pragma unroll
                for (int q = 0; q < 6; q++)
                {
                  const double v = src[(ks + q) * lds + q];
#pragma unroll
                  for (int r = 0; r < TM; r++)
                    c[r][q] = fma(PP[(a0 + r) * LDP + 1 + ks + q], v, c[r][q]);
                }
              }
            }
          }
#pragma unroll
          for (int r = 0; r < TM; r++)
#pragma unroll
            for (int q = 0; q < 6; q++)
              pab_dst[r * NAB + 6 * tc + q] = c[r][q];
        }
        else if (tid - TR * TC < NDX)
        {
          const int a = tid - TR * TC;
          double acc = PP[a * LDP];
          for (int b = 0; b < NDX; b++)
            acc = fma(PP[a * LDP + 1 + b], fk[b], acc);
          w[a] = acc;
        }
      }
      SMPC_END
      parity ^= 1u;
      RIC_T(3)
      // ---- 3. [Qh | Sh] = [Q | S] + A^T PAB ; Rh = R + B^T PB ; qh = q + A^T w ; rh = r + B^T w (in place) ----
      SMPC_PAR(tid)
      {
        constexpr int TR = NDX / TM, TCA = NDX / 6, TC = NAB / 6, TRU = NU / TM, TCU = NU / 6;
        static_assert(TR * TC + TRU * TCU <= TPB, "one tile per thread");
        if (tid < TR * TC + TRU * TCU)
        {
          const bool isR = tid >= TR * TC;
          const int tt = isR ? tid - TR * TC : tid;
          // left operand: columns a0.. of A (Qh, Sh rows) or of B (Rh rows); right operand: columns b0.. of PAB
          const int ar = isR ? TM * (tt / TCU) : TM * (tt / TC);
          const int tc = isR ? tt % TCU : tt % TC;
          const double * lhs = isR ? Bk + ar : Ak + ar;
          const int ldl = isR ? NU : NDX;
          const int b0 = isR ? NDX + 6 * tc : 6 * tc;
          double * out;
          int ldo;
          if (isR)
            out = Rh + ar * NU + 6 * tc, ldo = NU;
          else if (tc < TCA)
            out = Qh + ar * NDX + 6 * tc, ldo = NDX;
          else
            out = Sh + ar * NU + 6 * (tc - TCA), ldo = NU;
          double c[TM][6];
#pragma unroll
          for (int r = 0; r < TM; r++)
#pragma unroll
            for (int q = 0; q < 6; q++)
              c[r][q] = out[r * ldo + q];
          // rows [0, PR0) of PAB live in X, the rest in the C D d block of the knot copy: same k order, two loops
          if (SNV == 0)
          {
#pragma unroll 1
            for (int half = 0; half < 2; half++)
            {
              const int k0 = half ? PR0 : 0, k1 = half ? NDX : PR0;
              const double * pab = half ? PAB1 - PR0 * NAB : PAB;
#pragma unroll 4
              for (int k = k0; k < k1; k++)
              {
                double av[TM], bv[6];
#pragma unroll
                for (int r = 0; r < TM; r += 2)
                  ld2(lhs + k * ldl + r, av[r], av[r + 1]);
                ld2(pab + k * NAB + b0, bv[0], bv[1]);
                ld2(pab + k * NAB + b0 + 2, bv[2], bv[3]);
                ld2(pab + k * NAB + b0 + 4, bv[4], bv[5]);
#pragma unroll
                for (int r = 0; r < TM; r++)
#pragma unroll
                  for (int q = 0; q < 6; q++)
                    c[r][q] = fma(av[r], bv[q], c[r][q]);
              }
            }
          }
          else
          {
            // kinodynamics structure: output row i = column i of [A B] meets the dense rows and at most two sparse rows
            typedef KinoRows<(SNV > 0 ? SNV : 6), NU - (SNV > 0 ? SNV : 6) + 6> KR;
            static_assert(SNV == 0 || SNV + 6 <= PR0, "the dense rows of PAB all live in X");
            const int ig = isR ? NDX + ar : ar; // first output row as a column of [A B]
#pragma unroll 1
            for (int half = 0; half < 2; half++)
            {
              const int kd = half ? SNV : 0;
#pragma unroll 2
              for (int k = kd; k < kd + 6; k++)
              {
                double av[TM], bv[6];
#pragma unroll
                for (int r = 0; r < TM; r += 2)
                  ld2(lhs + k * ldl + r, av[r], av[r + 1]);
                ld2(PAB + k * NAB + b0, bv[0], bv[1]);
                ld2(PAB + k * NAB + b0 + 2, bv[2], bv[3]);
                ld2(PAB + k * NAB + b0 + 4, bv[4], bv[5]);
#pragma unroll
                for (int r = 0; r < TM; r++)
#pragma unroll
                  for (int q = 0; q < 6; q++)
                    c[r][q] = fma(av[r], bv[q], c[r][q]);
              }
#pragma unroll
              for (int r = 0; r < TM; r++)
              {
                const int ks = half ? KR::s2(ig + r) : KR::s1(ig + r);
                if (ks >= 0)
                {
                  const double v = lhs[ks * ldl + r];
                  const double * prow = (ks < PR0 ? PAB + ks * NAB : PAB1 + (ks - PR0) * NAB) + b0;
                  double bv[6];
                  ld2(prow, bv[0], bv[1]);
                  ld2(prow + 2, bv[2], bv[3]);
                  ld2(prow + 4, bv[4], bv[5]);
#pragma unroll
                  for (int q = 0; q < 6; q++)
                    c[r][q] = fma(v, bv[q], c[r][q]);
                }
              }
            }
          }
#pragma unroll
          for (int r = 0; r < TM; r++)
#pragma unroll
            for (int q = 0; q < 6; q++)
              out[r * ldo + q] = c[r][q];
        }
        if (tid < NAB)
        {
          const bool isq = tid < NDX;
          const double * col = isq ? Ak + tid : Bk + (tid - NDX);
          const int ld = isq ? NDX : NU;
          double * dst = isq ? qh + tid : rh + (tid - NDX);
          double acc = *dst;
          if (SNV == 0)
          {
            for (int k = 0; k < NDX; k++)
              acc = fma(col[k * ld], w[k], acc);
          }
          else
          {
            typedef KinoRows<(SNV > 0 ? SNV : 6), NU - (SNV > 0 ? SNV : 6) + 6> KR;
            for (int half = 0; half < 2; half++)
            {
              const int kd = half ? SNV : 0;
              for (int k = kd; k < kd + 6; k++)
                acc = fma(col[k * ld], w[k], acc);
              const int ks = half ? KR::s2(tid) : KR::s1(tid);
              if (ks >= 0)
                acc = fma(col[ks * ld], w[ks], acc);
            }
          }
          *dst = acc;
        }
      }
      SMPC_END
      RIC_T(4)
      // ---- 4. KKT by block LDL^T ----
      // PAB is dead: fetch C D d into its tail rows' storage while R^ is factorised
      SMPC_PAR(tid)
      if (tid == 0)
        bulk_load(KN + R::kC, kn + R::kC, R::KSIZE - R::kC, bar);
      SMPC_END
      chol2<NU, NU, TPB>(Rh, colb, invr, bad);
      RIC_T(5)
      constexpr int NYW = DZ ? LDP : NRY; // columns of [Y | W] that exist
      SMPC_PAR(tid)
      bulk_wait(bar, parity);
      for (int e = tid; e < NU * NYW; e += SMPC_NT)
      {
        const int a = e / NYW, c = e % NYW;
        double v;
        if (c == 0)
          v = rh[a];
        else if (c <= NDX)
          v = Sh[(c - 1) * NU + a];
        else
          v = Dk[(c - NDX - 1) * NU + a];
        YW[a * NRY + c] = v;
      }
      SMPC_END
      parity ^= 1u;
      sweep_cols<NU, NU, NYW, NRY, true, false>(Rh, invr, YW);
      RIC_T(6)
      if (NC > 0 && DZ)
      {
        // G = Sigma^-1 [d | C] with Sigma = mu I: the two sweeps of the general path reduce to two scalings by 1 / sqrt(mu)
        SMPC_PAR(tid)
        {
          const double inv = pivot_rsqrt(mueq);
          for (int e = tid; e < NC * LDP; e += SMPC_NT)
          {
            const int i = e / LDP, c = e % LDP;
            const double g = (c == 0) ? dk[i] : Ck[i * NDX + (c - 1)];
            G[e] = (g * inv) * inv;
          }
        }
        SMPC_END
        RIC_T(9)
      }
      if (NC > 0 && !DZ)
      {
        // Sigma = mu I + W^T W (lower) and G = [d | C] - W^T Y : register tiles (4x3 on G, 4x4 blocks of the
        // lower triangle of Sigma), one job list strided over the CTA
        SMPC_PAR(tid)
        {
          static_assert(NC % 4 == 0 && NDX % 3 == 0, "tile shapes");
          constexpr int NB = NC / 4, GT = NB * (NDX / 3), GL = NB, ST = NB * (NB + 1) / 2, NJOBS = GT + GL + ST;
          for (int job = tid; job < NJOBS; job += SMPC_NT)
          {
            if (job < GT)
            {
              const int i0 = 4 * (job / (NDX / 3)), c0 = 3 * (job % (NDX / 3));
              double acc[4][3];
#pragma unroll
              for (int r = 0; r < 4; r++)
#pragma unroll
                for (int q = 0; q < 3; q++)
                  acc[r][q] = (c0 + q == 0) ? dk[i0 + r] : Ck[(i0 + r) * NDX + (c0 + q - 1)];
#pragma unroll 4
              for (int k = 0; k < NU; k++)
              {
                double wv[4], yv[3];
#pragma unroll
                for (int r = 0; r < 4; r++)
                  wv[r] = YW[k * NRY + LDP + i0 + r];
#pragma unroll
                for (int q = 0; q < 3; q++)
                  yv[q] = YW[k * NRY + c0 + q];
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                  for (int q = 0; q < 3; q++)
                    acc[r][q] = fma(-wv[r], yv[q], acc[r][q]);
              }
#pragma unroll
              for (int r = 0; r < 4; r++)
#pragma unroll
                for (int q = 0; q < 3; q++)
                  G[(i0 + r) * LDP + c0 + q] = acc[r][q];
            }
            else if (job < GT + GL)
            {
              const int i0 = 4 * (job - GT);
              double acc[4];
#pragma unroll
              for (int r = 0; r < 4; r++)
                acc[r] = Ck[(i0 + r) * NDX + (NDX - 1)];
#pragma unroll 4
              for (int k = 0; k < NU; k++)
              {
                const double yv = YW[k * NRY + NDX];
#pragma unroll
                for (int r = 0; r < 4; r++)
                  acc[r] = fma(-YW[k * NRY + LDP + i0 + r], yv, acc[r]);
              }
#pragma unroll
              for (int r = 0; r < 4; r++)
                G[(i0 + r) * LDP + NDX] = acc[r];
            }
            else
            {
              int t = job - GT - GL, bi = 0;
              while (t > bi)
              {
                t -= bi + 1;
                bi++;
              }
              const int i0 = 4 * bi, j0 = 4 * t;
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
                  if (j0 + q <= i0 + r)
                    Sg[(i0 + r) * NC + j0 + q] = acc[r][q];
            }
          }
        }
        SMPC_END
        RIC_T(7)
        chol2<NC, NC, TPB>(Sg, colb, invs, bad);
        RIC_T(8)
        sweep_cols<NC, NC, LDP, LDP, true, false>(Sg, invs, G);
        sweep_cols<NC, NC, LDP, LDP, false, true>(Sg, invs, G);
        RIC_T(9)
        // Y <- Y + W G : 4x3 register tiles, the last column by four-row strips
        SMPC_PAR(tid)
        {
          constexpr int NBU = NU / 4, YT = NBU * (NDX / 3), NJOBS = YT + NBU;
          for (int job = tid; job < NJOBS; job += SMPC_NT)
          {
            if (job < YT)
            {
              const int a0 = 4 * (job / (NDX / 3)), c0 = 3 * (job % (NDX / 3));
              double acc[4][3];
#pragma unroll
              for (int r = 0; r < 4; r++)
#pragma unroll
                for (int q = 0; q < 3; q++)
                  acc[r][q] = YW[(a0 + r) * NRY + c0 + q];
#pragma unroll 4
              for (int i = 0; i < NC; i++)
              {
                double wv[4], gv[3];
#pragma unroll
                for (int r = 0; r < 4; r++)
                  wv[r] = YW[(a0 + r) * NRY + LDP + i];
#pragma unroll
                for (int q = 0; q < 3; q++)
                  gv[q] = G[i * LDP + c0 + q];
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                  for (int q = 0; q < 3; q++)
                    acc[r][q] = fma(wv[r], gv[q], acc[r][q]);
              }
#pragma unroll
              for (int r = 0; r < 4; r++)
#pragma unroll
                for (int q = 0; q < 3; q++)
                  YW[(a0 + r) * NRY + c0 + q] = acc[r][q];
            }
            else
            {
              const int a0 = 4 * (job - YT);
              double acc[4];
#pragma unroll
              for (int r = 0; r < 4; r++)
                acc[r] = YW[(a0 + r) * NRY + NDX];
#pragma unroll 4
              for (int i = 0; i < NC; i++)
              {
                const double gv = G[i * LDP + NDX];
#pragma unroll
                for (int r = 0; r < 4; r++)
                  acc[r] = fma(YW[(a0 + r) * NRY + LDP + i], gv, acc[r]);
              }
#pragma unroll
              for (int r = 0; r < 4; r++)
                YW[(a0 + r) * NRY + NDX] = acc[r];
            }
          }
        }
        SMPC_END
      }
      RIC_T(10)
      sweep_cols<NU, NU, LDP, NRY, false, true>(Rh, invr, YW);
      RIC_T(11)
      // ---- gains: [k | K] = -Y, [z | Z] = G ----
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
      RIC_T(12)
      // ---- 5. [p | P] = [qh | Qh] + Sh [k | K] + C^T [z | Z] : TM x 3 register tiles, p by the remaining threads ----
      SMPC_PAR(tid)
      {
        constexpr int TR = NDX / TM, TC = NDX / 3, PR = TM / 2; // PR rows of p per leftover thread
        static_assert(TR * TC + (NDX + PR - 1) / PR <= TPB, "one tile per thread plus the p rows");
        if (tid < TR * TC)
        {
          const int a0 = TM * (tid / TC), b0 = 3 * (tid % TC);
          double c[TM][3];
#pragma unroll
          for (int r = 0; r < TM; r++)
#pragma unroll
            for (int q = 0; q < 3; q++)
              c[r][q] = Qh[(a0 + r) * NDX + b0 + q];
#pragma unroll 4
          for (int k = 0; k < NU; k++)
          {
            double kv[3];
#pragma unroll
            for (int q = 0; q < 3; q++)
              kv[q] = YW[k * NRY + 1 + b0 + q];
#pragma unroll
            for (int r = 0; r < TM; r++)
            {
              const double sv = Sh[(a0 + r) * NU + k];
#pragma unroll
              for (int q = 0; q < 3; q++)
                c[r][q] = fma(sv, kv[q], c[r][q]);
            }
          }
#pragma unroll 4
          for (int i = 0; i < NC; i++)
          {
            double cv[TM], zv[3];
#pragma unroll
            for (int r = 0; r < TM; r += 2)
              ld2(Ck + i * NDX + a0 + r, cv[r], cv[r + 1]);
#pragma unroll
            for (int q = 0; q < 3; q++)
              zv[q] = G[i * LDP + 1 + b0 + q];
#pragma unroll
            for (int r = 0; r < TM; r++)
#pragma unroll
              for (int q = 0; q < 3; q++)
                c[r][q] = fma(cv[r], zv[q], c[r][q]);
          }
#pragma unroll
          for (int r = 0; r < TM; r++)
#pragma unroll
            for (int q = 0; q < 3; q++)
              PP[(a0 + r) * LDP + 1 + b0 + q] = c[r][q];
        }
        else
        {
          // p: PR rows per thread, independent chains
          const int a1 = PR * (tid - TR * TC);
          if (a1 < NDX)
          {
            double pv[PR];
            int ar[PR];
#pragma unroll
            for (int r = 0; r < PR; r++)
            {
              ar[r] = (a1 + r < NDX) ? a1 + r : a1;
              pv[r] = qh[ar[r]];
            }
            for (int k = 0; k < NU; k++)
            {
              const double kv = YW[k * NRY];
#pragma unroll
              for (int r = 0; r < PR; r++)
                pv[r] = fma(Sh[ar[r] * NU + k], kv, pv[r]);
            }
            for (int i = 0; i < NC; i++)
            {
              const double zv = G[i * LDP];
#pragma unroll
              for (int r = 0; r < PR; r++)
                pv[r] = fma(Ck[i * NDX + ar[r]], zv, pv[r]);
            }
#pragma unroll
            for (int r = 0; r < PR; r++)
              if (a1 + r < NDX)
                PP[(a1 + r) * LDP] = pv[r];
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
