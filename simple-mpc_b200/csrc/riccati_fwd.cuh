// riccati_fwd.cuh — forward sweep of the proximal Riccati recursion with the stage matrices staged through shared
// memory (supersedes riccati_forward_inst of riccati.cuh on the hot path).
//
// Same recursion and the same per-element operation order as riccati_forward_inst / oracle/orc_solver.hpp (every
// output row is one serial fma chain over its columns, in column order), so the results are bitwise identical; what
// changes is how the operands arrive. The first version read K, Z, A, B, P~ straight from global memory inside the
// dot products with one ROW per thread: neighbouring threads touched addresses 288 bytes apart, every load was its
// own 32-byte sector and sat in front of a dependent chain (profiles/r01f: 30 cycles per issued instruction, 4.3 ms
// per launch at B = 4096 for 2.2 MB per instance). A version that copied each stage into shared memory with coalesced
// loads and then computed was slower still (1.76 vs 1.15 ms at B = 1024): nothing was in flight while a CTA computed.
// Here the stage blocks arrive by bulk asynchronous copies (TMA) into two landing buffers, stage t+1 in flight while
// stage t is computed out of shared memory.
#pragma once
#include "par.cuh"
#include "dev_types.h"

namespace smpc
{

  // Two landing buffers per CTA, each holding one stage as it lies in HBM: the solution block K k Z z P~ p~ (contiguous
  // in `sol`) and the A B f block of the dense knot (contiguous in `knots`), fetched by two bulk asynchronous copies
  // (TMA) that complete on one mbarrier per buffer. Stage t+1 is in flight while stage t is computed.
  struct FwdSmem
  {
    int obuf[2]; // landing buffers (nbuf = 1: both name the same buffer)
    int nbuf;    // 2: stage t+1 in flight while stage t is computed (lowest latency per instance, two CTAs per SM);
                 // 1: one buffer, the next stage fetched while the last phase of this one runs - five CTAs per SM hide each other's
                 //    fetches (highest throughput for batches many waves deep)
    int okn;     // offset of the A B f block inside a landing buffer (after the solution block)
    int odx, odu, oy, olam, opart, obar, size;
    int sol_doubles, kn_doubles;
  };
  inline FwdSmem make_fwd_smem(int ndx, int nu, int ncmax, int nbuf = 2)
  {
    const KnotLayout L = make_knot_layout(ndx, nu, ncmax);
    const SolLayout SL = make_sol_layout(ndx, nu, ncmax);
    FwdSmem F;
    F.sol_doubles = SL.size;                  // even by construction
    F.kn_doubles = (L.of + ndx - L.oA + 1) & ~1; // A B f (+ one pad double when odd; stays inside the knot)
    const int stage = F.sol_doubles + F.kn_doubles;
    int o = 0;
    F.nbuf = nbuf;
    F.obuf[0] = o, o += stage;
    if (nbuf == 2)
      F.obuf[1] = o, o += stage;
    else
      F.obuf[1] = F.obuf[0];
    F.okn = F.sol_doubles;
    F.odx = o, o += ndx;
    F.odu = o, o += nu;
    F.oy = o, o += ndx;
    F.olam = o, o += ndx;
    F.opart = o, o += 64;
    o = (o + 1) & ~1;
    F.obar = o, o += 2; // two 8-byte mbarriers
    F.size = (o + 1) & ~1;
    return F;
  }

  // ---- bulk asynchronous copies: one arm (expected bytes) + any number of copies per mbarrier phase ----
  SMPC_DEV void fwd_bar_init(void * bar)
  {
#ifdef __CUDACC__
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
#else
    (void)bar;
#endif
  }
  SMPC_DEV void fwd_bar_fence_init()
  {
#ifdef __CUDACC__
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
  }
  SMPC_DEV void fwd_arm(void * bar, unsigned bytes)
  {
#ifdef __CUDACC__
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
#else
    (void)bar;
    (void)bytes;
#endif
  }
  SMPC_DEV void fwd_copy(double * dst, const double * src, int ndoubles, void * bar)
  {
#ifdef __CUDACC__
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src),
                 "r"((unsigned)ndoubles * 8u), "r"(b)
                 : "memory");
#else
    for (int i = 0; i < ndoubles; i++)
      dst[i] = src[i];
    (void)bar;
#endif
  }
  SMPC_DEV void fwd_wait(void * bar, unsigned parity)
  {
#ifdef __CUDACC__
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "FWD_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra FWD_DONE_%=;\n"
      "bra FWD_WAIT_%=;\n"
      "FWD_DONE_%=:\n"
      "}\n" ::"r"(b),
      "r"(parity)
      : "memory");
#else
    (void)bar;
    (void)parity;
#endif
  }

  SMPC_DEV void riccati_forward_staged(
    const KnotLayout & L, const SolLayout & SL, const FwdSmem & F, int N, const double * __restrict__ knots, const int * __restrict__ ncs,
    const double * __restrict__ sol, const double * __restrict__ P0, double mudyn, const double * __restrict__ gx,
    const double * __restrict__ gu, const double * __restrict__ lplus, double * __restrict__ dxs, double * __restrict__ dus,
    double * __restrict__ vnew, double * __restrict__ lnew, double * dphi_out, double * sm)
  {
    const int ndx = L.ndx, nu = L.nu, ncmax = L.ncmax;
    double * dx = sm + F.odx;
    double * du = sm + F.odu;
    double * y = sm + F.oy;
    double * lam = sm + F.olam;
    double * part = sm + F.opart; // partial dphi accumulators, one per thread slot < 64
    unsigned long long * bars = reinterpret_cast<unsigned long long *>(sm + F.obar);
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
      if (tid == 0)
      {
        fwd_bar_init(bars + 0);
        fwd_bar_init(bars + 1);
        fwd_bar_fence_init();
      }
    }
    SMPC_END
    // prologue: stage 0 into buffer 0
    SMPC_PAR(tid)
    if (tid == 0 && N > 0)
    {
      fwd_arm(bars + 0, (unsigned)(F.sol_doubles + F.kn_doubles) * 8u);
      fwd_copy(sm + F.obuf[0], sol, F.sol_doubles, bars + 0);
      fwd_copy(sm + F.obuf[0] + F.okn, knots + L.oA, F.kn_doubles, bars + 0);
    }
    SMPC_END
    const bool two = F.nbuf == 2;
    for (int t = 0; t < N; t++)
    {
      const int cur = two ? (t & 1) : 0;
      const double * buf = sm + F.obuf[cur];
      const double * Ks = buf + SL.oK;
      const double * kv = buf + SL.ok;
      const double * Zs = buf + SL.oZ;
      const double * zv = buf + SL.oz;
      const double * Ps = buf + SL.oPt;
      const double * ptv = buf + SL.opt;
      const double * As = buf + F.okn;
      const double * Bs = As + (L.oB - L.oA);
      const double * fv = As + (L.of - L.oA);
      const int nc = ncs[t];
      SMPC_PAR(tid)
      {
        // stage t+1 goes into the other buffer (last read in stage t-1; every thread has passed a barrier since)
        if (two && tid == 0 && t + 1 < N)
        {
          void * nb = bars + (cur ^ 1);
          double * nbuf = sm + F.obuf[cur ^ 1];
          fwd_arm(nb, (unsigned)(F.sol_doubles + F.kn_doubles) * 8u);
          fwd_copy(nbuf, sol + (size_t)(t + 1) * SL.size, F.sol_doubles, nb);
          fwd_copy(nbuf + F.okn, knots + (size_t)(t + 1) * L.size + L.oA, F.kn_doubles, nb);
        }
        fwd_wait(bars + cur, (unsigned)((two ? (t >> 1) : t) & 1));
        // (the dot products are unrolled by 6 so that the shared-memory loads of a group issue ahead of its fma chain;
        // the chain order itself is unchanged)
        for (int a = tid; a < nu; a += SMPC_NT)
        {
          double acc = kv[a];
#pragma unroll 6
          for (int b = 0; b < ndx; b++)
            acc = fma(Ks[a * ndx + b], dx[b], acc);
          du[a] = acc;
          dus[(size_t)t * nu + a] = acc;
        }
        // the multiplier rows start after the control rows so that both groups run side by side
        for (int i = tid - nu; i < ncmax; i += SMPC_NT)
        {
          if (i < 0)
            continue;
          double acc = 0;
          if (i < nc)
          {
            acc = zv[i];
#pragma unroll 6
            for (int b = 0; b < ndx; b++)
              acc = fma(Zs[i * ndx + b], dx[b], acc);
          }
          vnew[(size_t)t * ncmax + i] = acc;
        }
      }
      SMPC_END
      SMPC_PAR(tid)
      {
        for (int a = tid; a < ndx; a += SMPC_NT)
        {
          double acc = fv[a];
#pragma unroll 6
          for (int b = 0; b < ndx; b++)
            acc = fma(As[a * ndx + b], dx[b], acc);
#pragma unroll 6
          for (int b = 0; b < nu; b++)
            acc = fma(Bs[a * nu + b], du[b], acc);
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
          double acc = ptv[a];
#pragma unroll 6
          for (int b = 0; b < ndx; b++)
            acc = fma(Ps[a * ndx + b], y[b], acc);
          lam[a] = acc;
          lnew[(size_t)(t + 1) * ndx + a] = acc;
        }
        for (int a = tid; a < nu; a += SMPC_NT)
          part[a & 63] += gu[(size_t)t * nu + a] * du[a];
      }
      SMPC_END
      SMPC_PAR(tid)
      {
        // one landing buffer: nothing reads it after the barrier above, fetch stage t+1 behind this last phase
        if (!two && tid == 0 && t + 1 < N)
        {
          double * nbuf = sm + F.obuf[0];
          fwd_arm(bars + 0, (unsigned)(F.sol_doubles + F.kn_doubles) * 8u);
          fwd_copy(nbuf, sol + (size_t)(t + 1) * SL.size, F.sol_doubles, bars + 0);
          fwd_copy(nbuf + F.okn, knots + (size_t)(t + 1) * L.size + L.oA, F.kn_doubles, bars + 0);
        }
        for (int a = tid; a < ndx; a += SMPC_NT)
        {
          const double v = y[a] - mudyn * lam[a];
          dx[a] = v;
          dxs[(size_t)(t + 1) * ndx + a] = v;
        }
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
#pragma unroll 6
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
