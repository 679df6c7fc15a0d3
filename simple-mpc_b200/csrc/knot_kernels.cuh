// knot_kernels.cuh — the per-knot kernels (one CTA per (instance, knot)) and their host-side launch / setup templates.
// Included by the knots_*.cu translation units, each of which instantiates one robot shape.
#pragma once
#include "ctx.h"
#include "centro_knot.cuh"
#include "full_knot.cuh"

namespace smpc
{
  // kinodynamics: one warp per knot. The kernels are latency-bound and their throughput follows the resident warps per SM
  // (6 -> 8 warps: 1.33x, profiles/r01e vs r01f). A one-warp CTA lives on one SM sub-partition (16 K registers), so the
  // steps are 2 warps per sub-partition (<= 255 registers), 3 (<= 168) and 4 (<= 128). Linearisation kernel: 18.1 KB of
  // scratch (round 2: vf_dv, dhF_du, the velocity-product terms and the dIntegrate blocks are rebuilt where they are read or
  // share storage with dead arrays) allow 12 CTAs/SM at 168 registers: 15.6 -> 14.2 ms per launch at B = 4096 (profiles/r02o).
#ifndef SMPC_KNOT_MINB
#define SMPC_KNOT_MINB 12
#endif
// the value-only (trial) instantiation needs 13.0 KB of the scratch: four warps per SM sub-partition at 128 registers
// (224 bytes of spills) measure 4.65 ms against 5.1 ms for three at 168 registers (B = 4096, profiles/r02o)
#ifndef SMPC_KNOT_MINB_TRIAL
#define SMPC_KNOT_MINB_TRIAL 16
#endif
  template <class Cfg, bool TRIAL>
#ifdef SMPC_KNOT_MAXREG
  __global__ void __maxnreg__(SMPC_KNOT_MAXREG) k_knots(Arrays A)
#else
  __global__ void __launch_bounds__(32, (TRIAL ? SMPC_KNOT_MINB_TRIAL : SMPC_KNOT_MINB)) k_knots(Arrays A)
#endif
  {
    extern __shared__ __align__(16) double smem_raw[];
    KinoScratch<Cfg> & s = *reinterpret_cast<KinoScratch<Cfg> *>(smem_raw);
    const Dims & d = A.dm;
    const int b = blockIdx.x / (d.N + 1);
    const int t = blockIdx.x % (d.N + 1);
    const InstState & st = A.inst[b];
    if (st.phase != (TRIAL ? PH_LS : PH_ACTIVE))
      return;
    KnotIO io;
    int nc_dummy;
    make_knot_io<TRIAL>(A, st, b, t, io, &nc_dummy);
    if (t < d.N)
      kino_stage_knot<Cfg, !TRIAL>(*A.model, *A.ocp, A.L, io, s);
    else
      kino_terminal_knot<Cfg, !TRIAL>(*A.model, *A.ocp, A.L, io, s);
  }

  // centroidal: one warp per knot
  template <class Cfg, bool TRIAL>
  __global__ void __launch_bounds__(32) k_knots_centro(Arrays A)
  {
    extern __shared__ __align__(16) double smem_raw[];
    CentroScratch<Cfg> & s = *reinterpret_cast<CentroScratch<Cfg> *>(smem_raw);
    const Dims & d = A.dm;
    const int b = blockIdx.x / (d.N + 1);
    const int t = blockIdx.x % (d.N + 1);
    const InstState & st = A.inst[b];
    if (st.phase != (TRIAL ? PH_LS : PH_ACTIVE))
      return;
    KnotIO io;
    int nc_dummy;
    make_knot_io<TRIAL>(A, st, b, t, io, &nc_dummy);
    if (t < d.N)
      centro_stage_knot<Cfg, !TRIAL>(*A.ocp, A.L, io, s);
    else
      centro_terminal_knot<Cfg, !TRIAL>(*A.ocp, A.L, io, s);
  }

  // full dynamics: one CTA per knot; the terminal node is the kinodynamics one (same terminal cost / constraint,
  // reference src/fulldynamics.cpp:418-445 vs src/kinodynamics.cpp:352-377)
  template <class Cfg, class TermCfg, bool TRIAL>
  __global__ void __launch_bounds__(128) k_knots_full(Arrays A)
  {
    extern __shared__ __align__(16) double smem_raw[];
    const Dims & d = A.dm;
    const int b = blockIdx.x / (d.N + 1);
    const int t = blockIdx.x % (d.N + 1);
    const InstState & st = A.inst[b];
    if (st.phase != (TRIAL ? PH_LS : PH_ACTIVE))
      return;
    KnotIO io;
    int nc_dummy;
    make_knot_io<TRIAL>(A, st, b, t, io, &nc_dummy);
    if (t < d.N)
      full_stage_knot<Cfg, !TRIAL>(*A.model, *A.ocp, A.L, io, *reinterpret_cast<FullScratch<Cfg> *>(smem_raw));
    else
      kino_terminal_knot<TermCfg, !TRIAL>(*A.model, *A.ocp, A.L, io, *reinterpret_cast<KinoScratch<TermCfg> *>(smem_raw));
  }

  template <class Cfg>
  void launch_knots_impl(smpc_ctx * c, bool trial)
  {
    const Dims & d = c->A.dm;
    const int grid = d.B * (d.N + 1);
    if (trial)
      k_knots<Cfg, true><<<grid, c->knot_threads, kino_trial_bytes<Cfg>(), c->stream>>>(c->A); // value-only scratch prefix
    else
      k_knots<Cfg, false><<<grid, c->knot_threads, c->knot_smem, c->stream>>>(c->A);
    c->launches++;
  }
  template <class Cfg>
  int setup_knots(smpc_ctx * c)
  {
    c->knot_smem = sizeof(KinoScratch<Cfg>);
    c->knot_threads = 32; // one warp per knot: phase barriers are warp syncs (64 / 128 threads measured slower, profiles/r01a)
    CU(cudaFuncSetAttribute(k_knots<Cfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->knot_smem));
    CU(cudaFuncSetAttribute(k_knots<Cfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->knot_smem));
    CU(cudaFuncSetAttribute(k_knots<Cfg, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_knots<Cfg, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    c->launch_knots = &launch_knots_impl<Cfg>;
    if (getenv("SMPC_VERBOSE"))
    {
      int nl = 0, nt = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nl, k_knots<Cfg, false>, c->knot_threads, c->knot_smem);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nt, k_knots<Cfg, true>, c->knot_threads, kino_trial_bytes<Cfg>());
      fprintf(stderr, "smpc: k_knots %zu B shared memory per knot, resident CTAs per SM: lin %d, trial %d\n", c->knot_smem, nl, nt);
    }
    return SMPC_OK;
  }

  template <class Cfg>
  void launch_knots_centro_impl(smpc_ctx * c, bool trial)
  {
    const Dims & d = c->A.dm;
    const int grid = d.B * (d.N + 1);
    if (trial)
      k_knots_centro<Cfg, true><<<grid, 32, c->knot_smem, c->stream>>>(c->A);
    else
      k_knots_centro<Cfg, false><<<grid, 32, c->knot_smem, c->stream>>>(c->A);
    c->launches++;
  }
  template <class Cfg>
  int setup_knots_centro(smpc_ctx * c)
  {
    c->knot_smem = sizeof(CentroScratch<Cfg>);
    CU(cudaFuncSetAttribute(k_knots_centro<Cfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->knot_smem));
    CU(cudaFuncSetAttribute(k_knots_centro<Cfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->knot_smem));
    c->launch_knots = &launch_knots_centro_impl<Cfg>;
    return SMPC_OK;
  }

  template <class Cfg, class TermCfg>
  void launch_knots_full_impl(smpc_ctx * c, bool trial)
  {
    const Dims & d = c->A.dm;
    const int grid = d.B * (d.N + 1);
    if (trial)
      k_knots_full<Cfg, TermCfg, true><<<grid, c->knot_threads, c->knot_smem, c->stream>>>(c->A);
    else
      k_knots_full<Cfg, TermCfg, false><<<grid, c->knot_threads, c->knot_smem, c->stream>>>(c->A);
    c->launches++;
  }
  template <class Cfg, class TermCfg>
  int setup_knots_full(smpc_ctx * c)
  {
    c->knot_smem = sizeof(FullScratch<Cfg>) > sizeof(KinoScratch<TermCfg>) ? sizeof(FullScratch<Cfg>) : sizeof(KinoScratch<TermCfg>);
    c->knot_threads = 64;
    if (const char * e = getenv("SMPC_KNOT_THREADS"))
      c->knot_threads = atoi(e) == 32 ? 32 : (atoi(e) == 128 ? 128 : 64);
    CU(cudaFuncSetAttribute(k_knots_full<Cfg, TermCfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->knot_smem));
    CU(cudaFuncSetAttribute(k_knots_full<Cfg, TermCfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->knot_smem));
    c->launch_knots = &launch_knots_full_impl<Cfg, TermCfg>;
    return SMPC_OK;
  }
} // namespace smpc
