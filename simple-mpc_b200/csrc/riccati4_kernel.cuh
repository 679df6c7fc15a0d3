// riccati4_kernel.cuh — the generic-shape TMA-staged Riccati backward kernel (riccati4.cuh) and its host-side setup.
// Included by the ric4_*.cu translation units, each of which instantiates the (ndx, nu, ncmax) shapes of one robot.
#pragma once
#include "ctx.h"
#include "riccati4.cuh"

namespace smpc
{
  // resident CTAs per SM the shared-memory footprint allows (228 KB per SM, 1 KB reserved per CTA), capped by the
  // register file (TPB x CTAs x registers <= 64 K)
  template <int NDX, int NU, int NC, int TPB>
  constexpr int ric4_ctas_per_sm()
  {
    // 256 threads = the latency variant: one instance per SM with the whole register file (255 registers, no spills)
    constexpr int fit = 233472 / (Ric4<NDX, NU, NC>::SIZE * 8 + 1024 + 16);
    // at least 128 registers per thread; the small stages of the centroidal problem (ndx 12) live with 80
    constexpr int cap = TPB >= 256 ? 1 : (NDX <= 16 ? 768 / TPB : 512 / TPB);
    return fit < 1 ? 1 : (fit > cap ? cap : fit);
  }
  // GN: the problem's state dimension when it is not a multiple of 4 (padded mode of riccati4.cuh), else = NDX
  template <int NDX, int NU, int NC, int TPB, int GN>
  __global__ void __launch_bounds__(TPB, (ric4_ctas_per_sm<NDX, NU, NC, TPB>())) k_riccati4(Arrays A)
  {
    extern __shared__ __align__(16) double smem_raw[];
    const int b = blockIdx.x;
    const Dims & d = A.dm;
    InstState & st = A.inst[b];
    if (st.phase != PH_NEED_STEP)
      return;
    const size_t xb = (size_t)b * (d.N + 1);
    __shared__ int flag;
    riccati4_backward_inst<NDX, NU, NC, TPB, GN>(
      A.L, A.SL, d.N, A.knots + xb * A.L.size, A.sol + xb * A.SL.size, st.mu, st.mu, smem_raw, &flag,
      A.P0 + (size_t)b * d.ndx * (d.ndx + 1));
    if (threadIdx.x == 0 && flag)
      st.flags |= 1;
  }
  // one CTA per leg of the PARALLEL sweep, see k_riccati3_legs
  template <int NDX, int NU, int NC, int TPB, int GN>
  __global__ void __launch_bounds__(TPB, (ric4_ctas_per_sm<NDX, NU, NC, TPB>())) k_riccati4_legs(Arrays A, int pass)
  {
    extern __shared__ __align__(16) double smem_raw[];
    const Dims & d = A.dm;
    const int J = A.legs, b = blockIdx.x / (J - 1), j = blockIdx.x % (J - 1) + (pass == 1 ? 1 : 0);
    InstState & st = A.inst[b];
    if (st.phase != PH_NEED_STEP)
      return;
    const size_t pp = leg_pp_size(d.ndx);
    __shared__ int flag;
    if (pass == 1)
    {
      if (j == 1 && threadIdx.x == 0)
        A.legFail[b] = 0;
    }
    else if (A.legFail[b])
    {
      // the condensed system could not be factorised (a cost-to-go that is not positive definite): the CTA of leg 0 runs the
      // serial sweep over the whole horizon instead, the other legs have nothing to do
      if (j != 0)
        return;
      const size_t x0 = (size_t)b * (d.N + 1);
      riccati4_backward_inst<NDX, NU, NC, TPB, GN>(
        A.L, A.SL, d.N, A.knots + x0 * A.L.size, A.sol + x0 * A.SL.size, st.mu, st.mu, smem_raw, &flag, A.P0 + (size_t)b * pp);
      if (threadIdx.x == 0)
        atomicOr(&st.flags, flag ? 5 : 4);
      return;
    }
    const int a0 = leg_begin(d.N, J, j), a1 = leg_begin(d.N, J, j + 1);
    const size_t xb = (size_t)b * (d.N + 1) + a0;
    const double * Pterm = pass == 1 ? (j == J - 1 ? nullptr : A.legZero) : A.legT + ((size_t)b * J + j + 1) * pp;
    double * P0 = (pass != 1 && j == 0) ? A.P0 + (size_t)b * pp : A.legP0 + ((size_t)b * J + j) * pp;
    riccati4_backward_inst<NDX, NU, NC, TPB, GN>(
      A.L, A.SL, a1 - a0, A.knots + xb * A.L.size, A.sol + xb * A.SL.size, st.mu, st.mu, smem_raw, &flag, P0, Pterm);
    if (threadIdx.x == 0 && flag)
      atomicOr(&st.flags, 1);
  }
  template <int NDX, int NU, int NC, int TPB, int GN>
  void launch_riccati4_legs_impl(smpc_ctx * c, int pass)
  {
    k_riccati4_legs<NDX, NU, NC, TPB, GN><<<c->A.dm.B * (c->A.legs - 1), TPB, c->ric_smem, c->stream>>>(c->A, pass);
  }
  template <int NDX, int NU, int NC, int TPB, int GN>
  void launch_riccati4_impl(smpc_ctx * c)
  {
    k_riccati4<NDX, NU, NC, TPB, GN><<<c->A.dm.B, TPB, c->ric_smem, c->stream>>>(c->A);
  }
  // LEGS: also instantiate the leg kernel of the PARALLEL sweep (each instantiation of the sweep compiles for minutes, so
  // only the shapes whose single-instance latency matters carry it)
  template <int NDX, int NU, int NC, int TPB, int GN = NDX, bool LEGS = false>
  int setup_riccati4(smpc_ctx * c, const char * name)
  {
    typedef Ric4<NDX, NU, NC> R;
    const KnotLayout & L = c->A.L;
    if (GN != NDX)
    {
      if (L.ndx != GN || L.nu != NU || L.ncmax != NC)
      {
        smpc::err() = "riccati4 (padded): problem dimensions differ from the instantiated shape";
        return SMPC_ERR_INVALID;
      }
    }
    else if (L.oQ != R::kQ || L.oS != R::kS || L.oR != R::kR || L.oq != R::kq || L.or_ != R::kr || L.oA != R::kA || L.oB != R::kB
        || L.of != R::kf || L.oC != R::kC || L.oD != R::kD || L.od != R::kd || L.size != R::KSIZE)
    {
      smpc::err() = "riccati4: compile-time knot layout differs from make_knot_layout";
      return SMPC_ERR_INVALID;
    }
    c->ric_smem = (size_t)R::SIZE * sizeof(double);
    CU(cudaFuncSetAttribute(k_riccati4<NDX, NU, NC, TPB, GN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->ric_smem));
    CU(cudaFuncSetAttribute(k_riccati4<NDX, NU, NC, TPB, GN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    c->launch_riccati = &launch_riccati4_impl<NDX, NU, NC, TPB, GN>;
    if constexpr (LEGS)
    {
      CU(cudaFuncSetAttribute(k_riccati4_legs<NDX, NU, NC, TPB, GN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->ric_smem));
      CU(cudaFuncSetAttribute(k_riccati4_legs<NDX, NU, NC, TPB, GN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      c->launch_riccati_legs = &launch_riccati4_legs_impl<NDX, NU, NC, TPB, GN>;
    }
    c->riccati_variant = name;
    if (getenv("SMPC_VERBOSE"))
    {
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_riccati4<NDX, NU, NC, TPB, GN>, TPB, c->ric_smem);
      fprintf(stderr, "smpc: k_riccati4<%d,%d,%d> (%d threads) %zu B shared memory per instance, resident CTAs per SM: %d\n", NDX, NU, NC, TPB,
              c->ric_smem, nb);
    }
    return SMPC_OK;
  }
} // namespace smpc
