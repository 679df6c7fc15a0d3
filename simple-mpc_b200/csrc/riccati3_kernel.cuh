// riccati3_kernel.cuh — the TMA-staged Riccati backward kernel (riccati3.cuh) and its host-side setup template.
// Included by the ric3_*.cu translation units, each of which instantiates one (ndx, nu, ncmax) shape.
#pragma once
#include "ctx.h"
#include "riccati3.cuh"

namespace smpc
{
#ifndef SMPC_RIC_MINB
#define SMPC_RIC_MINB 2
#endif
#ifndef SMPC_RIC3_TPB
#define SMPC_RIC3_TPB 128
#endif
// per-instance sweep time with 1, 2, 3 CTAs per SM (ms at T = 50; only the ratios matter), see ric3_pick_ctas
#ifndef SMPC_RIC_T1
#define SMPC_RIC_T1 2.41
#define SMPC_RIC_T2 2.78
#define SMPC_RIC_T3 3.20
#endif
  // resident CTAs per SM the shared-memory footprint allows (228 KB per SM, 1 KB reserved per CTA): the register
  // budget follows it (3 CTAs x 128 threads -> 168 registers, no spills worth naming: ptxas reports 56 bytes)
  template <int NDX, int NU, int NC, int TPB>
  constexpr int ric3_ctas_per_sm()
  {
    return (233472 / (Ric3<NDX, NU, NC>::SIZE * 8 + 1024 + 16) >= 3 && TPB == 128) ? 3 : (TPB == 128 ? SMPC_RIC_MINB : 1);
  }
  // TPB = 128: throughput variant (three CTAs per SM). TPB = 256: latency variant for batches of at most one instance
  // per SM - twice the threads on the products and trailing updates of one instance, the whole register file to itself.
  // SNV > 0: the stage dynamics have the kinodynamics structure (riccati3.cuh KinoRows) with NV = SNV
  template <int NDX, int NU, int NC, int TPB, int SNV, bool DZ>
  __global__ void __launch_bounds__(TPB, (ric3_ctas_per_sm<NDX, NU, NC, TPB>())) k_riccati3(Arrays A)
  {
    extern __shared__ __align__(16) double smem_raw[];
    const int b = blockIdx.x;
    const Dims & d = A.dm;
    InstState & st = A.inst[b];
    if (st.phase != PH_NEED_STEP)
      return;
    const size_t xb = (size_t)b * (d.N + 1);
    __shared__ int flag;
    riccati3_backward_inst<NDX, NU, NC, TPB, SNV, DZ>(
      A.L, A.SL, d.N, A.knots + xb * A.L.size, A.sol + xb * A.SL.size, st.mu, st.mu, smem_raw, &flag,
      A.P0 + (size_t)b * d.ndx * (d.ndx + 1));
    if (threadIdx.x == 0 && flag)
      st.flags |= 1;
  }

  // One CTA per LEG of the PARALLEL sweep (riccati_par.cuh): grid = B x (legs - 1). pass 1: legs 1..J-1 from a zero cost-to-go
  // (the true terminal node for the last leg), (P_a, p_a) to legP0; pass 4: legs 0..J-2 from the true cost-to-go legT[j + 1]
  template <int NDX, int NU, int NC, int TPB, int SNV, bool DZ>
  __global__ void __launch_bounds__(TPB, (ric3_ctas_per_sm<NDX, NU, NC, TPB>())) k_riccati3_legs(Arrays A, int pass)
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
      riccati3_backward_inst<NDX, NU, NC, TPB, SNV, DZ>(
        A.L, A.SL, d.N, A.knots + x0 * A.L.size, A.sol + x0 * A.SL.size, st.mu, st.mu, smem_raw, &flag, A.P0 + (size_t)b * pp);
      if (threadIdx.x == 0)
        atomicOr(&st.flags, flag ? 5 : 4);
      return;
    }
    const int a0 = leg_begin(d.N, J, j), a1 = leg_begin(d.N, J, j + 1);
    const size_t xb = (size_t)b * (d.N + 1) + a0;
    const double * Pterm = pass == 1 ? (j == J - 1 ? nullptr : A.legZero) : A.legT + ((size_t)b * J + j + 1) * pp;
    double * P0 = (pass != 1 && j == 0) ? A.P0 + (size_t)b * pp : A.legP0 + ((size_t)b * J + j) * pp;
    riccati3_backward_inst<NDX, NU, NC, TPB, SNV, DZ>(
      A.L, A.SL, a1 - a0, A.knots + xb * A.L.size, A.sol + xb * A.SL.size, st.mu, st.mu, smem_raw, &flag, P0, Pterm);
    if (threadIdx.x == 0 && flag)
      atomicOr(&st.flags, 1);
  }

  // Wave model of the Riccati launch: an instance takes t[k] when k CTAs share its SM (measured on B200 for the
  // (36, 24, 24) sweep, relative units); `batch` instances run as full waves of k * sms CTAs plus a remainder that
  // runs at the occupancy it reaches.
  int ric3_pick_ctas(int batch, int sms, int fit)
  {
    static const double t[4] = {0.0, SMPC_RIC_T1, SMPC_RIC_T2, SMPC_RIC_T3};
    if (fit > 3)
      fit = 3;
    if (sms <= 0 || fit < 1)
      return fit;
    int best = fit;
    double best_time = 1e300;
    for (int k = fit; k >= 1; k--)
    {
      const int per_wave = k * sms;
      const int full = batch / per_wave, rem = batch % per_wave;
      const int krem = (rem + sms - 1) / sms;
      const double time = full * t[k] + (rem ? t[krem] : 0.0);
      if (time < best_time - 1e-12)
      {
        best_time = time;
        best = k;
      }
    }
    return best;
  }

  template <int NDX, int NU, int NC, int TPB, int SNV, bool DZ>
  void launch_riccati3_impl(smpc_ctx * c)
  {
    k_riccati3<NDX, NU, NC, TPB, SNV, DZ><<<c->A.dm.B, TPB, c->ric_smem, c->stream>>>(c->A);
  }
  template <int NDX, int NU, int NC, int TPB, int SNV, bool DZ>
  void launch_riccati3_legs_impl(smpc_ctx * c, int pass)
  {
    k_riccati3_legs<NDX, NU, NC, TPB, SNV, DZ><<<c->A.dm.B * (c->A.legs - 1), TPB, c->ric_smem, c->stream>>>(c->A, pass);
  }
  template <int NDX, int NU, int NC, int TPB, int SNV, bool DZ>
  int finish_setup_riccati3(smpc_ctx * c)
  {
    CU(cudaFuncSetAttribute(k_riccati3<NDX, NU, NC, TPB, SNV, DZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->ric_smem));
    CU(cudaFuncSetAttribute(k_riccati3<NDX, NU, NC, TPB, SNV, DZ>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    c->launch_riccati = &launch_riccati3_impl<NDX, NU, NC, TPB, SNV, DZ>;
    if constexpr (TPB == 256)
    {
      // the leg-parallel sweep only runs when SMs are idle, i.e. where the latency variant is selected
      CU(cudaFuncSetAttribute(k_riccati3_legs<NDX, NU, NC, TPB, SNV, DZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->ric_smem));
      CU(cudaFuncSetAttribute(k_riccati3_legs<NDX, NU, NC, TPB, SNV, DZ>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      c->launch_riccati_legs = &launch_riccati3_legs_impl<NDX, NU, NC, TPB, SNV, DZ>;
    }
    c->riccati_variant = SNV ? (DZ ? (TPB == 128 ? "tma-staged-kinostruct-D0" : "tma-staged-kinostruct-D0-256")
                                   : (TPB == 128 ? "tma-staged-kinostruct" : "tma-staged-kinostruct-256"))
                             : (TPB == 128 ? "tma-staged" : "tma-staged-256");
    if (getenv("SMPC_VERBOSE"))
    {
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_riccati3<NDX, NU, NC, TPB, SNV, DZ>, TPB, c->ric_smem);
      fprintf(stderr, "smpc: k_riccati3 (%d threads) %zu B shared memory per instance, resident CTAs per SM: %d\n", TPB, c->ric_smem, nb);
    }
    return SMPC_OK;
  }
  template <int NDX, int NU, int NC>
  int setup_riccati3(smpc_ctx * c)
  {
    typedef Ric3<NDX, NU, NC> R;
    const KnotLayout & L = c->A.L;
    if (L.oQ != R::kQ || L.oS != R::kS || L.oR != R::kR || L.oq != R::kq || L.or_ != R::kr || L.oA != R::kA || L.oB != R::kB
        || L.of != R::kf || L.oC != R::kC || L.oD != R::kD || L.od != R::kd || L.size != R::KSIZE)
    {
      smpc::err() = "riccati3: compile-time knot layout differs from make_knot_layout";
      return SMPC_ERR_INVALID;
    }
    c->ric_smem = (size_t)R::SIZE * sizeof(double);
    // Resident CTAs per SM. The footprint allows `fit`; a batch that does not fill whole waves at that occupancy can
    // finish sooner at a lower one (512 instances on 148 SMs: 444 + 68 at three per SM, 296 + 216 at two), because a
    // CTA runs faster the fewer neighbours share its SM. ric3_pick_ctas() minimises the wave model below; the
    // occupancy is imposed by asking for more dynamic shared memory than the kernel uses. SMPC_RIC_CTAS overrides.
    {
      const int fit = (int)(233472 / (c->ric_smem + 1024 + 16));
      int want = ric3_pick_ctas(c->A.dm.B, c->hmodel_sms, fit);
      if (const char * e = getenv("SMPC_RIC_CTAS"))
        want = atoi(e);
      if (want >= 1 && want < fit)
      {
        size_t req = (size_t)233472 / want - 1024 - 16;
        req &= ~(size_t)15;
        if (req > 232448 - 16)
          req = 232448 - 16;
        c->ric_smem = req;
      }
    }
    // latency variant when no SM holds more than one instance (SMPC_RIC3_TPB forces one of the two)
    bool wide = c->hmodel_sms > 0 && c->A.dm.B <= c->hmodel_sms;
    if (const char * e = getenv("SMPC_RIC3_TPB"))
      wide = atoi(e) == 256;
    // kinodynamics problems (kino_knot.cuh writes A, B with the KinoRows structure) take the structure-exploiting products;
    // SMPC_RIC_STRUCT=0 forces the dense ones (A/B timing)
    bool kino = c->hocp.kind == SMPC_OCP_KINODYNAMICS && 2 * c->hmodel.d.nv == NDX;
    if (const char * e = getenv("SMPC_RIC_STRUCT"))
      kino = kino && atoi(e) != 0;
    // ... and without force cones no constraint row depends on the control (DZ, riccati3.cuh)
    const bool dz = kino && !c->hocp.force_cone;
    if (kino && dz)
      return wide ? finish_setup_riccati3<NDX, NU, NC, 256, NDX / 2, true>(c) : finish_setup_riccati3<NDX, NU, NC, 128, NDX / 2, true>(c);
    if (kino)
      return wide ? finish_setup_riccati3<NDX, NU, NC, 256, NDX / 2, false>(c) : finish_setup_riccati3<NDX, NU, NC, 128, NDX / 2, false>(c);
    return wide ? finish_setup_riccati3<NDX, NU, NC, 256, 0, false>(c) : finish_setup_riccati3<NDX, NU, NC, 128, 0, false>(c);
  }
} // namespace smpc
