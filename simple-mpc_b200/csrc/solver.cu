// solver.cu — kernels, device-resident solver state and the C ABI (include/smpc.h) of libsmpc_cuda.
//
// One context = `batch` independent MPC instances that share a robot model, a problem recipe and the
// solver options. One call of smpc_solve() is one SolverProxDDP::run per instance
// (reference src/mpc.cpp:85,212), executed as rounds of
//     linearize (CTA per knot)  ->  decide (CTA per instance)  ->  riccati backward (CTA per instance)
//     ->  forward sweep (CTA per instance)  ->  { trial evaluation (CTA per knot) -> linesearch decide }*
// on the context's stream. Per-instance control flow (convergence, BCL updates, step acceptance,
// regularisation) lives on the device in InstState; the host only reads two counters per round.
//
// There is NO CPU fallback: every entry point that computes fails with SMPC_ERR_CUDA if the
// CUDA runtime does.
#include "ctx.h"
#include "riccati_par.cuh"
#include "host_build.h"

using namespace smpc;

namespace smpc
{
  std::string & err()
  {
    thread_local std::string e;
    return e;
  }
} // namespace smpc

namespace
{
  // ------------------------------------------------------------------------------------------
  // kernels
  // ------------------------------------------------------------------------------------------
  __global__ void k_run_begin(Arrays A, int max_iters)
  {
    const int b = blockIdx.x;
    const Dims & d = A.dm;
    InstState & st = A.inst[b];
    // force_initial_condition_: xs[0] = x0 (reference src/mpc.cpp:53)
    for (int k = threadIdx.x; k < d.nx; k += blockDim.x)
      A.xs[(size_t)b * (d.N + 1) * d.nx + k] = A.x0[(size_t)b * d.nx + k];
    // prev multipliers <- current (start of run)
    for (int k = threadIdx.x; k < (d.N + 1) * d.ncmax; k += blockDim.x)
      A.prev_vs[(size_t)b * (d.N + 1) * d.ncmax + k] = A.vs[(size_t)b * (d.N + 1) * d.ncmax + k];
    for (int k = threadIdx.x; k < (d.N + 1) * d.ndx; k += blockDim.x)
      A.prev_lams[(size_t)b * (d.N + 1) * d.ndx + k] = A.lams[(size_t)b * (d.N + 1) * d.ndx + k];
    if (threadIdx.x == 0)
    {
      const smpc_solver_opts & o = A.opt;
      st.mu = o.mu_init;
      st.prim_tol = fmax(o.prim_tol0 * pow(st.mu, o.prim_alpha), o.tol);
      st.inner_tol = fmax(o.inner_tol0 * pow(st.mu, o.dual_alpha), o.tol);
      st.num_iters = 0;
      st.al_iters = 0;
      st.conv = 0;
      st.ls_trials = 0;
      st.flags = 0;
      st.alpha = 0;
      st.ls_mov_avg = 0; // the nonmonotone reference value is per run (see oracle/orc_solver.hpp run())
      st.ls_avg_weight = 0;
      st.max_iters = max_iters;
      st.phase = (max_iters > 0 && o.max_al_iters > 0) ? PH_ACTIVE : PH_DONE;
    }
  }

  // per-instance reductions and the ProxDDP control flow between linearisation and the Newton step
  __global__ void k_decide(Arrays A)
  {
    const int b = blockIdx.x;
    const Dims & d = A.dm;
    InstState & st = A.inst[b];
    // the phase is read ONCE per CTA before thread 0 rewrites it below (a late warp must not see the new value)
    __shared__ int copy_prev, s_phase;
    if (threadIdx.x == 0)
      s_phase = st.phase;
    __syncthreads();
    if (s_phase != PH_ACTIVE)
      return;
    if (threadIdx.x == 0)
    {
      const smpc_solver_opts & o = A.opt;
      const KnotScalars * sc = A.sc + (size_t)b * (d.N + 1);
      double cost = 0, pen = 0, pinf = 0, dinf = 0, crit = 0;
      for (int t = 0; t <= d.N; t++)
      {
        cost += sc[t].cost;
        pen += sc[t].pen;
        pinf = fmax(pinf, sc[t].pinf);
        if (t > 0)
          dinf = fmax(dinf, sc[t].pad[0]);
        if (t < d.N)
          dinf = fmax(dinf, sc[t].pad[1]);
        crit = fmax(crit, sc[t].crit);
      }
      const double inner_crit = fmax(dinf, crit);
      st.cost0 = cost;
      st.phi0 = cost + pen;
      st.info.traj_cost = cost;
      st.info.merit = cost + pen;
      st.info.prim_infeas = pinf;
      st.info.dual_infeas = dinf;
      copy_prev = 0;
      if (fmax(pinf, dinf) <= o.tol)
      {
        st.conv = 1;
        st.phase = PH_DONE;
      }
      else if (inner_crit <= st.inner_tol)
      {
        // inner problem solved without a step: BCL outer update
        if (pinf <= st.prim_tol)
        {
          st.prim_tol = fmax(st.prim_tol * pow(st.mu, o.prim_beta), o.tol);
          st.inner_tol = fmax(st.inner_tol * pow(st.mu, o.dual_beta), o.tol);
          copy_prev = 1; // MultiplierUpdateMode::NEWTON
        }
        else
        {
          st.mu = fmax(st.mu * o.mu_update_factor, o.mu_lower_bound);
          st.prim_tol = fmax(o.prim_tol0 * pow(st.mu, o.prim_alpha), o.tol);
          st.inner_tol = fmax(o.inner_tol0 * pow(st.mu, o.dual_alpha), o.tol);
        }
        st.al_iters++;
        st.phase = (st.al_iters < o.max_al_iters) ? PH_ACTIVE : PH_DONE;
      }
      else
        st.phase = PH_NEED_STEP;
    }
    __syncthreads();
    if (copy_prev)
    {
      for (int k = threadIdx.x; k < (d.N + 1) * d.ncmax; k += blockDim.x)
        A.prev_vs[(size_t)b * (d.N + 1) * d.ncmax + k] = A.vs[(size_t)b * (d.N + 1) * d.ncmax + k];
      for (int k = threadIdx.x; k < (d.N + 1) * d.ndx; k += blockDim.x)
        A.prev_lams[(size_t)b * (d.N + 1) * d.ndx + k] = A.lams[(size_t)b * (d.N + 1) * d.ndx + k];
    }
  }

  __global__ void __launch_bounds__(128) k_riccati(Arrays A)
  {
    extern __shared__ __align__(16) double smem_raw[];
    const int b = blockIdx.x;
    const Dims & d = A.dm;
    InstState & st = A.inst[b];
    if (st.phase != PH_NEED_STEP)
      return;
    const size_t xb = (size_t)b * (d.N + 1);
    __shared__ int flag;
    riccati_backward_inst(
      A.L, A.SL, A.RS, d.N, A.knots + xb * A.L.size, A.ncs + xb, A.sol + xb * A.SL.size, st.mu, st.mu, smem_raw, &flag,
      A.P0 + (size_t)b * d.ndx * (d.ndx + 1));
    if (threadIdx.x == 0 && flag)
      st.flags |= 1;
  }

  // shared tail of the two forward kernels: linesearch initialisation
  __device__ __forceinline__ void forward_epilogue(const Arrays & A, InstState & st, double dphi)
  {
    const smpc_solver_opts & o = A.opt;
    st.dphi0 = dphi;
    // nonmonotone (Zhang-Hager) reference value
    st.ls_mov_avg = o.ls_avg_eta * st.ls_avg_weight * st.ls_mov_avg + st.phi0;
    st.ls_avg_weight = o.ls_avg_eta * st.ls_avg_weight + 1.0;
    st.ls_mov_avg /= st.ls_avg_weight;
    st.alpha = 1.0;
    st.ls_trials = 0;
    st.phase = PH_LS;
  }
  // forward sweep, operands read straight from global memory: many small CTAs (18 per SM) keep enough loads in flight
  // when the batch fills the GPU several times over; measured faster than the staged kernel from B ~ 3000 up
  __global__ void __launch_bounds__(64) k_forward_direct(Arrays A)
  {
    __shared__ double sm[4 * SMPC_MAX_NDX + SMPC_MAX_NU + 72];
    __shared__ double dphi;
    const int b = blockIdx.x;
    const Dims & d = A.dm;
    InstState & st = A.inst[b];
    if (st.phase != PH_NEED_STEP)
      return;
    const size_t xb = (size_t)b * (d.N + 1);
    riccati_forward_inst(
      A.L, A.SL, d.N, A.knots + xb * A.L.size, A.ncs + xb, A.sol + xb * A.SL.size, A.P0 + (size_t)b * d.ndx * (d.ndx + 1),
      st.mu, A.gx + xb * d.ndx, A.gu + (size_t)b * d.N * d.nu, A.lplus + xb * d.ndx, A.dxs + xb * d.ndx,
      A.dus + (size_t)b * d.N * d.nu, A.vnew + xb * d.ncmax, A.lnew + xb * d.ndx, &dphi, sm);
    if (threadIdx.x == 0)
      forward_epilogue(A, st, dphi);
  }
  // forward sweep, stage blocks by TMA into two landing buffers (riccati_fwd.cuh): lowest latency per instance
  __global__ void __launch_bounds__(128) k_forward(Arrays A)
  {
    extern __shared__ __align__(16) double sm[];
    __shared__ double dphi;
    const int b = blockIdx.x;
    const Dims & d = A.dm;
    InstState & st = A.inst[b];
    if (st.phase != PH_NEED_STEP)
      return;
    const size_t xb = (size_t)b * (d.N + 1);
    riccati_forward_staged(
      A.L, A.SL, A.FS, d.N, A.knots + xb * A.L.size, A.ncs + xb, A.sol + xb * A.SL.size, A.P0 + (size_t)b * d.ndx * (d.ndx + 1),
      st.mu, A.gx + xb * d.ndx, A.gu + (size_t)b * d.N * d.nu, A.lplus + xb * d.ndx, A.dxs + xb * d.ndx,
      A.dus + (size_t)b * d.N * d.nu, A.vnew + xb * d.ncmax, A.lnew + xb * d.ndx, &dphi, sm);
    if (threadIdx.x == 0)
      forward_epilogue(A, st, dphi);
  }

  // reduce the trial merit, test the (nonmonotone) Armijo condition, accept or shrink the step
  __global__ void k_ls_decide(Arrays A)
  {
    const int b = blockIdx.x;
    const Dims & d = A.dm;
    InstState & st = A.inst[b];
    // the phase is read ONCE per CTA before thread 0 rewrites it below (a late warp must not see the new value)
    __shared__ int accept, s_phase;
    __shared__ double s_alpha;
    if (threadIdx.x == 0)
      s_phase = st.phase;
    __syncthreads();
    if (s_phase != PH_LS)
    {
      if (threadIdx.x == 0 && s_phase == PH_ACTIVE)
        atomicAdd(A.counters + 1, 1);
      return;
    }
    if (threadIdx.x == 0)
    {
      const smpc_solver_opts & o = A.opt;
      const KnotScalars * sc = A.sct + (size_t)b * (d.N + 1);
      double cost = 0, pen = 0;
      for (int t = 0; t <= d.N; t++)
      {
        cost += sc[t].cost;
        pen += sc[t].pen;
      }
      const double phi = cost + pen;
      st.ls_trials++;
      const double alpha = st.alpha;
      s_alpha = alpha;
      const bool ok = phi <= st.ls_mov_avg + o.ls_armijo_c1 * alpha * st.dphi0;
      if (ok || alpha <= o.ls_alpha_min || st.ls_trials >= o.ls_max_steps)
      {
        accept = 1;
        st.info.traj_cost = cost;
        st.info.merit = phi;
        if (alpha <= o.ls_alpha_min)
          st.preg = fmin(st.preg * o.reg_inc_k, o.reg_max);
        else
          st.preg = fmax(st.preg * o.reg_dec_k, o.reg_min);
        st.num_iters++;
        if (st.num_iters >= st.max_iters)
        {
          st.al_iters++;
          st.phase = PH_DONE;
        }
        else
        {
          st.phase = PH_ACTIVE;
          atomicAdd(A.counters + 1, 1);
        }
      }
      else
      {
        accept = 0;
        st.alpha = fmax(alpha * o.ls_contraction, o.ls_alpha_min);
        atomicAdd(A.counters + 0, 1);
      }
    }
    __syncthreads();
    if (!accept)
      return;
    const double alpha = s_alpha;
    const size_t xb = (size_t)b * (d.N + 1);
    for (int k = threadIdx.x; k < (d.N + 1) * d.nx; k += blockDim.x)
      A.xs[xb * d.nx + k] = A.txs[xb * d.nx + k];
    for (int k = threadIdx.x; k < d.N * d.nu; k += blockDim.x)
      A.us[(size_t)b * d.N * d.nu + k] = A.tus[(size_t)b * d.N * d.nu + k];
    for (int k = threadIdx.x; k < d.N * d.ndx; k += blockDim.x)
      A.xdot[(size_t)b * d.N * d.ndx + k] = A.xdott[(size_t)b * d.N * d.ndx + k];
    if (A.fc)
      for (int k = threadIdx.x; k < d.N * d.nk * 3; k += blockDim.x)
        A.fc[(size_t)b * d.N * d.nk * 3 + k] = A.fct[(size_t)b * d.N * d.nk * 3 + k];
    for (int k = threadIdx.x; k < (d.N + 1) * d.ncmax; k += blockDim.x)
    {
      const int t = k / d.ncmax, i = k % d.ncmax;
      if (i < A.ncs[xb + t])
      {
        const double v = A.vs[xb * d.ncmax + k];
        A.vs[xb * d.ncmax + k] = v + alpha * (A.vnew[xb * d.ncmax + k] - v);
      }
    }
    for (int k = threadIdx.x; k < (d.N + 1) * d.ndx; k += blockDim.x)
    {
      const double v = A.lams[xb * d.ndx + k];
      A.lams[xb * d.ndx + k] = v + alpha * (A.lnew[xb * d.ndx + k] - v);
    }
  }

  __global__ void k_finish(Arrays A)
  {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= A.dm.B)
      return;
    InstState & st = A.inst[b];
    st.info.num_iters = st.num_iters;
    st.info.al_iters = st.al_iters;
    st.info.conv = st.conv;
    st.info.ls_trials = st.ls_trials;
    st.info.flags = st.flags;
    st.info.alpha = st.alpha;
    st.info.mu = st.mu;
    st.info.preg = st.preg;
  }

  // replaceStageCircular + cycleProblem (reference src/mpc.cpp:225-226): new tail slot, duals follow
  __global__ void k_cycle(Arrays A, int first, const smpc_slot * tails, const int * ids)
  {
    const int b = first + blockIdx.x;
    const Dims & d = A.dm;
    InstState & st = A.inst[b];
    const size_t xb = (size_t)b * (d.N + 1);
    // the slot at `head` is horizon position 0: overwrite it with the new tail, then advance head
    {
      const int * src = reinterpret_cast<const int *>(ids ? A.table + ids[blockIdx.x] : tails + blockIdx.x);
      int * dst = reinterpret_cast<int *>(A.slots + (size_t)b * d.N + st.head);
      for (int k = threadIdx.x; k < (int)(sizeof(smpc_slot) / sizeof(int)); k += blockDim.x)
        dst[k] = src[k];
    }
    for (int i = threadIdx.x; i < d.ncmax; i += blockDim.x)
    {
      for (int t = 0; t + 1 < d.N; t++)
        A.vs[(xb + t) * d.ncmax + i] = A.vs[(xb + t + 1) * d.ncmax + i];
      A.vs[(xb + d.N - 1) * d.ncmax + i] = 0.0;
    }
    for (int i = threadIdx.x; i < d.ndx; i += blockDim.x)
    {
      const double first_l = A.lams[(xb + 1) * d.ndx + i];
      for (int t = 1; t < d.N; t++)
        A.lams[(xb + t) * d.ndx + i] = A.lams[(xb + t + 1) * d.ndx + i];
      A.lams[(xb + d.N) * d.ndx + i] = first_l;
    }
    __syncthreads();
    if (threadIdx.x == 0)
      st.head = (st.head + 1) % d.N;
  }

  // warm-start shift of reference src/mpc.cpp:201-207
  __global__ void k_shift(Arrays A, int first)
  {
    const int b = first + blockIdx.x;
    const Dims & d = A.dm;
    const size_t xb = (size_t)b * (d.N + 1);
    for (int i = threadIdx.x; i < d.nx; i += blockDim.x)
      for (int t = 0; t < d.N; t++)
        A.xs[(xb + t) * d.nx + i] = A.xs[(xb + t + 1) * d.nx + i];
    for (int i = threadIdx.x; i < d.nu; i += blockDim.x)
      for (int t = 0; t + 1 < d.N; t++)
        A.us[((size_t)b * d.N + t) * d.nu + i] = A.us[((size_t)b * d.N + t + 1) * d.nu + i];
  }

  // slots arrive in horizon order; store them at ring positions
  __global__ void k_store_slots(Arrays A, int first, const smpc_slot * src, int t_only)
  {
    const Dims & d = A.dm;
    const int per = (t_only < 0) ? d.N : 1;
    const int bi = blockIdx.x / per;
    const int t = (t_only < 0) ? blockIdx.x % per : t_only;
    const int b = first + bi;
    const InstState & st = A.inst[b];
    const int * s = reinterpret_cast<const int *>(src + blockIdx.x);
    int * dst = reinterpret_cast<int *>(A.slots + (size_t)b * d.N + (st.head + t) % d.N);
    for (int k = threadIdx.x; k < (int)(sizeof(smpc_slot) / sizeof(int)); k += blockDim.x)
      dst[k] = s[k];
  }
  __global__ void k_store_foot_refs(Arrays A, int first, const double * src)
  {
    const Dims & d = A.dm;
    const int bi = blockIdx.x / d.N, t = blockIdx.x % d.N;
    const int b = first + bi;
    const InstState & st = A.inst[b];
    smpc_slot & sl = A.slots[(size_t)b * d.N + (st.head + t) % d.N];
    for (int k = threadIdx.x; k < d.nk * 12; k += blockDim.x)
      sl.foot_ref[k / 12][k % 12] = src[((size_t)blockIdx.x * d.nk) * 12 + k];
  }

  // ---- batched MPC driver (mpc_driver.cuh) ----
  __global__ void k_mpc_init(Arrays A)
  {
    // every instance starts from the same reference configuration: instance 0 computes, the others copy
    __shared__ MpcScratch s;
    __shared__ double pose[SMPC_MAX_FEET][3];
    if (threadIdx.x == 0)
    {
      MpcInst tmp;
      mpc_init_instance(*A.model, A.model->d.q_ref, blockIdx.x == 0 ? A.mpcS : nullptr, tmp, s);
      for (int f = 0; f < A.dm.nk; f++)
        for (int k = 0; k < 3; k++)
          pose[f][k] = tmp.init_pose[f][k];
    }
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < A.dm.B)
      for (int f = 0; f < A.dm.nk; f++)
        for (int k = 0; k < 3; k++)
          A.mpcI[b].init_pose[f][k] = A.mpcI[b].final_pose[f][k] = pose[f][k];
  }
  // velocity_base_ / now_ setters: vb == nullptr -> mode switch (walking 0: standing, zero velocity; 1: walking)
  __global__ void k_mpc_set_vb(Arrays A, int first, int n, const double * vb, int walking)
  {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
      return;
    MpcInst & I = A.mpcI[first + i];
    if (vb)
      for (int k = 0; k < 6; k++)
        I.velocity_base[k] = vb[i * 6 + k];
    if (walking >= 0)
    {
      I.walking = walking;
      if (!walking)
        for (int k = 0; k < 6; k++)
          I.velocity_base[k] = 0.0;
    }
  }
  // recedeWithCycle: one thread per instance picks the new tail and updates the take-off / landing queues
  __global__ void k_mpc_recede(Arrays A)
  {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const Dims & d = A.dm;
    if (b >= d.B)
      return;
    MpcInst & I = A.mpcI[b];
    const unsigned tail_mask = slot_at(A, A.inst[b], b, d.N - 1).contact_mask;
    const int before = I.overflow;
    A.mpc_ids[b] = mpc_recede(*A.mpcS, I, A.table, A.ncycle, d.nk, d.N, tail_mask);
    if (I.overflow && !before)
      atomicAdd(A.counters + 2, 1);
  }
  // updateInternalData + updateStepTrackerReferences + getProblemState: one CTA per instance
  __global__ void __launch_bounds__(64) k_mpc_refs(Arrays A)
  {
    __shared__ MpcScratch s;
    const int b = blockIdx.x;
    const Dims & d = A.dm;
    MpcRing ring;
    ring.base = A.slots + (size_t)b * d.N;
    ring.head = A.inst[b].head;
    ring.N = d.N;
    mpc_refs(*A.model, *A.ocp, *A.mpcS, A.mpcI[b], A.xmeas + (size_t)b * A.nxm, ring, A.term[b], A.x0 + (size_t)b * d.nx, s);
  }
  // ring -> horizon order (getters)
  __global__ void k_load_slots(Arrays A, int first, smpc_slot * dst)
  {
    const Dims & d = A.dm;
    const int bi = blockIdx.x / d.N, t = blockIdx.x % d.N;
    const int b = first + bi;
    const int * s = reinterpret_cast<const int *>(&slot_at(A, A.inst[b], b, t));
    int * o = reinterpret_cast<int *>(dst + blockIdx.x);
    for (int k = threadIdx.x; k < (int)(sizeof(smpc_slot) / sizeof(int)); k += blockDim.x)
      o[k] = s[k];
  }
  // u = us[0] - Ks[0] * difference(x_meas, xs[0]) for one instance per CTA (32 threads); dx by lane 0 (Lie-group
  // difference of the free-flyer), then one control row per lane with the columns in order
  __global__ void __launch_bounds__(32) k_feedback(Arrays A, int first, const double * xmeas, double * u)
  {
    __shared__ double dx[SMPC_MAX_NDX];
    const Dims & d = A.dm;
    const int b = first + blockIdx.x;
    const double * x0 = A.xs + (size_t)b * (d.N + 1) * d.nx;
    const double * xm = xmeas + (size_t)blockIdx.x * d.nx;
    if (threadIdx.x == 0)
    {
      if (A.ocp->kind == SMPC_OCP_CENTROIDAL)
        for (int k = 0; k < d.ndx; k++)
          dx[k] = x0[k] - xm[k];
      else
      {
        const int nv = d.ndx / 2, nq = d.nx - nv;
        difference_q(nv, xm, x0, dx);
        for (int k = 0; k < nv; k++)
          dx[nv + k] = x0[nq + k] - xm[nq + k];
      }
    }
    __syncthreads();
    const double * K = A.sol + (size_t)b * (d.N + 1) * A.SL.size + A.SL.oK;
    const double * u0 = A.us + (size_t)b * d.N * d.nu;
    for (int a = threadIdx.x; a < d.nu; a += blockDim.x)
    {
      double acc = 0;
      for (int k = 0; k < d.ndx; k++)
        acc = fma(K[a * d.ndx + k], dx[k], acc);
      u[(size_t)blockIdx.x * d.nu + a] = u0[a] - acc;
    }
  }
  // Interpolator::interpolateState / interpolateLinear / interpolateContacts on the plan of one instance per CTA
  __global__ void __launch_bounds__(32) k_interpolate(Arrays A, int first, int step, double progress, double * xo, double * uo, int * co)
  {
    const Dims & d = A.dm;
    const int b = first + blockIdx.x;
    const double * xs = A.xs + (size_t)b * (d.N + 1) * d.nx;
    const double * us = A.us + (size_t)b * d.N * d.nu;
    if (xo)
    {
      double * o = xo + (size_t)blockIdx.x * d.nx;
      if (step >= d.N) // xs.size() - 1
      {
        for (int k = threadIdx.x; k < d.nx; k += blockDim.x)
          o[k] = xs[(size_t)d.N * d.nx + k];
      }
      else
      {
        const double * xa = xs + (size_t)step * d.nx, *xb = xa + d.nx;
        if (A.ocp->kind == SMPC_OCP_CENTROIDAL)
        {
          for (int k = threadIdx.x; k < d.nx; k += blockDim.x)
            o[k] = xb[k] * progress + xa[k] * (1. - progress);
        }
        else
        {
          const int nv = d.ndx / 2, nq = d.nx - nv;
          if (threadIdx.x == 0)
          {
            // pinocchio::interpolate(q0, q1, u) = q0 (+) u (q1 (-) q0)
            double dq[SMPC_MAX_NV];
            difference_q(nv, xa, xb, dq);
            for (int k = 0; k < nv; k++)
              dq[k] = progress * dq[k];
            integrate_q(nv, xa, dq, o);
          }
          for (int k = threadIdx.x; k < nv; k += blockDim.x)
            o[nq + k] = xb[nq + k] * progress + xa[nq + k] * (1. - progress);
        }
      }
    }
    if (uo)
    {
      double * o = uo + (size_t)blockIdx.x * d.nu;
      for (int k = threadIdx.x; k < d.nu; k += blockDim.x)
        o[k] = (step >= d.N - 1) ? us[(size_t)(d.N - 1) * d.nu + k]
                                 : us[(size_t)(step + 1) * d.nu + k] * progress + us[(size_t)step * d.nu + k] * (1. - progress);
    }
    if (co)
    {
      const int sc = step > d.N - 1 ? d.N - 1 : step;
      const unsigned mask = slot_at(A, A.inst[b], b, sc).contact_mask;
      for (int f = threadIdx.x; f < d.nk; f += blockDim.x)
        co[(size_t)blockIdx.x * d.nk + f] = (mask >> f) & 1u;
    }
  }
  // gather rows of a strided device array into a dense staging buffer: dst[r][0..w) = src[r*stride + off .. + w)
  __global__ void k_pack(const double * src, double * dst, size_t rows, int w, size_t stride, int off)
  {
    const size_t n = rows * (size_t)w;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
    {
      const size_t r = e / w;
      const int c = (int)(e % w);
      dst[e] = src[r * stride + off + c];
    }
  }
} // namespace

namespace
{
  void launch_riccati_generic(smpc_ctx * c)
  {
    k_riccati<<<c->A.dm.B, c->ric_generic_tpb, c->ric_smem, c->stream>>>(c->A);
  }
  // ---- PARALLEL (leg-condensed) sweep, riccati_par.cuh: pass 2 (one CTA per inner leg) and pass 3 (one CTA per instance) ----
  __global__ void __launch_bounds__(256) k_leg_elements(Arrays A)
  {
    extern __shared__ __align__(16) double smem_raw[];
    const Dims & d = A.dm;
    const int J = A.legs, b = blockIdx.x / (J - 2), j = 1 + blockIdx.x % (J - 2);
    InstState & st = A.inst[b];
    if (st.phase != PH_NEED_STEP)
      return;
    const size_t xb = (size_t)b * (d.N + 1);
    __shared__ int flag;
    leg_element_inst(
      A.L, A.SL, A.PS, leg_begin(d.N, J, j), leg_begin(d.N, J, j + 1), A.knots + xb * A.L.size, A.ncs + xb, A.sol + xb * A.SL.size,
      st.mu, st.mu, smem_raw, &flag, A.legE + ((size_t)b * J + j) * leg_elem_size(d.ndx));
    if (threadIdx.x == 0 && flag)
      atomicOr(&A.legFail[b], 1); // pass 4 falls back to the serial sweep
  }
  __global__ void __launch_bounds__(256) k_leg_coarse(Arrays A)
  {
    extern __shared__ __align__(16) double smem_raw[];
    const Dims & d = A.dm;
    const int J = A.legs, b = blockIdx.x;
    InstState & st = A.inst[b];
    if (st.phase != PH_NEED_STEP)
      return;
    __shared__ int flag;
    const size_t pp = leg_pp_size(d.ndx);
    leg_coarse_inst(
      d.ndx, J, A.CS, A.legP0 + (size_t)b * J * pp, A.legE + (size_t)b * J * leg_elem_size(d.ndx), A.legT + (size_t)b * J * pp, smem_raw,
      &flag);
    if (threadIdx.x == 0 && (flag || A.leg_force_fail))
      atomicOr(&A.legFail[b], 1); // pass 4 falls back to the serial sweep
  }
  // the backward sweep of one round: serial (one CTA per instance) or leg-parallel (four launches, riccati_par.cuh)
  void launch_backward(smpc_ctx * c)
  {
    const int J = c->A.legs;
    if (J <= 1)
    {
      c->launch_riccati(c);
      return;
    }
    const int B = c->A.dm.B;
    c->launch_riccati_legs(c, 1);
    if (J > 2)
      k_leg_elements<<<B * (J - 2), 256, (size_t)c->A.PS.size * sizeof(double), c->stream>>>(c->A);
    k_leg_coarse<<<B, 256, (size_t)c->A.CS.size * sizeof(double), c->stream>>>(c->A);
    c->launch_riccati_legs(c, 4);
    c->launches += J > 2 ? 3 : 2;
  }
  template <class T>
  int dev_alloc(smpc_ctx * c, T ** p, size_t n)
  {
    void * q = nullptr;
    CU(cudaMalloc(&q, n * sizeof(T)));
    CU(cudaMemsetAsync(q, 0, n * sizeof(T), c->stream));
    c->allocs.push_back(q);
    *p = reinterpret_cast<T *>(q);
    return SMPC_OK;
  }
  int ensure_staging(smpc_ctx * c, size_t bytes)
  {
    if (bytes > c->staging_bytes)
    {
      if (c->staging)
        CU(cudaFree(c->staging));
      CU(cudaMalloc(&c->staging, bytes));
      c->staging_bytes = bytes;
    }
    if (bytes > c->h_staging_bytes)
    {
      if (c->h_staging)
        CU(cudaFreeHost(c->h_staging));
      CU(cudaMallocHost(&c->h_staging, bytes));
      c->h_staging_bytes = bytes;
    }
    return SMPC_OK;
  }
  int check_range(smpc_ctx * c, int first, int n)
  {
    if (!c || first < 0 || n < 0 || first + n > c->A.dm.B)
    {
      smpc::err() = "instance range out of bounds";
      return SMPC_ERR_INVALID;
    }
    return 0;
  }
  // host -> device through the pinned staging buffer (async on the context stream, then safe to reuse after sync)
  int h2d(smpc_ctx * c, void * dst, const void * src, size_t bytes)
  {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return SMPC_OK;
  }
  int d2h(smpc_ctx * c, void * dst, const void * src, size_t bytes)
  {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    return SMPC_OK;
  }
} // namespace

extern "C" {

const char * smpc_last_error(void) { return smpc::err().c_str(); }

void smpc_default_solver_opts(smpc_solver_opts * o)
{
  o->tol = 1e-4;
  o->mu_init = 1e-8;
  o->max_iters = 1;
  o->max_al_iters = 100;
  o->prim_alpha = 0.1;
  o->prim_beta = 0.9;
  o->dual_alpha = 1.0;
  o->dual_beta = 1.0;
  o->mu_update_factor = 0.01;
  o->mu_lower_bound = 1e-8;
  o->inner_tol0 = 1.0;
  o->prim_tol0 = 1.0;
  o->reg_init = 1e-9;
  o->reg_min = 1e-10;
  o->reg_max = 1e9;
  o->reg_inc_k = 10.0;
  o->reg_inc_critical = 100.0;
  o->reg_dec_k = 1.0 / 3.0;
  o->ls_armijo_c1 = 1e-4;
  o->ls_alpha_min = 1e-7;
  o->ls_contraction = 0.5;
  o->ls_avg_eta = 0.85;
  o->ls_max_steps = 20;
  o->num_threads = 1;
}

int smpc_problem_dims(const smpc_model_desc * m, const smpc_ocp_desc * o, int * nx, int * ndx, int * nu)
{
  if (!m || !o)
  {
    smpc::err() = "null descriptor";
    return SMPC_ERR_INVALID;
  }
  const int fs = o->force_size;
  if (o->kind == SMPC_OCP_KINODYNAMICS)
  {
    *nx = m->nq + m->nv;
    *ndx = 2 * m->nv;
    *nu = m->nv - 6 + fs * m->nfeet;
  }
  else if (o->kind == SMPC_OCP_FULLDYNAMICS)
  {
    *nx = m->nq + m->nv;
    *ndx = 2 * m->nv;
    *nu = m->nv - 6;
  }
  else
  {
    *nx = *ndx = 9;
    *nu = fs * m->nfeet;
  }
  return SMPC_OK;
}

static int create_impl(smpc_ctx * c, int device, int batch, const smpc_model_desc * model, const smpc_ocp_desc * ocp, const smpc_solver_opts * opts);

int smpc_create(smpc_ctx ** out, int device, int batch, const smpc_model_desc * model, const smpc_ocp_desc * ocp, const smpc_solver_opts * opts)
{
  if (!out || !model || !ocp || !opts || batch < 1)
  {
    smpc::err() = "smpc_create: bad arguments";
    return SMPC_ERR_INVALID;
  }
  smpc_ctx * c = new smpc_ctx;
  // every failure below releases what was acquired so far (device allocations, stream, events, pinned memory)
  const int rc = create_impl(c, device, batch, model, ocp, opts);
  if (rc != SMPC_OK)
  {
    smpc_destroy(c);
    return rc;
  }
  *out = c;
  return SMPC_OK;
}

static int create_impl(smpc_ctx * c, int device, int batch, const smpc_model_desc * model, const smpc_ocp_desc * ocp, const smpc_solver_opts * opts)
{
  try
  {
    build_dev_model(*model, c->hmodel);
    build_dev_ocp(*model, *ocp, c->hocp);
  }
  catch (const std::exception & e)
  {
    smpc::err() = e.what();
    const bool unsup = smpc::err().find("CUDA") != std::string::npos || smpc::err().find("supported") != std::string::npos;
    return unsup ? SMPC_ERR_UNSUPPORTED : SMPC_ERR_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
  {
    smpc::err() = "no CUDA device: libsmpc_cuda has no CPU fallback";
    return SMPC_ERR_CUDA;
  }
  c->device = device;
  CU(cudaSetDevice(device));
  CU(cudaDeviceGetAttribute(&c->hmodel_sms, cudaDevAttrMultiProcessorCount, device));
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  Arrays & A = c->A;
  // Row capacity of the problem rounded UP to the nearest shape with a tuned Riccati sweep: rows beyond a stage's active count
  // are zero-padded anyway, and the padded sweep is several times faster than the run-time-shape sweep a configuration without
  // any inequality would otherwise fall to (the reference's benchmark/go2.cpp: no limits, no cones -> capacity 1; its
  // single-instance MPC::iterate went 7.05 -> 1.97 ms on the (36, 12, 24) sweep, profiles/r02ab, r02ad). SMPC_RICCATI != default keeps the exact capacity.
  if (!getenv("SMPC_RICCATI"))
  {
    static const int tuned[][3] = {{36, 24, 24}, {36, 24, 36}, {36, 12, 24}, {36, 12, 60}, {56, 22, 44}, {56, 22, 78}, {56, 34, 34}, {56, 34, 46}, {9, 12, 8}, {9, 12, 34}};
    int best = 0;
    for (const auto & t : tuned)
      if (t[0] == c->hocp.ndx && t[1] == c->hocp.nu && t[2] >= c->hocp.ncmax && (best == 0 || t[2] < best))
        best = t[2];
    if (best > 0)
      c->hocp.ncmax = best;
  }
  const DevOcp & O = c->hocp;
  A.dm = Dims{batch, O.N, O.nx, O.ndx, O.nu, O.ncmax, O.nk};
  A.L = make_knot_layout(O.ndx, O.nu, O.ncmax);
  A.SL = make_sol_layout(O.ndx, O.nu, O.ncmax);
  A.RS = make_ric_smem(O.ndx, O.nu, O.ncmax);
  // which forward kernel. The TMA-staged sweep with two landing buffers (stage t+1 in flight while stage t is computed, two CTAs
  // per SM) has the lowest latency when no SM holds more than one instance; beyond that the same sweep with ONE landing buffer
  // wins at every batch size: five CTAs per SM hide each other's fetches (B = 512: 0.31 vs 0.46 ms, B = 4096: 1.90 vs 3.19 ms, and
  // 2.76 ms for the direct kernel that reads rows straight from global memory - profiles/r02p). The direct kernel is left for
  // shapes whose stage does not fit the shared memory of an SM. SMPC_FWD = staged | staged1 | direct overrides.
  // the MPC tick of a batch that leaves SMs idle is latency-bound: replay it as one CUDA Graph (61 us of a 1.57 ms
  // single-instance tick, profiles/r02r; bitwise equal results). Larger batches gain nothing and keep the per-kernel timers
  c->use_graph = batch <= (c->hmodel_sms > 0 ? c->hmodel_sms : 148);
  if (const char * e = getenv("SMPC_GRAPH"))
    c->use_graph = atoi(e) != 0;
  {
    int nbuf = batch <= (c->hmodel_sms > 0 ? c->hmodel_sms : 148) ? 2 : 1;
    c->forward_staged = true;
    if (const char * e = getenv("SMPC_FWD"))
    {
      c->forward_staged = strcmp(e, "direct") != 0;
      nbuf = strcmp(e, "staged1") == 0 ? 1 : (strcmp(e, "staged") == 0 ? 2 : nbuf);
    }
    A.FS = make_fwd_smem(O.ndx, O.nu, O.ncmax, nbuf);
    if (nbuf == 2 && (size_t)A.FS.size * sizeof(double) > 232448 - 64)
      A.FS = make_fwd_smem(O.ndx, O.nu, O.ncmax, 1);
    const size_t fwd_bytes = (size_t)A.FS.size * sizeof(double);
    const bool fits = fwd_bytes <= 232448 - 64;
    c->forward_staged = c->forward_staged && fits;
    if (fits)
    {
      CU(cudaFuncSetAttribute(k_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_bytes));
      CU(cudaFuncSetAttribute(k_forward, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
  }
  A.opt = *opts;
  // kernel instantiation for this (robot, problem) shape; each family lives in its own translation unit
  int rc = SMPC_ERR_UNSUPPORTED;
  const bool go2_shape = model->njoints == 13 && model->nv == 18 && O.nk == 4 && O.fs == 3;
  const bool talos_shape = model->njoints == 23 && model->nv == 28 && O.nk == 2 && O.fs == 6;
  if (O.kind == SMPC_OCP_KINODYNAMICS && go2_shape)
    rc = setup_knots_kino_go2(c);
  else if (O.kind == SMPC_OCP_KINODYNAMICS && talos_shape)
    rc = setup_knots_kino_talos(c);
  else if (O.kind == SMPC_OCP_FULLDYNAMICS && go2_shape)
    rc = setup_knots_full_go2(c);
  else if (O.kind == SMPC_OCP_FULLDYNAMICS && talos_shape)
    rc = setup_knots_full_talos(c);
  else if (O.kind == SMPC_OCP_CENTROIDAL && ((O.nk == 2 && O.fs == 6) || (O.nk == 4 && O.fs == 3)))
    rc = setup_knots_centro(c);
  else
    smpc::err() = "no CUDA kernel instantiation for this robot/problem shape (built: kinodynamics and full dynamics on a 12-DoF quadruped with "
                  "point feet and on a 22-DoF biped with quad feet; centroidal with 2 quad feet or 4 point feet)";
  if (rc != SMPC_OK)
    return rc;
  // Riccati sweep: compile-time specialised kernel when one is instantiated for (ndx, nu, ncmax). Default: the fastest
  // measured per shape - riccati3.cuh (4 x 6 tiles laid out for 128 threads) for the Go2-kinodynamics shapes, riccati4.cuh
  // (any even shape) for the rest. SMPC_RICCATI (A/B timing): 4 = riccati4 wherever instantiated, 3 = riccati3 only,
  // 0 = the generic run-time-shape sweep of riccati.cuh
  const char * rv = getenv("SMPC_RICCATI");
  const int ric = rv ? atoi(rv) : -1;
  bool tuned = false;
  if (ric == -1 || ric == 3)
    rc = setup_ric3_go2(c, O.ndx, O.nu, O.ncmax, &tuned);
  if (rc == SMPC_OK && !tuned && (ric == -1 || ric == 4))
    rc = setup_ric4_go2(c, O.ndx, O.nu, O.ncmax, &tuned);
  if (rc == SMPC_OK && !tuned && (ric == -1 || ric == 4))
    rc = setup_ric4_big(c, O.ndx, O.nu, O.ncmax, &tuned);
  if (rc == SMPC_OK && !tuned && (ric == -1 || ric == 4))
    rc = setup_ric4_talos_kino(c, O.ndx, O.nu, O.ncmax, &tuned);
  if (rc == SMPC_OK && !tuned && (ric == -1 || ric == 4))
    rc = setup_ric4_centro(c, O.ndx, O.nu, O.ncmax, &tuned);
  if (rc != SMPC_OK)
    return rc;
  if (!tuned)
  {
    c->ric_smem = (size_t)A.RS.size * sizeof(double);
    CU(cudaFuncSetAttribute(k_riccati, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->ric_smem));
    c->launch_riccati = &launch_riccati_generic;
    // CTA width of the run-time-shape sweep: its phases are loops strided over the CTA, so any multiple of 32 works; measured on
    // the smallest stage there is (centroidal, ndx 9, nu 12, 34 rows) 128 threads still win (49 ms vs 53 / 63 ms for 64 / 32 at
    // B = 4096, T = 100), so 128 everywhere; SMPC_RIC_GENERIC_TPB overrides
    c->ric_generic_tpb = 128;
    if (const char * e = getenv("SMPC_RIC_GENERIC_TPB"))
      c->ric_generic_tpb = (atoi(e) == 32 || atoi(e) == 64) ? atoi(e) : 128;
  }
  const Dims & d = A.dm;
  const size_t B = d.B, N1 = d.N + 1;
  DevModel * dmodel;
  DevOcp * docp;
#define AL(p, n)                                                                                                       \
  if ((rc = dev_alloc(c, &(p), (n))) != SMPC_OK)                                                                       \
  return rc
  AL(dmodel, 1);
  AL(docp, 1);
  CU(cudaMemcpyAsync(dmodel, &c->hmodel, sizeof(DevModel), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(docp, &c->hocp, sizeof(DevOcp), cudaMemcpyHostToDevice, c->stream));
  A.model = dmodel;
  A.ocp = docp;
  AL(A.slots, B * d.N);
  AL(A.term, B);
  AL(A.x0, B * d.nx);
  AL(A.xs, B * N1 * d.nx);
  AL(A.us, B * d.N * d.nu);
  AL(A.txs, B * N1 * d.nx);
  AL(A.tus, B * d.N * d.nu);
  AL(A.vs, B * N1 * d.ncmax);
  AL(A.prev_vs, B * N1 * d.ncmax);
  AL(A.lams, B * N1 * d.ndx);
  AL(A.prev_lams, B * N1 * d.ndx);
  AL(A.knots, B * N1 * A.L.size);
  AL(A.ncs, B * N1);
  AL(A.sol, B * N1 * A.SL.size);
  AL(A.P0, B * d.ndx * (d.ndx + 1));
  AL(A.gx, B * N1 * d.ndx);
  AL(A.gu, B * d.N * d.nu);
  AL(A.lplus, B * N1 * d.ndx);
  AL(A.dxs, B * N1 * d.ndx);
  AL(A.dus, B * d.N * d.nu);
  AL(A.vnew, B * N1 * d.ncmax);
  AL(A.lnew, B * N1 * d.ndx);
  AL(A.sc, B * N1);
  AL(A.sct, B * N1);
  AL(A.xdot, B * d.N * d.ndx);
  AL(A.xdott, B * d.N * d.ndx);
  if (O.kind == SMPC_OCP_FULLDYNAMICS)
  {
    AL(A.fc, B * d.N * d.nk * 3);
    AL(A.fct, B * d.N * d.nk * 3);
  }
  AL(A.inst, B);
  AL(A.counters, 4);
  // PARALLEL Riccati sweep (smpc_solver_opts::num_threads > 1, reference src/mpc.cpp:46-50): only where it shortens the
  // critical path - SMs left idle by the batch, a tuned sweep for the shape, at least three legs (two legs are the serial
  // sweep cut in two) and a shape for which it was MEASURED to pay on B200 (profiles/r02v, single instance): the
  // centroidal problems (T = 100, 9 states: backward sweep 3.56 ms serial, 1.05 ms with 10 legs, 0.73 ms with 20), not the
  // Go2 kinodynamics problem (T = 50, 36 states: 1.04 ms serial against 1.35 ms with 10 legs - its run-time-shape element and
  // condensation kernels cost more than the tuned stages they save). SMPC_LEGS = n forces n legs for any supported shape.
  {
    int J = opts->num_threads;
    bool pays = O.kind == SMPC_OCP_CENTROIDAL;
    if (const char * e = getenv("SMPC_LEGS"))
    {
      J = atoi(e);
      pays = true;
    }
    J = riccati_legs_used(d.N, J);
    const int sms = c->hmodel_sms > 0 ? c->hmodel_sms : 148;
    A.legs = 1;
    A.PS = make_par_smem(d.ndx, d.nu, d.ncmax);
    A.CS = make_coarse_smem(d.ndx);
    const size_t par_bytes = (size_t)A.PS.size * sizeof(double), coarse_bytes = (size_t)A.CS.size * sizeof(double);
    if (J >= 3 && pays && c->launch_riccati_legs && (int)B * (J - 1) <= sms && par_bytes <= 227 * 1024 && coarse_bytes <= 227 * 1024)
    {
      A.legs = J;
      double * zero;
      AL(A.legP0, B * J * leg_pp_size(d.ndx));
      AL(A.legT, B * J * leg_pp_size(d.ndx));
      AL(A.legE, B * J * leg_elem_size(d.ndx));
      AL(zero, (size_t)leg_pp_size(d.ndx));
      A.legZero = zero;
      AL(A.legFail, B);
      A.leg_force_fail = getenv("SMPC_LEGS_FORCE_FALLBACK") != nullptr;
      CU(cudaFuncSetAttribute(k_leg_elements, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)par_bytes));
      CU(cudaFuncSetAttribute(k_leg_coarse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)coarse_bytes));
    }
  }
#undef AL
  // persistent solver scalars (aligator: set in setup())
  {
    std::vector<InstState> h(B);
    memset(h.data(), 0, B * sizeof(InstState));
    for (auto & s : h)
    {
      s.mu = opts->mu_init;
      s.preg = opts->reg_init;
    }
    CU(cudaMemcpyAsync(A.inst, h.data(), B * sizeof(InstState), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  CU(cudaMallocHost(&c->h_counters, 4 * sizeof(int)));
  CU(cudaEventCreate(&c->ev0));
  CU(cudaEventCreate(&c->ev1));
  CU(cudaEventCreate(&c->evr0));
  CU(cudaEventCreate(&c->evr1));
  for (cudaEvent_t & e : c->evk)
    CU(cudaEventCreate(&e));
  CU(cudaEventCreate(&c->evu0));
  CU(cudaEventCreate(&c->evu1));
  return SMPC_OK;
}

void smpc_destroy(smpc_ctx * c)
{
  if (!c)
    return;
  cudaSetDevice(c->device);
  if (c->stream)
    cudaStreamSynchronize(c->stream);
  for (void * p : c->allocs)
    cudaFree(p);
  if (c->staging)
    cudaFree(c->staging);
  if (c->h_staging)
    cudaFreeHost(c->h_staging);
  if (c->h_counters)
    cudaFreeHost(c->h_counters);
  for (cudaEvent_t e : {c->ev0, c->ev1, c->evr0, c->evr1, c->evu0, c->evu1, c->evk[0], c->evk[1], c->evk[2], c->evk[3], c->evk[4], c->evk[5]})
    if (e)
      cudaEventDestroy(e);
  if (c->tick_graph)
    cudaGraphExecDestroy(c->tick_graph);
  if (c->h_head)
    cudaFreeHost(c->h_head);
  if (c->stream)
    cudaStreamDestroy(c->stream);
  delete c;
}

int smpc_set_slots(smpc_ctx * c, int first, int n, const smpc_slot * slots)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  if (n == 0)
    return SMPC_OK;
  CU(cudaSetDevice(c->device));
  const size_t bytes = (size_t)n * c->A.dm.N * sizeof(smpc_slot);
  int rc;
  if ((rc = ensure_staging(c, bytes)) != SMPC_OK)
    return rc;
  if ((rc = h2d(c, c->staging, slots, bytes)) != SMPC_OK)
    return rc;
  k_store_slots<<<n * c->A.dm.N, 128, 0, c->stream>>>(c->A, first, (const smpc_slot *)c->staging, -1);
  c->launches++;
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}
int smpc_set_slot(smpc_ctx * c, int first, int n, int t, const smpc_slot * slots)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  if (t < 0 || t >= c->A.dm.N)
  {
    smpc::err() = "Stage index exceeds stage vector size";
    return SMPC_ERR_INVALID;
  }
  if (n == 0)
    return SMPC_OK;
  CU(cudaSetDevice(c->device));
  const size_t bytes = (size_t)n * sizeof(smpc_slot);
  int rc;
  if ((rc = ensure_staging(c, bytes)) != SMPC_OK)
    return rc;
  if ((rc = h2d(c, c->staging, slots, bytes)) != SMPC_OK)
    return rc;
  k_store_slots<<<n, 128, 0, c->stream>>>(c->A, first, (const smpc_slot *)c->staging, t);
  c->launches++;
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}
int smpc_set_foot_refs(smpc_ctx * c, int first, int n, const double * foot_ref)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  if (n == 0)
    return SMPC_OK;
  CU(cudaSetDevice(c->device));
  const Dims & d = c->A.dm;
  const size_t bytes = (size_t)n * d.N * d.nk * 12 * sizeof(double);
  int rc;
  if ((rc = ensure_staging(c, bytes)) != SMPC_OK)
    return rc;
  if ((rc = h2d(c, c->staging, foot_ref, bytes)) != SMPC_OK)
    return rc;
  k_store_foot_refs<<<n * d.N, 64, 0, c->stream>>>(c->A, first, (const double *)c->staging);
  c->launches++;
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}
int smpc_set_terminal(smpc_ctx * c, int first, int n, const smpc_terminal * term)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(c->A.term + first, term, (size_t)n * sizeof(smpc_terminal), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}
int smpc_cycle(smpc_ctx * c, int first, int n, const smpc_slot * new_tail)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  if (n == 0)
    return SMPC_OK;
  CU(cudaSetDevice(c->device));
  const size_t bytes = (size_t)n * sizeof(smpc_slot);
  int rc;
  if ((rc = ensure_staging(c, bytes)) != SMPC_OK)
    return rc;
  if ((rc = h2d(c, c->staging, new_tail, bytes)) != SMPC_OK)
    return rc;
  k_cycle<<<n, 64, 0, c->stream>>>(c->A, first, (const smpc_slot *)c->staging, nullptr);
  c->launches++;
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}
int smpc_upload_stage_table(smpc_ctx * c, const smpc_slot * stages, int n)
{
  if (!c || !stages || n < 1)
  {
    smpc::err() = "smpc_upload_stage_table: bad arguments";
    return SMPC_ERR_INVALID;
  }
  CU(cudaSetDevice(c->device));
  if (c->A.table)
  {
    // a new table replaces the previous one: release it
    CU(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < c->allocs.size(); i++)
      if (c->allocs[i] == (void *)c->A.table)
      {
        cudaFree(c->allocs[i]);
        c->allocs.erase(c->allocs.begin() + (long)i);
        break;
      }
    c->A.table = nullptr;
    c->A.ntable = 0;
  }
  smpc_slot * tbl = nullptr;
  int rc = dev_alloc(c, &tbl, (size_t)n);
  if (rc != SMPC_OK)
    return rc;
  CU(cudaMemcpyAsync(tbl, stages, (size_t)n * sizeof(smpc_slot), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->A.table = tbl;
  c->A.ntable = n;
  c->h_table.assign(stages, stages + n);
  c->mpc_ready = false; // the driver's timing tables refer to the previous stage table
  return SMPC_OK;
}
int smpc_cycle_ids(smpc_ctx * c, int first, int n, const int * ids)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  if (!c->A.table)
  {
    smpc::err() = "smpc_cycle_ids: no stage table uploaded";
    return SMPC_ERR_STATE;
  }
  for (int i = 0; i < n; i++)
    if (ids[i] < 0 || ids[i] >= c->A.ntable)
    {
      smpc::err() = "smpc_cycle_ids: stage id out of range";
      return SMPC_ERR_INVALID;
    }
  if (n == 0)
    return SMPC_OK;
  CU(cudaSetDevice(c->device));
  int rc;
  if ((rc = ensure_staging(c, (size_t)n * sizeof(int))) != SMPC_OK)
    return rc;
  CU(cudaMemcpyAsync(c->staging, ids, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  k_cycle<<<n, 64, 0, c->stream>>>(c->A, first, nullptr, (const int *)c->staging);
  c->launches++;
  // the staging buffer is reused by later calls: the copy above must have been consumed
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}
int smpc_set_warmstart(smpc_ctx * c, int first, int n, const double * xs, const double * us)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  c->head_cached = false;
  CU(cudaSetDevice(c->device));
  const Dims & d = c->A.dm;
  if (xs)
    CU(cudaMemcpyAsync(c->A.xs + (size_t)first * (d.N + 1) * d.nx, xs, (size_t)n * (d.N + 1) * d.nx * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (us)
    CU(cudaMemcpyAsync(c->A.us + (size_t)first * d.N * d.nu, us, (size_t)n * d.N * d.nu * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}
int smpc_shift_warmstart(smpc_ctx * c, int first, int n)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  if (n == 0)
    return SMPC_OK;
  c->head_cached = false;
  CU(cudaSetDevice(c->device));
  k_shift<<<n, 64, 0, c->stream>>>(c->A, first);
  c->launches++;
  CU(cudaGetLastError());
  return SMPC_OK;
}

int smpc_set_x0(smpc_ctx * c, const double * x0)
{
  if (!c || !x0)
  {
    smpc::err() = "smpc_set_x0: null argument";
    return SMPC_ERR_INVALID;
  }
  CU(cudaSetDevice(c->device));
  const Dims & d = c->A.dm;
  CU(cudaMemcpyAsync(c->A.x0, x0, (size_t)d.B * d.nx * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return SMPC_OK;
}

// One round of the solve up to the step direction: linearisation, decision, backward and forward sweeps. `tk`: record the
// per-kernel CUDA events of smpc_kernel_times (never inside a stream capture)
static int enqueue_round_head(smpc_ctx * c, bool tk)
{
  const Dims & d = c->A.dm;
  cudaStream_t s = c->stream;
  if (tk)
    CU(cudaEventRecord(c->evk[0], s));
  c->launch_knots(c, false);
  if (tk)
    CU(cudaEventRecord(c->evk[1], s));
  k_decide<<<d.B, 64, 0, s>>>(c->A);
  if (tk)
    CU(cudaEventRecord(c->evr0, s));
  launch_backward(c);
  if (tk)
  {
    CU(cudaEventRecord(c->evr1, s));
    CU(cudaEventRecord(c->evk[2], s));
  }
  if (c->forward_staged)
    k_forward<<<d.B, 128, (size_t)c->A.FS.size * sizeof(double), s>>>(c->A);
  else
    k_forward_direct<<<d.B, 64, 0, s>>>(c->A);
  if (tk)
    CU(cudaEventRecord(c->evk[3], s));
  c->launches += 4;
  return SMPC_OK;
}

// One linesearch trial: value-only knots at the trial point, the Armijo decision, the two counters to the host, and k_finish
// (scalars -> smpc_solve_info) behind every decision: when the step is accepted and no instance needs another round - the
// normal MPC tick - the synchronisation that follows ends the solve; otherwise the next decision's k_finish overwrites it
static int enqueue_trial(smpc_ctx * c, bool tk)
{
  const Dims & d = c->A.dm;
  cudaStream_t s = c->stream;
  if (tk)
    CU(cudaEventRecord(c->evk[4], s));
  // counters: [0] instances whose linesearch needs another trial, [1] instances that need another round
  // (both recounted by every k_ls_decide launch; the host reads the last one)
  CU(cudaMemsetAsync(c->A.counters, 0, 2 * sizeof(int), s));
  c->launch_knots(c, true);
  k_ls_decide<<<d.B, 128, 0, s>>>(c->A);
  if (tk)
    CU(cudaEventRecord(c->evk[5], s));
  CU(cudaMemcpyAsync(c->h_counters, c->A.counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  k_finish<<<(d.B + 127) / 128, 128, 0, s>>>(c->A);
  c->launches += 3;
  return SMPC_OK;
}

// The round / trial loops of a solve. `resumed`: the first round and its first trial have already run (the captured tick
// graph of smpc_mpc_iterate) and h_counters holds their outcome
static int solve_rounds(smpc_ctx * c, int max_iters, bool resumed)
{
  cudaStream_t s = c->stream;
  const int max_rounds = max_iters + c->A.opt.max_al_iters + 1;
  int rc;
  for (int round = 0; round < max_rounds; round++)
  {
    // the first round of a solve is timed kernel by kernel (CUDA events on this stream; smpc_kernel_times)
    const bool tk = round == 0 && !resumed;
    if (!(resumed && round == 0))
      if ((rc = enqueue_round_head(c, tk)) != SMPC_OK)
        return rc;
    for (int trial = 0; trial <= c->A.opt.ls_max_steps + 1; trial++)
    {
      if (!(resumed && round == 0 && trial == 0))
      {
        if ((rc = enqueue_trial(c, tk && trial == 0)) != SMPC_OK)
          return rc;
        CU(cudaEventRecord(c->ev1, s));
        CU(cudaStreamSynchronize(s));
      }
      if (c->h_counters[0] == 0)
        break;
    }
    if (c->h_counters[1] == 0)
      break;
  }
  CU(cudaGetLastError());
  return SMPC_OK;
}

int smpc_solve(smpc_ctx * c, const double * x0, int max_iters)
{
  if (!c)
  {
    smpc::err() = "smpc_solve: null context";
    return SMPC_ERR_INVALID;
  }
  CU(cudaSetDevice(c->device));
  const Dims & d = c->A.dm;
  if (x0)
  {
    int rc = smpc_set_x0(c, x0);
    if (rc != SMPC_OK)
      return rc;
  }
  cudaStream_t s = c->stream;
  c->head_cached = false;
  CU(cudaEventRecord(c->ev0, s));
  k_run_begin<<<d.B, 128, 0, s>>>(c->A, max_iters);
  c->launches++;
  int rc = solve_rounds(c, max_iters, false);
  if (rc != SMPC_OK)
    return rc;
  CU(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  CU(cudaEventElapsedTime(&c->last_ric_ms, c->evr0, c->evr1));
  c->acc_ric_ms += c->last_ric_ms;
  {
    float ms;
    CU(cudaEventElapsedTime(&ms, c->evk[0], c->evk[1]));
    c->acc_kernel_ms[0] += ms;
    c->acc_kernel_ms[1] += c->last_ric_ms;
    CU(cudaEventElapsedTime(&ms, c->evk[2], c->evk[3]));
    c->acc_kernel_ms[2] += ms;
    CU(cudaEventElapsedTime(&ms, c->evk[4], c->evk[5]));
    c->acc_kernel_ms[3] += ms;
  }
  c->acc_solve_ms += c->last_ms;
  c->acc_ric_launches += 1;
  c->acc_solves += 1;
  return SMPC_OK;
}

static int pack_to_host(smpc_ctx * c, const double * src, size_t rows, int w, size_t stride, int off, double * dst_host)
{
  const size_t bytes = rows * (size_t)w * sizeof(double);
  int rc;
  if ((rc = ensure_staging(c, bytes)) != SMPC_OK)
    return rc;
  const size_t n = rows * (size_t)w;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 16)
    grid = 148 * 16;
  k_pack<<<grid, 256, 0, c->stream>>>(src, (double *)c->staging, rows, w, stride, off);
  c->launches++;
  CU(cudaMemcpyAsync(dst_host, c->staging, bytes, cudaMemcpyDeviceToHost, c->stream));
  return SMPC_OK;
}

int smpc_get_results(smpc_ctx * c, int first, int n, double * xs, double * us, double * Ks, double * ks)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  const Dims & d = c->A.dm;
  const size_t N1 = d.N + 1;
  int rc;
  if (xs)
    if ((rc = d2h(c, xs, c->A.xs + (size_t)first * N1 * d.nx, (size_t)n * N1 * d.nx * sizeof(double))) != SMPC_OK)
      return rc;
  if (us)
    if ((rc = d2h(c, us, c->A.us + (size_t)first * d.N * d.nu, (size_t)n * d.N * d.nu * sizeof(double))) != SMPC_OK)
      return rc;
  // gains live in the per-stage solution records; rows t = 0..N-1 of each instance
  for (int i = 0; i < n && (Ks || ks); i++)
  {
    const double * base = c->A.sol + ((size_t)(first + i) * N1) * c->A.SL.size;
    if (Ks)
    {
      if ((rc = pack_to_host(c, base, d.N, d.nu * d.ndx, c->A.SL.size, c->A.SL.oK, Ks + (size_t)i * d.N * d.nu * d.ndx)) != SMPC_OK)
        return rc;
      CU(cudaStreamSynchronize(c->stream));
    }
    if (ks)
    {
      if ((rc = pack_to_host(c, base, d.N, d.nu, c->A.SL.size, c->A.SL.ok, ks + (size_t)i * d.N * d.nu)) != SMPC_OK)
        return rc;
      CU(cudaStreamSynchronize(c->stream));
    }
  }
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}

int smpc_get_head(smpc_ctx * c, int first, int n, double * x1, double * u0, double * u1, double * K0)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  const Dims & d = c->A.dm;
  const size_t N1 = d.N + 1;
  cudaStream_t s = c->stream;
  if (c->head_cached && c->h_head)
  {
    // written by the tick graph (capture_tick_graph): [B][nx] | [B][nu] | [B][nu] | [B][nu * ndx]
    const double * h = c->h_head;
    const size_t K = (size_t)d.nu * d.ndx;
    if (x1)
      memcpy(x1, h + (size_t)first * d.nx, (size_t)n * d.nx * sizeof(double));
    h += (size_t)d.B * d.nx;
    if (u0)
      memcpy(u0, h + (size_t)first * d.nu, (size_t)n * d.nu * sizeof(double));
    h += (size_t)d.B * d.nu;
    if (u1)
      memcpy(u1, h + (size_t)first * d.nu, (size_t)n * d.nu * sizeof(double));
    h += (size_t)d.B * d.nu;
    if (K0)
      memcpy(K0, h + (size_t)first * K, (size_t)n * K * sizeof(double));
    return SMPC_OK;
  }
  if (x1)
    CU(cudaMemcpy2DAsync(x1, d.nx * sizeof(double), c->A.xs + ((size_t)first * N1 + 1) * d.nx, N1 * d.nx * sizeof(double), d.nx * sizeof(double), n, cudaMemcpyDeviceToHost, s));
  if (u0)
    CU(cudaMemcpy2DAsync(u0, d.nu * sizeof(double), c->A.us + (size_t)first * d.N * d.nu, (size_t)d.N * d.nu * sizeof(double), d.nu * sizeof(double), n, cudaMemcpyDeviceToHost, s));
  if (u1)
    CU(cudaMemcpy2DAsync(u1, d.nu * sizeof(double), c->A.us + ((size_t)first * d.N + 1) * d.nu, (size_t)d.N * d.nu * sizeof(double), d.nu * sizeof(double), n, cudaMemcpyDeviceToHost, s));
  if (K0)
    CU(cudaMemcpy2DAsync(
      K0, (size_t)d.nu * d.ndx * sizeof(double), c->A.sol + (size_t)first * N1 * c->A.SL.size + c->A.SL.oK,
      N1 * c->A.SL.size * sizeof(double), (size_t)d.nu * d.ndx * sizeof(double), n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SMPC_OK;
}

int smpc_get_info(smpc_ctx * c, int first, int n, smpc_solve_info * info)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpy2DAsync(
    info, sizeof(smpc_solve_info), reinterpret_cast<const char *>(c->A.inst + first) + offsetof(InstState, info), sizeof(InstState),
    sizeof(smpc_solve_info), n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}

int smpc_get_ncmax(smpc_ctx * c) { return c ? c->A.dm.ncmax : SMPC_ERR_INVALID; }

int smpc_get_duals(smpc_ctx * c, int first, int n, double * vs, double * lams)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  const Dims & d = c->A.dm;
  const size_t N1 = d.N + 1;
  int rc;
  if (vs)
    if ((rc = d2h(c, vs, c->A.vs + (size_t)first * N1 * d.ncmax, (size_t)n * N1 * d.ncmax * sizeof(double))) != SMPC_OK)
      return rc;
  if (lams)
    if ((rc = d2h(c, lams, c->A.lams + (size_t)first * N1 * d.ndx, (size_t)n * N1 * d.ndx * sizeof(double))) != SMPC_OK)
      return rc;
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}

int smpc_get_xdot(smpc_ctx * c, int first, int n, int t, double * xdot)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  const Dims & d = c->A.dm;
  if (t < 0 || t >= d.N)
  {
    smpc::err() = "Stage index exceeds stage vector size";
    return SMPC_ERR_INVALID;
  }
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpy2DAsync(
    xdot, d.ndx * sizeof(double), c->A.xdot + ((size_t)first * d.N + t) * d.ndx, (size_t)d.N * d.ndx * sizeof(double),
    d.ndx * sizeof(double), n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}

int smpc_get_contact_forces(smpc_ctx * c, int first, int n, int t, double * forces)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  if (!c->A.fc)
  {
    // reference include/simple-mpc/mpc.hpp:174-178: only defined for the full-dynamics handler
    smpc::err() = "getContactForces is only available for full-dynamics problems";
    return SMPC_ERR_UNSUPPORTED;
  }
  const Dims & d = c->A.dm;
  if (t < 0 || t >= d.N || !forces)
  {
    smpc::err() = "smpc_get_contact_forces: stage index exceeds the horizon, or null output";
    return SMPC_ERR_INVALID;
  }
  if (n == 0)
    return SMPC_OK;
  CU(cudaSetDevice(c->device));
  const size_t w = (size_t)d.nk * 3 * sizeof(double);
  CU(cudaMemcpy2DAsync(forces, w, c->A.fc + ((size_t)first * d.N + t) * d.nk * 3, (size_t)d.N * w, w, n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}

// x0 <- xs[1] of every instance (closed loop without a plant: the predicted next state is the measurement)
int smpc_x0_from_prediction(smpc_ctx * c)
{
  if (!c)
    return SMPC_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  const Dims & d = c->A.dm;
  CU(cudaMemcpy2DAsync(
    c->A.x0, d.nx * sizeof(double), c->A.xs + d.nx, (size_t)(d.N + 1) * d.nx * sizeof(double), d.nx * sizeof(double), d.B,
    cudaMemcpyDeviceToDevice, c->stream));
  return SMPC_OK;
}

// ------------------------------------------------------------------------------------------------
// batched MPC driver (mpc_driver.cuh)
// ------------------------------------------------------------------------------------------------
int smpc_mpc_setup(smpc_ctx * c, const smpc_mpc_settings * s, const int * phase, const double * x_reference)
{
  if (!c || !s || !x_reference)
  {
    smpc::err() = "smpc_mpc_setup: null argument";
    return SMPC_ERR_INVALID;
  }
  if (c->h_table.empty())
  {
    smpc::err() = "smpc_mpc_setup: no stage table uploaded";
    return SMPC_ERR_STATE;
  }
  const Dims & d = c->A.dm;
  const int ntable = (int)c->h_table.size();
  if (s->ncycle < 2 || s->ncycle > ntable || s->standing_id >= ntable || s->T_fly < 1 || s->T_contact < 0)
  {
    smpc::err() = "smpc_mpc_setup: ncycle / standing_id outside the stage table, or bad T_fly / T_contact";
    return SMPC_ERR_INVALID;
  }
  if (s->ncycle < d.N)
  {
    smpc::err() = "smpc_mpc_setup: the cycle must be at least as long as the horizon (tile it as generateCycleHorizon does)";
    return SMPC_ERR_INVALID;
  }
  CU(cudaSetDevice(c->device));
  int rc;
  if (!c->A.mpcS)
  {
    if ((rc = dev_alloc(c, &c->A.mpcS, 1)) != SMPC_OK || (rc = dev_alloc(c, &c->A.mpcI, (size_t)d.B)) != SMPC_OK
        || (rc = dev_alloc(c, &c->A.mpc_ids, (size_t)d.B)) != SMPC_OK
        || (rc = dev_alloc(c, &c->A.xmeas, (size_t)d.B * (c->hmodel.d.nq + c->hmodel.d.nv))) != SMPC_OK)
      return rc;
  }
  c->A.nxm = c->hmodel.d.nq + c->hmodel.d.nv;
  c->A.ncycle = s->ncycle;
  MpcShared & S = c->h_mpc;
  memset(&S, 0, sizeof(S));
  S.T_fly = s->T_fly;
  S.T_contact = s->T_contact;
  S.swing_apex = s->swing_apex;
  S.timestep = s->timestep;
  S.standing_id = s->standing_id;
  std::vector<unsigned> masks((size_t)s->ncycle);
  for (int k = 0; k < s->ncycle; k++)
    masks[(size_t)k] = c->h_table[(size_t)k].contact_mask;
  // one MpcInst per distinct phase, then replicated
  std::vector<MpcInst> per_phase((size_t)s->ncycle);
  std::vector<char> have((size_t)s->ncycle, 0);
  std::vector<MpcInst> h((size_t)d.B);
  for (int b = 0; b < d.B; b++)
  {
    const int p = phase ? phase[b] : 0;
    if (p < 0 || p >= s->ncycle)
    {
      smpc::err() = "smpc_mpc_setup: phase outside [0, ncycle)";
      return SMPC_ERR_INVALID;
    }
    if (!have[(size_t)p])
    {
      MpcInst & I = per_phase[(size_t)p];
      memset(&I, 0, sizeof(I));
      mpc_init_queues(masks.data(), s->ncycle, d.nk, d.N, p, I);
      if (I.overflow)
      {
        smpc::err() = "smpc_mpc_setup: more than SMPC_MPC_QUEUE contact switches of one foot in the cycle";
        return SMPC_ERR_UNSUPPORTED;
      }
      I.walking = 1;
      for (int k = 0; k < d.nx; k++)
        I.x_reference[k] = x_reference[k];
      have[(size_t)p] = 1;
    }
    h[(size_t)b] = per_phase[(size_t)p];
  }
  CU(cudaMemcpyAsync(c->A.mpcS, &S, sizeof(S), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->A.mpcI, h.data(), h.size() * sizeof(MpcInst), cudaMemcpyHostToDevice, c->stream));
  // starting foot poses and com0 from the forward kinematics of the reference configuration
  k_mpc_init<<<(d.B + 63) / 64, 64, 0, c->stream>>>(c->A);
  c->launches++;
  CU(cudaMemsetAsync(c->A.counters + 2, 0, sizeof(int), c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  c->mpc_ready = true;
  return SMPC_OK;
}

static int mpc_check(smpc_ctx * c, int first, int n, const char * what)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  if (!c->mpc_ready)
  {
    smpc::err() = std::string(what) + ": smpc_mpc_setup has not been called";
    return SMPC_ERR_STATE;
  }
  return SMPC_OK;
}

int smpc_mpc_set_velocity_base(smpc_ctx * c, int first, int n, const double * vb)
{
  int rc = mpc_check(c, first, n, "smpc_mpc_set_velocity_base");
  if (rc != SMPC_OK)
    return rc;
  if (!vb)
  {
    smpc::err() = "smpc_mpc_set_velocity_base: null argument";
    return SMPC_ERR_INVALID;
  }
  if (n == 0)
    return SMPC_OK;
  CU(cudaSetDevice(c->device));
  if ((rc = ensure_staging(c, (size_t)n * 6 * sizeof(double))) != SMPC_OK)
    return rc;
  CU(cudaMemcpyAsync(c->staging, vb, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  k_mpc_set_vb<<<(n + 127) / 128, 128, 0, c->stream>>>(c->A, first, n, (const double *)c->staging, -1);
  c->launches++;
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}

int smpc_mpc_set_walking(smpc_ctx * c, int first, int n, int walking)
{
  int rc = mpc_check(c, first, n, "smpc_mpc_set_walking");
  if (rc != SMPC_OK)
    return rc;
  if (n == 0)
    return SMPC_OK;
  CU(cudaSetDevice(c->device));
  k_mpc_set_vb<<<(n + 127) / 128, 128, 0, c->stream>>>(c->A, first, n, nullptr, walking ? 1 : 0);
  c->launches++;
  CU(cudaGetLastError());
  return SMPC_OK;
}

// The device-side part of MPC::iterate up to the solve (reference src/mpc.cpp:189-209), enqueued on the context's stream
static int enqueue_tick_driver(smpc_ctx * c)
{
  const Dims & d = c->A.dm;
  cudaStream_t s = c->stream;
  k_mpc_recede<<<(d.B + 127) / 128, 128, 0, s>>>(c->A);
  k_cycle<<<d.B, 64, 0, s>>>(c->A, 0, nullptr, c->A.mpc_ids);
  k_mpc_refs<<<d.B, 64, 0, s>>>(c->A);
  k_shift<<<d.B, 64, 0, s>>>(c->A, 0);
  c->launches += 4;
  CU(cudaMemcpyAsync(c->h_counters + 2, c->A.counters + 2, sizeof(int), cudaMemcpyDeviceToHost, s));
  return SMPC_OK;
}

static void drop_tick_graph(smpc_ctx * c)
{
  if (c->tick_graph)
    cudaGraphExecDestroy(c->tick_graph);
  c->tick_graph = nullptr;
}

// Capture the normal tick - driver kernels, k_run_begin, one round, its first linesearch trial, k_finish and the two
// read-backs of counters - into one executable graph. The kernels take their arguments (the Arrays view) by value, so the
// graph is re-captured whenever that view changes (stage table re-upload, smpc_mpc_setup)
static int capture_tick_graph(smpc_ctx * c, int max_iters)
{
  drop_tick_graph(c);
  const Dims & d = c->A.dm;
  cudaStream_t s = c->stream;
  if (!c->h_head && d.B <= 256)
    CU(cudaMallocHost(&c->h_head, (size_t)d.B * (d.nx + 2 * d.nu + (size_t)d.nu * d.ndx) * sizeof(double)));
  const int64_t launches0 = c->launches;
  cudaGraph_t g = nullptr;
  CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int rc = enqueue_tick_driver(c);
  if (rc == SMPC_OK)
  {
    k_run_begin<<<d.B, 128, 0, s>>>(c->A, max_iters);
    c->launches++;
    rc = enqueue_round_head(c, false);
  }
  if (rc == SMPC_OK)
    rc = enqueue_trial(c, false);
  if (rc == SMPC_OK && c->h_head)
  {
    const size_t N1 = d.N + 1, K = (size_t)d.nu * d.ndx;
    double * h = c->h_head;
    cudaMemcpy2DAsync(h, d.nx * sizeof(double), c->A.xs + d.nx, N1 * d.nx * sizeof(double), d.nx * sizeof(double), d.B, cudaMemcpyDeviceToHost, s);
    h += (size_t)d.B * d.nx;
    cudaMemcpy2DAsync(h, d.nu * sizeof(double), c->A.us, (size_t)d.N * d.nu * sizeof(double), d.nu * sizeof(double), d.B, cudaMemcpyDeviceToHost, s);
    h += (size_t)d.B * d.nu;
    cudaMemcpy2DAsync(h, d.nu * sizeof(double), c->A.us + d.nu, (size_t)d.N * d.nu * sizeof(double), d.nu * sizeof(double), d.B, cudaMemcpyDeviceToHost, s);
    h += (size_t)d.B * d.nu;
    cudaMemcpy2DAsync(h, K * sizeof(double), c->A.sol + c->A.SL.oK, N1 * c->A.SL.size * sizeof(double), K * sizeof(double), d.B, cudaMemcpyDeviceToHost, s);
  }
  cudaError_t e = cudaStreamEndCapture(s, &g);
  c->launches = launches0; // nothing ran
  if (rc != SMPC_OK || e != cudaSuccess)
  {
    if (g)
      cudaGraphDestroy(g);
    if (rc == SMPC_OK)
    {
      smpc::err() = std::string("smpc_mpc_iterate: stream capture failed: ") + cudaGetErrorString(e);
      rc = SMPC_ERR_CUDA;
    }
    return rc;
  }
  e = cudaGraphInstantiate(&c->tick_graph, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess)
  {
    c->tick_graph = nullptr;
    smpc::err() = std::string("smpc_mpc_iterate: cudaGraphInstantiate: ") + cudaGetErrorString(e);
    return SMPC_ERR_CUDA;
  }
  c->tick_graph_A = c->A;
  c->tick_graph_iters = max_iters;
  c->tick_graph_fwd = c->forward_staged ? 1 : 0;
  return SMPC_OK;
}

int smpc_mpc_use_graph(smpc_ctx * c, int on)
{
  if (!c)
  {
    smpc::err() = "smpc_mpc_use_graph: null context";
    return SMPC_ERR_INVALID;
  }
  c->use_graph = on ? 1 : 0;
  if (!on)
    drop_tick_graph(c);
  return SMPC_OK;
}

int smpc_mpc_iterate(smpc_ctx * c, const double * x_meas, int max_iters)
{
  int rc = mpc_check(c, 0, 0, "smpc_mpc_iterate");
  if (rc != SMPC_OK)
    return rc;
  CU(cudaSetDevice(c->device));
  const Dims & d = c->A.dm;
  cudaStream_t s = c->stream;
  if (x_meas)
    CU(cudaMemcpyAsync(c->A.xmeas, x_meas, (size_t)d.B * c->A.nxm * sizeof(double), cudaMemcpyHostToDevice, s));
  c->head_cached = false;
  if (c->use_graph)
  {
    // one graph launch and one host synchronisation per tick; a linesearch that backtracks or an instance that needs
    // another round (both rare in MPC mode, profiles/r02m) continues on the stream below
    if (!c->tick_graph || c->tick_graph_iters != max_iters || c->tick_graph_fwd != (c->forward_staged ? 1 : 0) ||
        memcmp(&c->tick_graph_A, &c->A, sizeof(c->A)) != 0)
      if ((rc = capture_tick_graph(c, max_iters)) != SMPC_OK)
        return rc;
    CU(cudaEventRecord(c->ev0, s));
    CU(cudaGraphLaunch(c->tick_graph, s));
    CU(cudaEventRecord(c->ev1, s));
    CU(cudaStreamSynchronize(s));
    c->launches += 12;
    c->graph_ticks++;
    c->head_cached = c->h_head && c->h_counters[0] == 0 && c->h_counters[1] == 0;
    if ((rc = solve_rounds(c, max_iters, true)) != SMPC_OK)
      return rc;
    CU(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1)); // the per-kernel timers only see ticks outside the graph
  }
  else
  {
    if ((rc = enqueue_tick_driver(c)) != SMPC_OK)
      return rc;
    rc = smpc_solve(c, nullptr, max_iters); // synchronises the stream
    if (rc != SMPC_OK)
      return rc;
  }
  if (c->h_counters[2] != 0)
  {
    smpc::err() = "smpc_mpc_iterate: a take-off / landing queue overflowed (more than SMPC_MPC_QUEUE pending events of one foot)";
    return SMPC_ERR_UNSUPPORTED;
  }
  return SMPC_OK;
}

int smpc_mpc_xmeas_from_prediction(smpc_ctx * c)
{
  int rc = mpc_check(c, 0, 0, "smpc_mpc_xmeas_from_prediction");
  if (rc != SMPC_OK)
    return rc;
  const Dims & d = c->A.dm;
  if (d.nx != c->A.nxm)
  {
    smpc::err() = "smpc_mpc_xmeas_from_prediction: the problem state is not the multibody state (centroidal OCP)";
    return SMPC_ERR_UNSUPPORTED;
  }
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpy2DAsync(
    c->A.xmeas, d.nx * sizeof(double), c->A.xs + d.nx, (size_t)(d.N + 1) * d.nx * sizeof(double), d.nx * sizeof(double), d.B,
    cudaMemcpyDeviceToDevice, c->stream));
  return SMPC_OK;
}

int smpc_mpc_get_timings(smpc_ctx * c, int first, int n, int * takeoff, int * land)
{
  int rc = mpc_check(c, first, n, "smpc_mpc_get_timings");
  if (rc != SMPC_OK)
    return rc;
  if (n == 0)
    return SMPC_OK;
  CU(cudaSetDevice(c->device));
  std::vector<MpcInst> h((size_t)n);
  CU(cudaMemcpyAsync(h.data(), c->A.mpcI + first, (size_t)n * sizeof(MpcInst), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  const int nk = c->A.dm.nk;
  for (int i = 0; i < n; i++)
    for (int f = 0; f < nk; f++)
    {
      if (takeoff)
        takeoff[i * nk + f] = h[(size_t)i].ntake[f] > 0 ? h[(size_t)i].take[f][0] : -1;
      if (land)
        land[i * nk + f] = h[(size_t)i].nland[f] > 0 ? h[(size_t)i].land[f][0] : -1;
    }
  return SMPC_OK;
}

int smpc_get_slots(smpc_ctx * c, int first, int n, smpc_slot * slots, smpc_terminal * term)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  if (n == 0)
    return SMPC_OK;
  CU(cudaSetDevice(c->device));
  const Dims & d = c->A.dm;
  if (slots)
  {
    const size_t bytes = (size_t)n * d.N * sizeof(smpc_slot);
    int rc;
    if ((rc = ensure_staging(c, bytes)) != SMPC_OK)
      return rc;
    k_load_slots<<<n * d.N, 64, 0, c->stream>>>(c->A, first, (smpc_slot *)c->staging);
    c->launches++;
    CU(cudaMemcpyAsync(slots, c->staging, bytes, cudaMemcpyDeviceToHost, c->stream));
  }
  if (term)
    CU(cudaMemcpyAsync(term, c->A.term + first, (size_t)n * sizeof(smpc_terminal), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}

int smpc_feedback_control(smpc_ctx * c, int first, int n, const double * x_meas, double * u)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  if (!x_meas || !u)
  {
    smpc::err() = "smpc_feedback_control: null argument";
    return SMPC_ERR_INVALID;
  }
  if (n == 0)
    return SMPC_OK;
  CU(cudaSetDevice(c->device));
  const Dims & d = c->A.dm;
  const size_t in_b = (size_t)n * d.nx * sizeof(double), out_b = (size_t)n * d.nu * sizeof(double);
  int rc;
  if ((rc = ensure_staging(c, in_b + out_b)) != SMPC_OK)
    return rc;
  double * din = (double *)c->staging;
  double * dout = din + (size_t)n * d.nx;
  CU(cudaMemcpyAsync(din, x_meas, in_b, cudaMemcpyHostToDevice, c->stream));
  k_feedback<<<n, 32, 0, c->stream>>>(c->A, first, din, dout);
  c->launches++;
  CU(cudaMemcpyAsync(u, dout, out_b, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return SMPC_OK;
}

int smpc_interpolate_plan(smpc_ctx * c, int first, int n, double delay, double * x_interp, double * u_interp, int * contacts)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  if (!(delay >= 0))
  {
    smpc::err() = "smpc_interpolate_plan: negative delay";
    return SMPC_ERR_INVALID;
  }
  if (n == 0)
    return SMPC_OK;
  CU(cudaSetDevice(c->device));
  const Dims & d = c->A.dm;
  const double timestep = c->hocp.dt;
  // reference src/interpolator.cpp:14-15
  const size_t step_nb = static_cast<size_t>(delay / timestep);
  const double step_progress = (delay - (double)step_nb * timestep) / timestep;
  const int step = step_nb > (size_t)d.N ? d.N : (int)step_nb;
  const size_t xb = (size_t)n * d.nx * sizeof(double), ub = (size_t)n * d.nu * sizeof(double), cb = (size_t)n * d.nk * sizeof(int);
  int rc;
  if ((rc = ensure_staging(c, xb + ub + cb)) != SMPC_OK)
    return rc;
  double * dx = (double *)c->staging;
  double * du = dx + (size_t)n * d.nx;
  int * dc = (int *)(du + (size_t)n * d.nu);
  k_interpolate<<<n, 32, 0, c->stream>>>(c->A, first, step, step_progress, x_interp ? dx : nullptr, u_interp ? du : nullptr,
                                        contacts ? dc : nullptr);
  c->launches++;
  if (x_interp)
    CU(cudaMemcpyAsync(x_interp, dx, xb, cudaMemcpyDeviceToHost, c->stream));
  if (u_interp)
    CU(cudaMemcpyAsync(u_interp, du, ub, cudaMemcpyDeviceToHost, c->stream));
  if (contacts)
    CU(cudaMemcpyAsync(contacts, dc, cb, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return SMPC_OK;
}

// device timer on the context's stream: which = 0 start, 1 stop
int smpc_timer_record(smpc_ctx * c, int which)
{
  if (!c)
    return SMPC_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  CU(cudaEventRecord(which ? c->evu1 : c->evu0, c->stream));
  return SMPC_OK;
}
int smpc_timer_elapsed_ms(smpc_ctx * c, float * ms)
{
  if (!c || !ms)
    return SMPC_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  CU(cudaEventSynchronize(c->evu1));
  CU(cudaEventElapsedTime(ms, c->evu0, c->evu1));
  return SMPC_OK;
}
// accumulated device time of the Riccati kernel / whole solves since the last reset
int smpc_timer_stats(smpc_ctx * c, int reset, double * riccati_ms, int64_t * riccati_launches, double * solve_ms, int64_t * solves)
{
  if (!c)
    return SMPC_ERR_INVALID;
  if (riccati_ms)
    *riccati_ms = c->acc_ric_ms;
  if (riccati_launches)
    *riccati_launches = c->acc_ric_launches;
  if (solve_ms)
    *solve_ms = c->acc_solve_ms;
  if (solves)
    *solves = c->acc_solves;
  if (reset)
  {
    c->acc_ric_ms = c->acc_solve_ms = 0;
    c->acc_ric_launches = c->acc_solves = 0;
    for (double & v : c->acc_kernel_ms)
      v = 0;
  }
  return SMPC_OK;
}
// accumulated device time per kernel of the FIRST round of every solve since the last smpc_timer_stats(reset):
// [0] linearisation (k_knots), [1] Riccati backward sweep, [2] forward sweep, [3] first trial evaluation + linesearch decision
int smpc_kernel_times(smpc_ctx * c, double * ms4)
{
  if (!c || !ms4)
    return SMPC_ERR_INVALID;
  for (int k = 0; k < 4; k++)
    ms4[k] = c->acc_kernel_ms[k];
  return SMPC_OK;
}
int smpc_sync(smpc_ctx * c)
{
  if (!c)
    return SMPC_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}

int smpc_last_solve_ms(smpc_ctx * c, float * total_ms, float * riccati_ms)
{
  if (!c)
    return SMPC_ERR_INVALID;
  if (total_ms)
    *total_ms = c->last_ms;
  if (riccati_ms)
    *riccati_ms = c->last_ric_ms;
  return SMPC_OK;
}
int64_t smpc_launch_count(smpc_ctx * c) { return c ? c->launches : 0; }
const char * smpc_riccati_variant(smpc_ctx * c) { return c ? c->riccati_variant : ""; }
int smpc_riccati_legs(smpc_ctx * c) { return c ? (c->A.legs > 1 ? c->A.legs : 1) : SMPC_ERR_INVALID; }

// LQ knot t of instance `inst` as assembled by the last linearisation (test hook; device layout -> dense blocks)
int smpc_get_knot(smpc_ctx * c, int inst, int t, int * nc, double * knot_out /* KnotLayout.size doubles */)
{
  if (check_range(c, inst, 1))
    return SMPC_ERR_INVALID;
  const Dims & d = c->A.dm;
  if (t < 0 || t > d.N)
  {
    smpc::err() = "bad knot index";
    return SMPC_ERR_INVALID;
  }
  CU(cudaSetDevice(c->device));
  const size_t k = (size_t)inst * (d.N + 1) + t;
  CU(cudaMemcpyAsync(knot_out, c->A.knots + k * c->A.L.size, c->A.L.size * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (nc)
    CU(cudaMemcpyAsync(nc, c->A.ncs + k, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SMPC_OK;
}
int smpc_knot_layout(smpc_ctx * c, int * out /*12*/)
{
  if (!c)
    return SMPC_ERR_INVALID;
  const KnotLayout & L = c->A.L;
  const int v[12] = {L.oQ, L.oS, L.oR, L.oq, L.or_, L.oA, L.oB, L.of, L.oC, L.oD, L.od, L.size};
  memcpy(out, v, sizeof(v));
  return SMPC_OK;
}

} // extern "C"
