// ORACLE — test infrastructure only (see orc_math.hpp header). PARITY UNPINNED.
//
// orc_api.cpp — C entry points of the CPU oracle. The solver-level ones mirror include/smpc.h
// one to one (orc_X has the signature of smpc_X without the device argument) so that tests drive the
// oracle and the CUDA library with the same call sequence. The unit-level ones expose single
// functions for finite-difference and known-answer checks.
#include "orc_solver.hpp"
#include <cstdio>
#include <memory>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

struct OrcMpcDriver; // orc_mpc_* below (timed CPU baseline of the whole tick)
struct orc_ctx
{
  int batch;
  std::vector<std::unique_ptr<Solver>> inst;
  std::vector<smpc_slot> table;
  std::shared_ptr<OrcMpcDriver> mpc;
};

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH                                                                                                      \
  }                                                                                                                    \
  catch (const std::exception & e)                                                                                     \
  {                                                                                                                    \
    g_err = e.what();                                                                                                  \
    return SMPC_ERR_INVALID;                                                                                           \
  }                                                                                                                    \
  return SMPC_OK;

static int check_range(orc_ctx * c, int first, int n)
{
  if (!c || first < 0 || n < 0 || first + n > c->batch)
  {
    g_err = "instance range out of bounds";
    return SMPC_ERR_INVALID;
  }
  return 0;
}

extern "C" {

const char * orc_last_error(void) { return g_err.c_str(); }

void orc_default_solver_opts(smpc_solver_opts * o)
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

int orc_problem_dims(const smpc_model_desc * m, const smpc_ocp_desc * o, int * nx, int * ndx, int * nu)
{
  ORC_TRY
  smpc_ocp_desc oc = *o;
  int fs = oc.force_size;
  if (oc.kind == SMPC_OCP_KINODYNAMICS)
  {
    *nx = m->nq + m->nv;
    *ndx = 2 * m->nv;
    *nu = m->nv - 6 + fs * m->nfeet;
  }
  else if (oc.kind == SMPC_OCP_FULLDYNAMICS)
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
  ORC_CATCH
}

int orc_create(orc_ctx ** out, int batch, const smpc_model_desc * model, const smpc_ocp_desc * ocp, const smpc_solver_opts * opts)
{
  ORC_TRY
  if (batch < 1)
    throw std::runtime_error("batch must be >= 1");
  auto c = new orc_ctx;
  c->batch = batch;
  for (int i = 0; i < batch; i++)
    c->inst.emplace_back(new Solver(*model, *ocp, *opts));
  *out = c;
  ORC_CATCH
}
void orc_destroy(orc_ctx * c) { delete c; }

int orc_set_slots(orc_ctx * c, int first, int n, const smpc_slot * slots)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
  {
    Solver & s = *c->inst[first + i];
    for (int t = 0; t < s.N; t++)
      s.slots[t] = slots[(size_t)i * s.N + t];
  }
  return SMPC_OK;
}
int orc_set_slot(orc_ctx * c, int first, int n, int t, const smpc_slot * slots)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  if (t < 0 || t >= c->inst[0]->N)
  {
    g_err = "Stage index exceeds stage vector size";
    return SMPC_ERR_INVALID;
  }
  for (int i = 0; i < n; i++)
    c->inst[first + i]->slots[t] = slots[i];
  return SMPC_OK;
}
int orc_set_foot_refs(orc_ctx * c, int first, int n, const double * foot_ref)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
  {
    Solver & s = *c->inst[first + i];
    const int nk = s.P.nk;
    for (int t = 0; t < s.N; t++)
      for (int f = 0; f < nk; f++)
        memcpy(s.slots[t].foot_ref[f], foot_ref + (((size_t)i * s.N + t) * nk + f) * 12, 12 * sizeof(double));
  }
  return SMPC_OK;
}
int orc_set_terminal(orc_ctx * c, int first, int n, const smpc_terminal * term)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
    c->inst[first + i]->term = term[i];
  return SMPC_OK;
}
int orc_cycle(orc_ctx * c, int first, int n, const smpc_slot * tail)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
    c->inst[first + i]->cycle(tail[i]);
  return SMPC_OK;
}
int orc_upload_stage_table(orc_ctx * c, const smpc_slot * stages, int n)
{
  if (!c || !stages || n < 1)
  {
    g_err = "orc_upload_stage_table: bad arguments";
    return SMPC_ERR_INVALID;
  }
  c->table.assign(stages, stages + n);
  return SMPC_OK;
}
int orc_cycle_ids(orc_ctx * c, int first, int n, const int * ids)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
  {
    if (ids[i] < 0 || ids[i] >= (int)c->table.size())
    {
      g_err = "orc_cycle_ids: stage id out of range";
      return SMPC_ERR_INVALID;
    }
    c->inst[first + i]->cycle(c->table[ids[i]]);
  }
  return SMPC_OK;
}
int orc_set_warmstart(orc_ctx * c, int first, int n, const double * xs, const double * us)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
  {
    Solver & s = *c->inst[first + i];
    for (int t = 0; t <= s.N; t++)
      s.xs[t].assign(xs + ((size_t)i * (s.N + 1) + t) * s.nx, xs + ((size_t)i * (s.N + 1) + t + 1) * s.nx);
    for (int t = 0; t < s.N; t++)
      s.us[t].assign(us + ((size_t)i * s.N + t) * s.nu, us + ((size_t)i * s.N + t + 1) * s.nu);
  }
  return SMPC_OK;
}
// reference src/mpc.cpp:201-207: xs.erase(begin); xs[0] = x0 (done by solve); xs.push_back(xs.back()); same for us
int orc_shift_warmstart(orc_ctx * c, int first, int n)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
  {
    Solver & s = *c->inst[first + i];
    std::rotate(s.xs.begin(), s.xs.begin() + 1, s.xs.end());
    s.xs[s.N] = s.xs[s.N - 1];
    std::rotate(s.us.begin(), s.us.begin() + 1, s.us.end());
    s.us[s.N - 1] = s.us[s.N - 2];
  }
  return SMPC_OK;
}
int orc_solve(orc_ctx * c, const double * x0, int max_iters)
{
  ORC_TRY
  std::string err;
  // one instance: no team over instances, so that Solver::knot_threads can open one over the knots
#pragma omp parallel for schedule(dynamic, 1) if (c->batch > 1)
  for (int i = 0; i < c->batch; i++)
  {
    try
    {
      Solver & s = *c->inst[i];
      s.run(x0 + (size_t)i * s.nx, max_iters);
    }
    catch (const std::exception & e)
    {
#pragma omp critical
      err = e.what();
    }
  }
  if (!err.empty())
    throw std::runtime_error(err);
  ORC_CATCH
}
int orc_get_results(orc_ctx * c, int first, int n, double * xs, double * us, double * Ks, double * ks)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
  {
    Solver & s = *c->inst[first + i];
    const int N = s.N, nx = s.nx, nu = s.nu, ndx = s.ndx;
    if (xs)
      for (int t = 0; t <= N; t++)
        memcpy(xs + ((size_t)i * (N + 1) + t) * nx, s.xs[t].data(), nx * sizeof(double));
    if (us)
      for (int t = 0; t < N; t++)
        memcpy(us + ((size_t)i * N + t) * nu, s.us[t].data(), nu * sizeof(double));
    if ((Ks || ks) && (int)s.sol.size() == N + 1)
      for (int t = 0; t < N; t++)
      {
        if (Ks)
          memcpy(Ks + ((size_t)i * N + t) * nu * ndx, s.sol[t].K.d.data(), (size_t)nu * ndx * sizeof(double));
        if (ks)
          memcpy(ks + ((size_t)i * N + t) * nu, s.sol[t].k.data(), nu * sizeof(double));
      }
  }
  return SMPC_OK;
}
int orc_get_head(orc_ctx * c, int first, int n, double * x1, double * u0, double * u1, double * K0)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
  {
    Solver & s = *c->inst[first + i];
    if (x1)
      memcpy(x1 + (size_t)i * s.nx, s.xs[1].data(), s.nx * sizeof(double));
    if (u0)
      memcpy(u0 + (size_t)i * s.nu, s.us[0].data(), s.nu * sizeof(double));
    if (u1)
      memcpy(u1 + (size_t)i * s.nu, s.us[1].data(), s.nu * sizeof(double));
    if (K0 && !s.sol.empty())
      memcpy(K0 + (size_t)i * s.nu * s.ndx, s.sol[0].K.d.data(), (size_t)s.nu * s.ndx * sizeof(double));
  }
  return SMPC_OK;
}
int orc_get_info(orc_ctx * c, int first, int n, smpc_solve_info * info)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
    info[i] = c->inst[first + i]->info;
  return SMPC_OK;
}
int orc_get_ncmax(orc_ctx * c) { return c->inst[0]->ncmax; }
int orc_get_duals(orc_ctx * c, int first, int n, double * vs, double * lams)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
  {
    Solver & s = *c->inst[first + i];
    const int N = s.N;
    if (vs)
      for (int t = 0; t <= N; t++)
        memcpy(vs + ((size_t)i * (N + 1) + t) * s.ncmax, s.vs[t].data(), s.ncmax * sizeof(double));
    if (lams)
      for (int t = 0; t <= N; t++)
        memcpy(lams + ((size_t)i * (N + 1) + t) * s.ndx, s.lams[t].data(), s.ndx * sizeof(double));
  }
  return SMPC_OK;
}
int orc_get_xdot(orc_ctx * c, int first, int n, int t, double * xdot)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  if (t < 0 || t >= c->inst[0]->N)
  {
    g_err = "Stage index exceeds stage vector size";
    return SMPC_ERR_INVALID;
  }
  for (int i = 0; i < n; i++)
  {
    Solver & s = *c->inst[first + i];
    memcpy(xdot + (size_t)i * s.ndx, s.xdots[t].data(), s.ndx * sizeof(double));
  }
  return SMPC_OK;
}
// MPC::getContactForces(t) (reference src/mpc.cpp:354-380): [n][nfeet][3], zero rows for swing feet
int orc_get_contact_forces(orc_ctx * c, int first, int n, int t, double * forces)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
  {
    Solver & s = *c->inst[first + i];
    if (s.P.ocp.kind != SMPC_OCP_FULLDYNAMICS)
    {
      g_err = "getContactForces is only available for full-dynamics problems";
      return SMPC_ERR_UNSUPPORTED;
    }
    if (t < 0 || t >= s.N || s.fcs[(size_t)t].size() != (size_t)s.P.nk * 3)
    {
      g_err = "orc_get_contact_forces: stage index exceeds the horizon (or no evaluation yet)";
      return SMPC_ERR_INVALID;
    }
    memcpy(forces + (size_t)i * s.P.nk * 3, s.fcs[(size_t)t].data(), (size_t)s.P.nk * 3 * sizeof(double));
  }
  return SMPC_OK;
}
// u = us[0] - Ks[0] * difference(x_meas, xs[0])   (reference examples/talos_fulldynamics.py:203-205;
// RobotModelHandler::difference(x1, x2) = x2 (-) x1, reference src/robot-handler.cpp:81-96)
int orc_feedback_control(orc_ctx * c, int first, int n, const double * x_meas, double * u)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
  {
    Solver & s = *c->inst[first + i];
    std::vector<double> dx((size_t)s.ndx);
    state_difference(s.P, x_meas + (size_t)i * s.nx, s.xs[0].data(), dx.data());
    for (int a = 0; a < s.nu; a++)
    {
      double acc = 0;
      for (int k = 0; k < s.ndx; k++)
        acc = std::fma(s.sol[0].K.d[(size_t)a * s.ndx + k], dx[(size_t)k], acc);
      u[(size_t)i * s.nu + a] = s.us[0][a] - acc;
    }
  }
  return SMPC_OK;
}
// Interpolator::interpolateState / interpolateLinear / interpolateContacts on the plan (reference src/interpolator.cpp:5-78)
int orc_interpolate_plan(orc_ctx * c, int first, int n, double delay, double * x_interp, double * u_interp, int * contacts)
{
  if (check_range(c, first, n))
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
  {
    Solver & s = *c->inst[first + i];
    const double timestep = s.P.ocp.timestep;
    size_t step_nb = static_cast<size_t>(delay / timestep);
    const double step_progress = (delay - (double)step_nb * timestep) / timestep;
    if (x_interp)
    {
      double * o = x_interp + (size_t)i * s.nx;
      if (step_nb >= s.xs.size() - 1)
        memcpy(o, s.xs.back().data(), s.nx * sizeof(double));
      else if (s.P.ocp.kind == SMPC_OCP_CENTROIDAL)
        for (int k = 0; k < s.nx; k++)
          o[k] = s.xs[step_nb + 1][k] * step_progress + s.xs[step_nb][k] * (1. - step_progress);
      else
      {
        const Model & m = s.P.model;
        std::vector<double> dq((size_t)m.nv);
        difference_q(m, s.xs[step_nb].data(), s.xs[step_nb + 1].data(), dq.data());
        for (auto & v : dq)
          v = step_progress * v;
        integrate_q(m, s.xs[step_nb].data(), dq.data(), o);
        for (int k = 0; k < m.nv; k++)
          o[m.nq + k] = s.xs[step_nb + 1][m.nq + k] * step_progress + s.xs[step_nb][m.nq + k] * (1. - step_progress);
      }
    }
    if (u_interp)
    {
      double * o = u_interp + (size_t)i * s.nu;
      if (step_nb >= s.us.size() - 1)
        memcpy(o, s.us.back().data(), s.nu * sizeof(double));
      else
        for (int k = 0; k < s.nu; k++)
          o[k] = s.us[step_nb + 1][k] * step_progress + s.us[step_nb][k] * (1. - step_progress);
    }
    if (contacts)
    {
      const size_t sc = std::min(step_nb, (size_t)s.N - 1);
      for (int f = 0; f < s.P.nk; f++)
        contacts[(size_t)i * s.P.nk + f] = (s.slots[sc].contact_mask >> f) & 1u;
    }
  }
  return SMPC_OK;
}
// OpenMP threads over the knots of each instance (aligator's num_threads); meant for batch == 1
int orc_set_knot_threads(orc_ctx * c, int n)
{
  if (!c || n < 1)
  {
    g_err = "orc_set_knot_threads: bad arguments";
    return SMPC_ERR_INVALID;
  }
  for (auto & s : c->inst)
    s->knot_threads = n;
  return SMPC_OK;
}
// seconds spent in {linearisation, Riccati backward, forward sweep, linesearch trials}, summed over instances
int orc_get_timers(orc_ctx * c, double * out4)
{
  if (!c || !out4)
    return SMPC_ERR_INVALID;
  for (int k = 0; k < 4; k++)
    out4[k] = 0;
  for (auto & s : c->inst)
    for (int k = 0; k < 4; k++)
      out4[k] += s->timers[k];
  return SMPC_OK;
}
// torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU-baseline leg of bench.py asks for the host cores explicitly
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0)
    omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------
// The whole MPC tick on the CPU, for the TIMED BASELINE only (bench.py cpu_baseline / --impl reference): the part
// of MPC::iterate around run() - updateInternalData, recedeWithCycle, updateStepTrackerReferences, warm-start
// shift (reference src/mpc.cpp:189-324) - per instance, so that the CPU arm does the same work per tick as
// smpc_mpc_iterate. This section is NOT a checker: it compiles the product's own driver routines
// (simple-mpc_b200/csrc/mpc_driver.cuh, written in the phase style of par.cuh that g++ serialises) instead of
// restating them; the checker of the device driver is the host MPC class (tests/test_mpc_driver_emul.py,
// tests/test_gpu_parity.py). Signatures mirror smpc_mpc_* (include/smpc.h).
// ---------------------------------------------------------------------------------------------------
#include "../simple-mpc_b200/csrc/mpc_driver.cuh"
#include "../simple-mpc_b200/csrc/host_build.h"

struct OrcMpcDriver
{
  smpc::DevModel M;
  smpc::DevOcp O;
  smpc::MpcShared S;
  std::vector<smpc::MpcInst> I;
  int ncycle = 0, nxm = 0;
};

extern "C" {
int orc_mpc_setup(orc_ctx * c, const smpc_mpc_settings * st, const int * phase, const double * x_reference)
{
  ORC_TRY
  if (!c || !st || !x_reference || c->table.empty())
    throw std::runtime_error("orc_mpc_setup: null argument or no stage table uploaded");
  auto d = std::make_shared<OrcMpcDriver>();
  Solver & s0 = *c->inst[0];
  smpc_ocp_desc od = s0.P.ocp; // weights are not needed by the driver
  smpc::build_dev_model(s0.P.model.d, d->M);
  memset(&d->O, 0, sizeof(d->O));
  d->O.terminal_constraint = od.terminal_constraint;
  d->O.kind = od.kind, d->O.N = s0.N, d->O.fs = s0.P.fs, d->O.nk = s0.P.nk, d->O.nx = s0.nx, d->O.ndx = s0.ndx, d->O.nu = s0.nu;
  d->O.dt = od.timestep, d->O.mass = s0.P.mass;
  memset(&d->S, 0, sizeof(d->S));
  d->S.T_fly = st->T_fly, d->S.T_contact = st->T_contact, d->S.swing_apex = st->swing_apex, d->S.timestep = st->timestep;
  d->S.standing_id = st->standing_id;
  d->ncycle = st->ncycle;
  d->nxm = s0.P.model.nq + s0.P.model.nv;
  std::vector<unsigned> masks((size_t)st->ncycle);
  for (int k = 0; k < st->ncycle; k++)
    masks[(size_t)k] = c->table[(size_t)k].contact_mask;
  d->I.resize((size_t)c->batch);
  smpc::MpcScratch scratch;
  for (int b = 0; b < c->batch; b++)
  {
    smpc::MpcInst & I = d->I[(size_t)b];
    memset(&I, 0, sizeof(I));
    smpc::mpc_init_queues(masks.data(), st->ncycle, d->O.nk, d->O.N, phase ? phase[b] : 0, I);
    I.walking = 1;
    for (int k = 0; k < s0.nx; k++)
      I.x_reference[k] = x_reference[k];
    smpc::mpc_init_instance(d->M, s0.P.model.d.q_ref, b == 0 ? &d->S : nullptr, I, scratch);
  }
  c->mpc = d;
  ORC_CATCH
}
int orc_mpc_set_velocity_base(orc_ctx * c, int first, int n, const double * vb)
{
  if (check_range(c, first, n) || !c->mpc)
    return SMPC_ERR_INVALID;
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 6; k++)
      c->mpc->I[(size_t)(first + i)].velocity_base[k] = vb[i * 6 + k];
  return SMPC_OK;
}
// x_meas [batch][nq + nv]: recede -> cycle -> references -> warm-start shift -> run(max_iters), OpenMP over instances
int orc_mpc_iterate(orc_ctx * c, const double * x_meas, int max_iters)
{
  ORC_TRY
  if (!c || !c->mpc || !x_meas)
    throw std::runtime_error("orc_mpc_iterate: orc_mpc_setup has not been called");
  OrcMpcDriver & D = *c->mpc;
  std::string err;
#pragma omp parallel for schedule(dynamic, 1) if (c->batch > 1)
  for (int b = 0; b < c->batch; b++)
  {
    try
    {
      Solver & s = *c->inst[(size_t)b];
      smpc::MpcInst & I = D.I[(size_t)b];
      const int id = smpc::mpc_recede(D.S, I, c->table.data(), D.ncycle, D.O.nk, s.N, s.slots[(size_t)s.N - 1].contact_mask);
      s.cycle(c->table[(size_t)id]);
      smpc::MpcRing ring;
      ring.base = s.slots.data();
      ring.head = 0;
      ring.N = s.N;
      smpc::MpcScratch scratch;
      std::vector<double> x0((size_t)s.nx);
      smpc::mpc_refs(D.M, D.O, D.S, I, x_meas + (size_t)b * D.nxm, ring, s.term, x0.data(), scratch);
      std::rotate(s.xs.begin(), s.xs.begin() + 1, s.xs.end());
      s.xs[(size_t)s.N] = s.xs[(size_t)s.N - 1];
      std::rotate(s.us.begin(), s.us.begin() + 1, s.us.end());
      s.us[(size_t)s.N - 1] = s.us[(size_t)s.N - 2];
      s.run(x0.data(), max_iters);
    }
    catch (const std::exception & e)
    {
#pragma omp critical
      err = e.what();
    }
  }
  if (!err.empty())
    throw std::runtime_error(err);
  ORC_CATCH
}
} // extern "C" (driver port)

extern "C" {

// ------------------------------- unit-level entry points -------------------------------------

// LQ knot t of instance `inst` as assembled by the last solve: row-major dense blocks; any pointer may be NULL.
// sizes: Q ndx*ndx, S ndx*nu, R nu*nu, q ndx, r nu, A ndx*ndx, B ndx*nu, f ndx, C ncmax*ndx, D ncmax*nu, d ncmax
int orc_get_knot(
  orc_ctx * c, int inst, int t, int * nc, double * Q, double * S, double * R, double * q, double * r, double * A, double * B,
  double * f, double * C, double * D, double * d)
{
  if (check_range(c, inst, 1))
    return SMPC_ERR_INVALID;
  Solver & s = *c->inst[inst];
  if (t < 0 || t > s.N || (int)s.knots.size() != s.N + 1)
  {
    g_err = "no LQ problem assembled / bad index";
    return SMPC_ERR_INVALID;
  }
  const Knot & k = s.knots[t];
  auto cp = [](double * dst, const std::vector<double> & v) {
    if (dst && !v.empty())
      memcpy(dst, v.data(), v.size() * sizeof(double));
  };
  if (nc)
    *nc = k.nc;
  cp(Q, k.Q.d);
  cp(q, k.q);
  cp(C, k.C.d);
  cp(d, k.d);
  if (t < s.N)
  {
    cp(S, k.S.d);
    cp(R, k.R.d);
    cp(r, k.r);
    cp(A, k.A.d);
    cp(B, k.B.d);
    cp(f, k.f);
    cp(D, k.D.d);
  }
  return SMPC_OK;
}
// value function / gains of node t from the last backward sweep
int orc_get_stage_sol(orc_ctx * c, int inst, int t, double * P, double * p, double * K, double * k, double * Z, double * z)
{
  if (check_range(c, inst, 1))
    return SMPC_ERR_INVALID;
  Solver & s = *c->inst[inst];
  if (t < 0 || t > s.N || (int)s.sol.size() != s.N + 1)
  {
    g_err = "no Riccati solution / bad index";
    return SMPC_ERR_INVALID;
  }
  const StageSol & so = s.sol[t];
  auto cp = [](double * dst, const std::vector<double> & v) {
    if (dst && !v.empty())
      memcpy(dst, v.data(), v.size() * sizeof(double));
  };
  cp(P, so.P.d);
  cp(p, so.p);
  cp(K, so.K.d);
  cp(k, so.k);
  cp(Z, so.Z.d);
  cp(z, so.z);
  return SMPC_OK;
}
// Newton step of the last iteration: dxs [N+1][ndx], dus [N][nu]
int orc_get_step(orc_ctx * c, int inst, double * dxs, double * dus)
{
  if (check_range(c, inst, 1))
    return SMPC_ERR_INVALID;
  Solver & s = *c->inst[inst];
  if ((int)s.dxs.size() != s.N + 1)
  {
    g_err = "no step computed";
    return SMPC_ERR_INVALID;
  }
  for (int t = 0; t <= s.N; t++)
    if (dxs)
      memcpy(dxs + (size_t)t * s.ndx, s.dxs[t].data(), s.ndx * sizeof(double));
  for (int t = 0; t < s.N; t++)
    if (dus)
      memcpy(dus + (size_t)t * s.nu, s.dus[t].data(), s.nu * sizeof(double));
  return SMPC_OK;
}

// One stage evaluation (values, and derivatives if derivs != 0). Outputs (any may be NULL):
// xnext nx, xdot ndx, cost 1, c ncmax, ctype ncmax, A ndx*ndx, B ndx*nu, lx ndx, lu nu, lxx, lxu, luu, Cx nc*ndx, Cu nc*nu
int orc_stage_eval(
  const smpc_model_desc * model, const smpc_ocp_desc * ocp, const smpc_slot * slot, const double * x, const double * u,
  int derivs, int * nc, double * xnext, double * xdot, double * cost, double * cval, int * ctype, double * A, double * B,
  double * lx, double * lu, double * lxx, double * lxu, double * luu, double * Cx, double * Cu)
{
  ORC_TRY
  Problem P(*model, *ocp);
  StageEval e;
  stage_eval(P, *slot, x, u, derivs != 0, e);
  auto cp = [](double * dst, const std::vector<double> & v) {
    if (dst && !v.empty())
      memcpy(dst, v.data(), v.size() * sizeof(double));
  };
  if (nc)
    *nc = e.nc;
  cp(xnext, e.xnext);
  cp(xdot, e.xdot);
  if (cost)
    *cost = e.cost;
  cp(cval, e.c);
  if (ctype)
    for (int i = 0; i < e.nc; i++)
      ctype[i] = e.ctype[i];
  if (derivs)
  {
    cp(A, e.A.d);
    cp(B, e.B.d);
    cp(lx, e.lx);
    cp(lu, e.lu);
    cp(lxx, e.lxx.d);
    cp(lxu, e.lxu.d);
    cp(luu, e.luu.d);
    cp(Cx, e.Cx.d);
    cp(Cu, e.Cu.d);
  }
  ORC_CATCH
}
int orc_terminal_eval(
  const smpc_model_desc * model, const smpc_ocp_desc * ocp, const smpc_terminal * term, const double * x, int derivs, int * nc,
  double * cost, double * cval, double * lx, double * lxx, double * Cx)
{
  ORC_TRY
  Problem P(*model, *ocp);
  StageEval e;
  terminal_eval(P, *term, x, derivs != 0, e);
  if (nc)
    *nc = e.nc;
  if (cost)
    *cost = e.cost;
  if (cval && e.nc)
    memcpy(cval, e.c.data(), e.nc * sizeof(double));
  if (derivs)
  {
    if (lx)
      memcpy(lx, e.lx.data(), e.lx.size() * sizeof(double));
    if (lxx)
      memcpy(lxx, e.lxx.d.data(), e.lxx.d.size() * sizeof(double));
    if (Cx && e.nc)
      memcpy(Cx, e.Cx.d.data(), e.Cx.d.size() * sizeof(double));
  }
  ORC_CATCH
}

// Rigid-contact forward dynamics (orc_full.hpp) exposed for the finite-difference tests. Outputs may be NULL.
int orc_full_dynamics(
  const smpc_model_desc * model, const double * q, const double * v, const double * tau, unsigned mask, const double * gravity,
  const double * Kp, const double * Kd, int derivs, double * a, double * lam, double * da_dq, double * da_dv, double * da_dtau,
  double * dl_dq, double * dl_dv, double * dl_dtau, int contact_dim)
{
  ORC_TRY
  Model m(*model);
  FullDyn fd;
  full_dynamics(m, q, v, tau, mask, gravity, Kp, Kd, derivs != 0, fd, contact_dim == 6 ? 6 : 3);
  auto cp = [](double * dst, const std::vector<double> & vv) {
    if (dst && !vv.empty())
      memcpy(dst, vv.data(), vv.size() * sizeof(double));
  };
  cp(a, fd.a);
  cp(lam, fd.lam);
  if (derivs)
  {
    cp(da_dq, fd.da_dq.d);
    cp(da_dv, fd.da_dv.d);
    cp(da_dtau, fd.da_dtau.d);
    cp(dl_dq, fd.dl_dq.d);
    cp(dl_dv, fd.dl_dv.d);
    cp(dl_dtau, fd.dl_dtau.d);
  }
  ORC_CATCH
}

// x (+) dx and x1 (-) x0 of the OCP's state space
int orc_state_integrate(const smpc_model_desc * model, const smpc_ocp_desc * ocp, const double * x, const double * dx, double * xo)
{
  ORC_TRY
  Problem P(*model, *ocp);
  state_integrate(P, x, dx, xo);
  ORC_CATCH
}
int orc_state_difference(const smpc_model_desc * model, const smpc_ocp_desc * ocp, const double * x0, const double * x1, double * d)
{
  ORC_TRY
  Problem P(*model, *ocp);
  state_difference(P, x0, x1, d);
  ORC_CATCH
}

// Kinematics + centroidal quantities at (q, v, a): what RobotDataHandler::updateInternalData exposes
// (reference src/robot-handler.cpp:114-127,142-149) plus the centroidal derivatives.
// outputs (any may be NULL): foot_pose nfeet*12, footref_pose nfeet*12, base_pose 12, com 3, hg 6, hdg 6,
// Ag 6*nv, dh_dq 6*nv, dhd_dq 6*nv, dhd_dv 6*nv, Jcom 3*nv
int orc_kinematics(
  const smpc_model_desc * model, const double * q, const double * v, const double * a, double * foot_pose,
  double * footref_pose, double * base_pose, double * com, double * hg, double * hdg, double * Ag, double * dh_dq,
  double * dhd_dq, double * dhd_dv, double * Jcom)
{
  ORC_TRY
  Model m(*model);
  Kin k;
  centroidal_pass(m, q, v, a, true, k);
  auto pose = [](double * dst, const SE3 & M) {
    memcpy(dst, M.R, 9 * sizeof(double));
    memcpy(dst + 9, M.p, 3 * sizeof(double));
  };
  for (int f = 0; f < m.nfeet; f++)
  {
    if (foot_pose)
      pose(foot_pose + 12 * f, k.oMf[f]);
    if (footref_pose)
      pose(footref_pose + 12 * f, k.oMfref[f]);
  }
  if (base_pose)
    pose(base_pose, k.oMbase);
  if (com)
    memcpy(com, k.com, sizeof(k.com));
  if (hg)
    memcpy(hg, k.hg, sizeof(k.hg));
  if (hdg)
    memcpy(hdg, k.hdg, sizeof(k.hdg));
  auto cp = [](double * dst, const Mat & M) {
    if (dst)
      memcpy(dst, M.d.data(), M.d.size() * sizeof(double));
  };
  cp(Ag, k.Ag);
  cp(dh_dq, k.dh_dq);
  cp(dhd_dq, k.dhd_dq);
  cp(dhd_dv, k.dhd_dv);
  cp(Jcom, k.Jcom);
  ORC_CATCH
}

// SE(3) helpers for known-answer tests
void orc_exp6(const double * nu, double * pose12)
{
  SE3 M = exp6(nu);
  memcpy(pose12, M.R, 9 * sizeof(double));
  memcpy(pose12 + 9, M.p, 3 * sizeof(double));
}
void orc_log6(const double * pose12, double * nu)
{
  SE3 M;
  memcpy(M.R, pose12, 9 * sizeof(double));
  memcpy(M.p, pose12 + 9, 3 * sizeof(double));
  log6(M, nu);
}
void orc_Jexp6(const double * nu, double * J) { Jexp6(nu, J); }
void orc_Jlog6(const double * nu, double * J) { Jlog6(nu, J); }

// Standalone proximal Riccati solve on caller-supplied dense knots (all stages share nc = ncs, terminal nc = nct).
// Layout per stage: Q S R q r A B f C D d (row-major, sizes as in orc_get_knot with ncmax = ncs); terminal: Q q C d.
// Outputs: K [N][nu][ndx], k [N][nu], P0 ndx*ndx, p0 ndx, dxs [N+1][ndx], dus [N][nu]
int orc_riccati(
  int N, int ndx, int nu, int ncs, int nct, double mudyn, double mueq, const double * stages, const double * terminal,
  double * K, double * k, double * P0, double * p0, double * dxs_o, double * dus_o)
{
  ORC_TRY
  std::vector<Knot> knots(N + 1);
  const double * p = stages;
  auto take = [&](Mat & M, int r, int c) {
    M.resize(r, c);
    memcpy(M.d.data(), p, sizeof(double) * r * c);
    p += (size_t)r * c;
  };
  auto takev = [&](Vec & v, int n) {
    v.assign(p, p + n);
    p += n;
  };
  for (int t = 0; t < N; t++)
  {
    Knot & kn = knots[t];
    kn.nc = ncs;
    take(kn.Q, ndx, ndx);
    take(kn.S, ndx, nu);
    take(kn.R, nu, nu);
    takev(kn.q, ndx);
    takev(kn.r, nu);
    take(kn.A, ndx, ndx);
    take(kn.B, ndx, nu);
    takev(kn.f, ndx);
    take(kn.C, ncs, ndx);
    take(kn.D, ncs, nu);
    takev(kn.d, ncs);
  }
  p = terminal;
  {
    Knot & kn = knots[N];
    kn.nc = nct;
    take(kn.Q, ndx, ndx);
    takev(kn.q, ndx);
    take(kn.C, nct, ndx);
    takev(kn.d, nct);
  }
  std::vector<StageSol> sol;
  RiccatiStats st = riccati_backward(knots, ndx, nu, mudyn, mueq, sol);
  std::vector<Vec> dxs, dus, vnew, lnew;
  riccati_forward(knots, sol, ndx, nu, mudyn, dxs, dus, vnew, lnew);
  for (int t = 0; t < N; t++)
  {
    if (K)
      memcpy(K + (size_t)t * nu * ndx, sol[t].K.d.data(), sizeof(double) * nu * ndx);
    if (k)
      memcpy(k + (size_t)t * nu, sol[t].k.data(), sizeof(double) * nu);
    if (dus_o)
      memcpy(dus_o + (size_t)t * nu, dus[t].data(), sizeof(double) * nu);
  }
  for (int t = 0; t <= N; t++)
    if (dxs_o)
      memcpy(dxs_o + (size_t)t * ndx, dxs[t].data(), sizeof(double) * ndx);
  if (P0)
    memcpy(P0, sol[0].P.d.data(), sizeof(double) * ndx * ndx);
  if (p0)
    memcpy(p0, sol[0].p.data(), sizeof(double) * ndx);
  if (!st.ok)
    throw std::runtime_error("riccati: non-positive pivot");
  ORC_CATCH
}

} // extern "C"
