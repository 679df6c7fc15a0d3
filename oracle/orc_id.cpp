// ORACLE — test infrastructure only (see orc_math.hpp header). PARITY UNPINNED.
//
// orc_id.cpp — C entry points of the CPU oracle for the inverse-dynamics QP; signatures mirror include/smpc_id.h
// (prefix orc_id_ instead of smpc_id_). Loaded by tests/, __graft_entry__.smoke() and bench.py's CPU legs only.
#include "orc_full.hpp"
#include "orc_id.hpp"
#include <string>
#include <vector>
#include <omp.h>

namespace
{
  thread_local std::string g_id_err;
  struct IdCtx
  {
    orc::Model model;
    smpc_id_desc desc;
    int batch;
    bool warm = true;
    std::vector<orc::IdTarget> tgt;
    std::vector<orc::IdSolution> sol;
    IdCtx(const smpc_model_desc & m, const smpc_id_desc & d, int B) : model(m), desc(d), batch(B), tgt(B), sol(B) {}
  };
  int fail(const char * what)
  {
    g_id_err = what;
    return SMPC_ERR_INVALID;
  }
  // CentroidalID::setTarget for one instance (centroidal-id.cpp:88-145)
  void set_target_centroidal(
    IdCtx * c, int b, const double * com, const double * vcom, const double * pose, const double * vel, const int32_t * contact,
    const double * f_target)
  {
    orc::IdTarget & T = c->tgt[b];
    const orc::Model & m = c->model;
    for (int r = 0; r < 3; r++)
    {
      T.com_ref[r] = com[r];
      T.com_vel[r] = vcom[r];
    }
    if (c->desc.settings.w_feet_tracking > 0.)
      for (int f = 0; f < m.nfeet; f++)
      {
        T.tracked[f] = contact[f] ? 0 : 1;
        if (!contact[f])
        {
          memcpy(T.foot_ref[f].R, pose + 12 * f, 9 * sizeof(double));
          memcpy(T.foot_ref[f].p, pose + 12 * f + 9, 3 * sizeof(double));
          memcpy(T.foot_vel[f], vel + 6 * f, 6 * sizeof(double));
        }
      }
    std::vector<double> q(m.d.q_ref, m.d.q_ref + m.nq), z(m.nv, 0.0);
    orc::id_set_target(m, T, q.data(), z.data(), z.data(), contact, f_target);
  }
} // namespace

extern "C" {

const char * orc_id_last_error(void) { return g_id_err.c_str(); }

void orc_id_default_settings(smpc_id_settings * s)
{
  memset(s, 0, sizeof(*s));
  s->friction_coefficient = 0.6;
  s->contact_weight_ratio_max = 10.0;
  s->contact_weight_ratio_min = 0.01;
  s->w_base = s->w_posture = s->w_contact_motion = s->w_contact_force = -1.;
  s->w_com = s->w_feet_tracking = -1.;
}

int orc_id_solve(IdCtx * c, double t, const double * q_meas, const double * v_meas, double * tau_res);

int orc_id_create(IdCtx ** out, const smpc_model_desc * model, const smpc_id_desc * desc, int batch)
{
  if (!out || !model || !desc || batch < 1)
    return fail("orc_id_create: bad argument");
  try
  {
    IdCtx * c = new IdCtx(*model, *desc, batch);
    const orc::Model & m = c->model;
    const int nfeet = m.nfeet;
    std::vector<double> q(m.d.q_ref, m.d.q_ref + m.nq), z(m.nv, 0.0), f(6 * nfeet, 0.0);
    std::vector<int32_t> cs(nfeet, 1);
    for (int i = 0; i < nfeet; i++)
      f[6 * i + 2] = m.total_mass * 9.81 / nfeet;
    orc::Kin k;
    orc::centroidal_pass(m, q.data(), z.data(), nullptr, false, k);
    std::vector<double> pose(12 * nfeet), vel(6 * nfeet, 0.0);
    for (int i = 0; i < nfeet; i++)
    { // data_handler_.getFootRefPose(i) (centroidal-id.cpp:81)
      memcpy(&pose[12 * i], k.oMfref[i].R, 9 * sizeof(double));
      memcpy(&pose[12 * i + 9], k.oMfref[i].p, 3 * sizeof(double));
    }
    const double vz[3] = {0, 0, 0};
    for (int b = 0; b < batch; b++)
    {
      orc::IdTarget & T = c->tgt[b];
      memset(T.tracked, 0, sizeof(T.tracked));
      memset(T.com_ref, 0, sizeof(T.com_ref));
      memset(T.com_vel, 0, sizeof(T.com_vel));
      memset(T.foot_vel, 0, sizeof(T.foot_vel));
      if (desc->settings.centroidal)
        set_target_centroidal(c, b, k.com, vz, pose.data(), vel.data(), cs.data(), f.data());
      else
        orc::id_set_target(m, T, q.data(), z.data(), z.data(), cs.data(), f.data());
    }
    // dry run at the reference state (kinodynamics-id.cpp:116-118)
    std::vector<double> qb((size_t)batch * m.nq), vb((size_t)batch * m.nv, 0.0), tau((size_t)batch * (m.nv - 6));
    for (int b = 0; b < batch; b++)
      memcpy(&qb[(size_t)b * m.nq], q.data(), m.nq * sizeof(double));
    *out = c;
    int rc = orc_id_solve(c, 0.0, qb.data(), vb.data(), tau.data());
    if (rc)
    {
      delete c;
      *out = nullptr;
    }
    return rc;
  }
  catch (const std::exception & e)
  {
    return fail(e.what());
  }
}

void orc_id_destroy(IdCtx * c) { delete c; }

int orc_id_set_target(
  IdCtx * c, int first, int n_inst, const double * q, const double * v, const double * a, const int32_t * contact, const double * f_target)
{
  if (!c || first < 0 || n_inst < 0 || first + n_inst > c->batch)
    return fail("orc_id_set_target: bad range");
  if (c->desc.settings.centroidal)
    return fail("orc_id_set_target: CentroidalID takes set_target_centroidal (KinodynamicsID::setTarget is private there)");
  const orc::Model & m = c->model;
  for (int i = 0; i < n_inst; i++)
    orc::id_set_target(
      m, c->tgt[first + i], q + (size_t)i * m.nq, v + (size_t)i * m.nv, a + (size_t)i * m.nv, contact + (size_t)i * m.nfeet,
      f_target ? f_target + (size_t)i * m.nfeet * 6 : nullptr);
  return 0;
}

int orc_id_set_target_centroidal(
  IdCtx * c, int first, int n_inst, const double * com, const double * vcom, const double * pose, const double * vel,
  const int32_t * contact, const double * f_target)
{
  if (!c || first < 0 || n_inst < 0 || first + n_inst > c->batch)
    return fail("orc_id_set_target_centroidal: bad range");
  if (!c->desc.settings.centroidal)
    return fail("orc_id_set_target_centroidal: not a CentroidalID context");
  const int nf = c->model.nfeet;
  for (int i = 0; i < n_inst; i++)
    set_target_centroidal(
      c, first + i, com + 3 * i, vcom + 3 * i, pose + (size_t)i * nf * 12, vel + (size_t)i * nf * 6, contact + (size_t)i * nf,
      f_target + (size_t)i * nf * 6);
  return 0;
}

int orc_id_solve(IdCtx * c, double, const double * q_meas, const double * v_meas, double * tau_res)
{
  if (!c || !q_meas || !v_meas || !tau_res)
    return fail("orc_id_solve: bad argument");
  const orc::Model & m = c->model;
  const int na = m.nv - 6;
  int bad = 0;
#pragma omp parallel for schedule(dynamic)
  for (int b = 0; b < c->batch; b++)
  {
    try
    {
      orc::id_solve(m, c->desc, c->tgt[b], q_meas + (size_t)b * m.nq, v_meas + (size_t)b * m.nv, c->sol[b], c->warm);
      memcpy(tau_res + (size_t)b * na, c->sol[b].tau.data(), na * sizeof(double));
    }
    catch (const std::exception & e)
    {
#pragma omp critical
      {
        bad = 1;
        g_id_err = e.what();
      }
    }
  }
  return bad ? SMPC_ERR_INVALID : 0;
}

int orc_id_set_warm_start(IdCtx * c, int on)
{
  if (!c)
    return fail("orc_id_set_warm_start: null context");
  c->warm = on != 0;
  if (!on)
    for (auto & s : c->sol)
      s.sig = -1.0;
  return 0;
}

int orc_id_get_accelerations(IdCtx * c, int first, int n_inst, double * a)
{
  if (!c || first < 0 || first + n_inst > c->batch)
    return fail("orc_id_get_accelerations: bad range");
  for (int i = 0; i < n_inst; i++)
    memcpy(a + (size_t)i * c->model.nv, c->sol[first + i].a.data(), c->model.nv * sizeof(double));
  return 0;
}

int orc_id_get_forces(IdCtx * c, int first, int n_inst, double * f)
{
  if (!c || first < 0 || first + n_inst > c->batch)
    return fail("orc_id_get_forces: bad range");
  const int nf = c->model.nfeet;
  for (int i = 0; i < n_inst; i++)
    for (int k = 0; k < nf; k++)
      memcpy(f + ((size_t)i * nf + k) * SMPC_ID_MAX_FVAR, c->sol[first + i].f[k], SMPC_ID_MAX_FVAR * sizeof(double));
  return 0;
}

int orc_id_get_info(IdCtx * c, int first, int n_inst, smpc_id_info * info)
{
  if (!c || first < 0 || first + n_inst > c->batch)
    return fail("orc_id_get_info: bad range");
  for (int i = 0; i < n_inst; i++)
    info[i] = c->sol[first + i].info;
  return 0;
}

} // extern "C"
