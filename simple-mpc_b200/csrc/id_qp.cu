// id_qp.cu — the C ABI of include/smpc_id.h: batched KinodynamicsID / CentroidalID (reference
// src/inverse-dynamics/*.cpp) on sm_100a. One CTA of 128 threads per controller instance, everything in shared memory
// (id_qp.cuh); a context owns the device-resident model / settings / targets and pinned staging buffers, one stream.
// There is no CPU path: every entry point that computes launches k_id_solve or fails with SMPC_ERR_CUDA.
#include "ctx.h"
#include "host_build.h"
#include "id_host.h"
#include <cuda_runtime.h>
#include <string>
#include <vector>

using namespace smpc;

namespace
{
  __global__ void __launch_bounds__(ID_THREADS, 5) k_id_solve(
    const DevModel * __restrict__ M, const smpc_id_desc * __restrict__ D, const IdTargetDev * __restrict__ T, const IdDims L,
    const double * __restrict__ q, const double * __restrict__ v, double * __restrict__ tau, double * __restrict__ a,
    double * __restrict__ f, smpc_id_info * __restrict__ info, double * __restrict__ ws, int batch)
  {
    extern __shared__ __align__(16) double sm[];
    const int b = blockIdx.x;
    if (b >= batch)
      return;
    id_solve_cta(
      *M, *D, T[b], L, q + (size_t)b * L.nq, v + (size_t)b * L.nv, sm, tau + (size_t)b * L.na, a + (size_t)b * L.nv,
      f + (size_t)b * L.nfeet * SMPC_ID_MAX_FVAR, info + b, ws ? ws + (size_t)b * L.ws_stride : nullptr);
  }
} // namespace

struct smpc_id_ctx
{
  smpc_model_desc model;
  smpc_id_desc desc;
  IdDims L;
  int batch = 0;
  size_t smem = 0;
  DevModel * d_model = nullptr;
  smpc_id_desc * d_desc = nullptr;
  IdTargetDev * d_tgt = nullptr;
  std::vector<IdTargetDev> h_tgt;
  bool tgt_dirty = true;
  double *d_q = nullptr, *d_v = nullptr, *d_tau = nullptr, *d_a = nullptr, *d_f = nullptr;
  smpc_id_info * d_info = nullptr;
  double * d_ws = nullptr; // previous solution of every instance (warm start)
  bool warm_start = true;
  double *h_q = nullptr, *h_v = nullptr, *h_tau = nullptr; // pinned staging
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int64_t launches = 0;
  bool timed = false;
};

namespace
{
  int bad(const char * what)
  {
    smpc::err() = what;
    return SMPC_ERR_INVALID;
  }
  int range_ok(smpc_id_ctx * c, int first, int n)
  {
    if (!c)
      return bad("null inverse-dynamics context");
    if (first < 0 || n < 0 || first + n > c->batch)
      return bad("instance range out of bounds");
    return 0;
  }
  int upload_targets(smpc_id_ctx * c)
  {
    if (c->tgt_dirty)
    {
      CU(cudaMemcpyAsync(c->d_tgt, c->h_tgt.data(), sizeof(IdTargetDev) * c->batch, cudaMemcpyHostToDevice, c->stream));
      c->tgt_dirty = false;
    }
    return 0;
  }
  int launch(smpc_id_ctx * c, const double * q_dev, const double * v_dev, double * tau_dev)
  {
    int rc = upload_targets(c);
    if (rc)
      return rc;
    CU(cudaEventRecord(c->e0, c->stream));
    k_id_solve<<<c->batch, ID_THREADS, c->smem, c->stream>>>(
      c->d_model, c->d_desc, c->d_tgt, c->L, q_dev, v_dev, tau_dev, c->d_a, c->d_f, c->d_info, c->warm_start ? c->d_ws : nullptr, c->batch);
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->e1, c->stream));
    c->launches++;
    c->timed = true;
    return 0;
  }
} // namespace

extern "C" {

void smpc_id_default_settings(smpc_id_settings * s)
{
  memset(s, 0, sizeof(*s));
  s->friction_coefficient = 0.6;
  s->contact_weight_ratio_max = 10.0;
  s->contact_weight_ratio_min = 0.01;
  s->w_base = s->w_posture = s->w_contact_motion = s->w_contact_force = -1.;
  s->w_com = s->w_feet_tracking = -1.;
}

void smpc_id_destroy(smpc_id_ctx * c)
{
  if (!c)
    return;
  cudaFree(c->d_model), cudaFree(c->d_desc), cudaFree(c->d_tgt);
  cudaFree(c->d_ws);
  cudaFree(c->d_q), cudaFree(c->d_v), cudaFree(c->d_tau), cudaFree(c->d_a), cudaFree(c->d_f), cudaFree(c->d_info);
  cudaFreeHost(c->h_q), cudaFreeHost(c->h_v), cudaFreeHost(c->h_tau);
  if (c->e0)
    cudaEventDestroy(c->e0);
  if (c->e1)
    cudaEventDestroy(c->e1);
  if (c->stream)
    cudaStreamDestroy(c->stream);
  delete c;
}

static int id_create_impl(smpc_id_ctx * c, const smpc_model_desc * model, const smpc_id_desc * desc, int batch)
{
  DevModel hM;
  try
  {
    build_dev_model(*model, hM);
  }
  catch (const std::exception & e)
  {
    smpc::err() = e.what();
    return SMPC_ERR_INVALID;
  }
  for (int f = 0; f < model->nfeet; f++)
    if (desc->foot_type[f] != SMPC_ID_POINT && desc->foot_type[f] != SMPC_ID_QUAD)
      return bad("smpc_id_create: foot_type must be SMPC_ID_POINT or SMPC_ID_QUAD");
  if (!(desc->control_dt > 0))
    return bad("smpc_id_create: control_dt must be positive");
  c->model = *model;
  c->desc = *desc;
  c->batch = batch;
  c->L = id_make_dims(*model, *desc);
  c->smem = sizeof(double) * (size_t)c->L.total;
  if (c->L.nx > 64) // the back substitution keeps two rows per lane of one warp
  {
    smpc::err() = "smpc_id_create: " + std::to_string(c->L.nx) + " QP variables exceed what the one-warp back substitution is laid out for (64)";
    return SMPC_ERR_UNSUPPORTED;
  }
  int dev = 0, max_smem = 0;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (c->smem > (size_t)max_smem)
  {
    smpc::err() = "smpc_id_create: the QP of this robot needs " + std::to_string(c->smem) + " bytes of shared memory per CTA";
    return SMPC_ERR_UNSUPPORTED;
  }
  CU(cudaFuncSetAttribute(k_id_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem));
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&c->e0));
  CU(cudaEventCreate(&c->e1));
  const size_t B = (size_t)batch;
  CU(cudaMalloc(&c->d_model, sizeof(DevModel)));
  CU(cudaMalloc(&c->d_desc, sizeof(smpc_id_desc)));
  CU(cudaMalloc(&c->d_tgt, sizeof(IdTargetDev) * B));
  CU(cudaMalloc(&c->d_q, sizeof(double) * B * model->nq));
  CU(cudaMalloc(&c->d_v, sizeof(double) * B * model->nv));
  CU(cudaMalloc(&c->d_tau, sizeof(double) * B * (model->nv - 6)));
  CU(cudaMalloc(&c->d_a, sizeof(double) * B * model->nv));
  CU(cudaMalloc(&c->d_f, sizeof(double) * B * model->nfeet * SMPC_ID_MAX_FVAR));
  CU(cudaMalloc(&c->d_info, sizeof(smpc_id_info) * B));
  CU(cudaMalloc(&c->d_ws, sizeof(double) * B * c->L.ws_stride));
  CU(cudaMemset(c->d_ws, 0xff, sizeof(double) * B * c->L.ws_stride)); // NaN signatures: no previous solution
  CU(cudaMallocHost(&c->h_q, sizeof(double) * B * model->nq));
  CU(cudaMallocHost(&c->h_v, sizeof(double) * B * model->nv));
  CU(cudaMallocHost(&c->h_tau, sizeof(double) * B * (model->nv - 6)));
  CU(cudaMemcpy(c->d_model, &hM, sizeof(DevModel), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(c->d_desc, desc, sizeof(smpc_id_desc), cudaMemcpyHostToDevice));
  // default target + dry run at the reference state (kinodynamics-id.cpp:100-118)
  IdTargetDev T0;
  id_host_default_target(*model, *desc, T0);
  c->h_tgt.assign(B, T0);
  c->tgt_dirty = true;
  for (size_t b = 0; b < B; b++)
  {
    memcpy(c->h_q + b * model->nq, model->q_ref, sizeof(double) * model->nq);
    memset(c->h_v + b * model->nv, 0, sizeof(double) * model->nv);
  }
  CU(cudaMemcpyAsync(c->d_q, c->h_q, sizeof(double) * B * model->nq, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_v, c->h_v, sizeof(double) * B * model->nv, cudaMemcpyHostToDevice, c->stream));
  int rc = launch(c, c->d_q, c->d_v, c->d_tau);
  if (rc)
    return rc;
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int smpc_id_create(smpc_id_ctx ** out, const smpc_model_desc * model, const smpc_id_desc * desc, int batch)
{
  if (!out || !model || !desc || batch < 1)
    return bad("smpc_id_create: bad arguments");
  *out = nullptr;
  smpc_id_ctx * c = new smpc_id_ctx;
  const int rc = id_create_impl(c, model, desc, batch);
  if (rc)
  {
    smpc_id_destroy(c); // every failure exit releases what was allocated so far
    return rc;
  }
  *out = c;
  return 0;
}

int smpc_id_set_target(
  smpc_id_ctx * c, int first, int n_inst, const double * q, const double * v, const double * a, const int32_t * contact,
  const double * f_target)
{
  if (int rc = range_ok(c, first, n_inst))
    return rc;
  if (c->desc.settings.centroidal)
    return bad("smpc_id_set_target: a CentroidalID context takes smpc_id_set_target_centroidal (KinodynamicsID::setTarget is private there)");
  if (!q || !v || !a || !contact)
    return bad("smpc_id_set_target: null argument");
  const smpc_model_desc & m = c->model;
  for (int i = 0; i < n_inst; i++)
    id_host_set_target(
      m, c->h_tgt[first + i], q + (size_t)i * m.nq, v + (size_t)i * m.nv, a + (size_t)i * m.nv, contact + (size_t)i * m.nfeet,
      f_target ? f_target + (size_t)i * m.nfeet * 6 : nullptr);
  c->tgt_dirty = true;
  return 0;
}

int smpc_id_set_target_centroidal(
  smpc_id_ctx * c, int first, int n_inst, const double * com, const double * vcom, const double * pose, const double * vel,
  const int32_t * contact, const double * f_target)
{
  if (int rc = range_ok(c, first, n_inst))
    return rc;
  if (!c->desc.settings.centroidal)
    return bad("smpc_id_set_target_centroidal: not a CentroidalID context");
  if (!com || !vcom || !pose || !vel || !contact || !f_target)
    return bad("smpc_id_set_target_centroidal: null argument");
  const int nf = c->model.nfeet;
  for (int i = 0; i < n_inst; i++)
    id_host_set_target_centroidal(
      c->model, c->desc, c->h_tgt[first + i], com + 3 * (size_t)i, vcom + 3 * (size_t)i, pose + (size_t)i * nf * 12,
      vel + (size_t)i * nf * 6, contact + (size_t)i * nf, f_target + (size_t)i * nf * 6);
  c->tgt_dirty = true;
  return 0;
}

int smpc_id_solve(smpc_id_ctx * c, double, const double * q_meas, const double * v_meas, double * tau_res)
{
  if (!c || !q_meas || !v_meas || !tau_res)
    return bad("smpc_id_solve: null argument");
  const size_t B = (size_t)c->batch, nq = c->model.nq, nv = c->model.nv, na = nv - 6;
  memcpy(c->h_q, q_meas, sizeof(double) * B * nq);
  memcpy(c->h_v, v_meas, sizeof(double) * B * nv);
  CU(cudaMemcpyAsync(c->d_q, c->h_q, sizeof(double) * B * nq, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_v, c->h_v, sizeof(double) * B * nv, cudaMemcpyHostToDevice, c->stream));
  if (int rc = launch(c, c->d_q, c->d_v, c->d_tau))
    return rc;
  CU(cudaMemcpyAsync(c->h_tau, c->d_tau, sizeof(double) * B * na, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  memcpy(tau_res, c->h_tau, sizeof(double) * B * na);
  return 0;
}

int smpc_id_solve_dev(smpc_id_ctx * c, double, const double * q_dev, const double * v_dev, double * tau_dev)
{
  if (!c || !q_dev || !v_dev || !tau_dev)
    return bad("smpc_id_solve_dev: null argument");
  return launch(c, q_dev, v_dev, tau_dev);
}

int smpc_id_get_accelerations(smpc_id_ctx * c, int first, int n_inst, double * a)
{
  if (int rc = range_ok(c, first, n_inst))
    return rc;
  const size_t nv = c->model.nv;
  CU(cudaMemcpyAsync(a, c->d_a + first * nv, sizeof(double) * n_inst * nv, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int smpc_id_get_forces(smpc_id_ctx * c, int first, int n_inst, double * f)
{
  if (int rc = range_ok(c, first, n_inst))
    return rc;
  const size_t w = (size_t)c->model.nfeet * SMPC_ID_MAX_FVAR;
  CU(cudaMemcpyAsync(f, c->d_f + first * w, sizeof(double) * n_inst * w, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int smpc_id_get_info(smpc_id_ctx * c, int first, int n_inst, smpc_id_info * info)
{
  if (int rc = range_ok(c, first, n_inst))
    return rc;
  CU(cudaMemcpyAsync(info, c->d_info + first, sizeof(smpc_id_info) * n_inst, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int smpc_id_last_solve_ms(smpc_id_ctx * c, float * ms)
{
  if (!c || !ms)
    return bad("smpc_id_last_solve_ms: null argument");
  if (!c->timed)
  {
    smpc::err() = "smpc_id_last_solve_ms: no solve yet";
    return SMPC_ERR_STATE;
  }
  CU(cudaEventSynchronize(c->e1));
  CU(cudaEventElapsedTime(ms, c->e0, c->e1));
  return 0;
}

int64_t smpc_id_launch_count(smpc_id_ctx * c) { return c ? c->launches : 0; }

int smpc_id_set_warm_start(smpc_id_ctx * c, int on)
{
  if (!c)
    return bad("smpc_id_set_warm_start: null context");
  c->warm_start = on != 0;
  if (!on) // forget the stored solutions, so that switching it back on starts cold once
    CU(cudaMemsetAsync(c->d_ws, 0xff, sizeof(double) * (size_t)c->batch * c->L.ws_stride, c->stream));
  return 0;
}

} // extern "C"
