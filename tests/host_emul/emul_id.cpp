// TEST HARNESS — not part of the product (see emul.cpp). CPU emulation of the inverse-dynamics QP CTA routine
// (simple-mpc_b200/csrc/id_qp.cuh) so that its logic can be compared with the oracle on the GPU-less development box.
#include "../../simple-mpc_b200/csrc/id_host.h"
#include "../../simple-mpc_b200/csrc/host_build.h"
#include <vector>

using namespace smpc;

extern "C" {

// One KinodynamicsID / CentroidalID solve for one instance whose target was set by the kinodynamics-style setTarget
// (centroidal != 0 in the settings: the CentroidalID arguments). Outputs: tau na, a nv, f nfeet x 12.
int emul_id_solve(
  const smpc_model_desc * model, const smpc_id_desc * desc, const double * q_target, const double * v_target, const double * a_target,
  const double * com, const double * vcom, const double * feet_pose, const double * feet_vel, const int32_t * contact,
  const double * f_target, int use_default_target, const double * q, const double * v, double * tau, double * a, double * f,
  smpc_id_info * info, double * ws /* warm-start record of emul_id_ws_doubles() entries, or NULL */)
{
  try
  {
    DevModel Mo;
    build_dev_model(*model, Mo);
    IdDims L = id_make_dims(*model, *desc);
    IdTargetDev T;
    id_host_default_target(*model, *desc, T);
    if (!use_default_target)
    {
      if (desc->settings.centroidal)
        id_host_set_target_centroidal(*model, *desc, T, com, vcom, feet_pose, feet_vel, contact, f_target);
      else
        id_host_set_target(*model, T, q_target, v_target, a_target, contact, f_target);
    }
    std::vector<double> sm(L.total, 0.0);
    id_solve_cta(Mo, *desc, T, L, q, v, sm.data(), tau, a, f, info, ws);
    return 0;
  }
  catch (const std::exception &)
  {
    return -1;
  }
}

int emul_id_ws_doubles(const smpc_model_desc * model, const smpc_id_desc * desc) { return id_make_dims(*model, *desc).ws_stride; }
int emul_id_smem_bytes(const smpc_model_desc * model, const smpc_id_desc * desc) { return id_make_dims(*model, *desc).total * 8; }

} // extern "C"
