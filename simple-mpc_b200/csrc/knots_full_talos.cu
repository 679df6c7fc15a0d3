// full-dynamics knots, 22-DoF biped with two quad feet (reduced-Talos shape: nv 28, 6-D contacts) - config C4
#include "knot_kernels.cuh"
namespace smpc
{
  int setup_knots_full_talos(smpc_ctx * c) { return setup_knots_full<FullCfg<23, 28, 2, 6>, KinoCfg<23, 28, 2, 6>>(c); }
} // namespace smpc
