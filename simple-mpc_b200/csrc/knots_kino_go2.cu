// kinodynamics knots, 12-DoF quadruped with point feet (Go2 / Solo12 shape)
#include "knot_kernels.cuh"
namespace smpc
{
  int setup_knots_kino_go2(smpc_ctx * c) { return setup_knots<KinoCfg<13, 18, 4, 3>>(c); }
} // namespace smpc
