// kinodynamics knots, 22-DoF biped with two quad feet (reduced-Talos shape: nv 28, 6-D contact wrenches as controls,
// CentroidalWrenchCone + 6-D LOCAL contact velocity rows, FramePlacement pose cost; reference src/kinodynamics.cpp:40-152)
#include "knot_kernels.cuh"
namespace smpc
{
  int setup_knots_kino_talos(smpc_ctx * c) { return setup_knots<KinoCfg<23, 28, 2, 6>>(c); }
} // namespace smpc
