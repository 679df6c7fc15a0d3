// centroidal knots: two quad feet (6-D contacts) or four point feet
#include "knot_kernels.cuh"
namespace smpc
{
  int setup_knots_centro(smpc_ctx * c)
  {
    if (c->hocp.nk == 2 && c->hocp.fs == 6)
      return setup_knots_centro<CentroCfg<2, 6>>(c);
    return setup_knots_centro<CentroCfg<4, 3>>(c);
  }
} // namespace smpc
