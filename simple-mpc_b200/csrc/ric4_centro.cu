// two-column Riccati sweeps (riccati4.cuh, padded mode) for the centroidal problem: state dimension 9 computed on 12 x 12 blocks
// with exact-zero padding. Shapes: biped with two quad feet and wrench cones (nu 12, 2 x 17 rows; config C2, reference
// examples/talos_centroidal.py) and quadruped with four point feet and friction cones (nu 12, 4 x 2 rows)
#include "riccati4_kernel.cuh"
namespace smpc
{
  int setup_ric4_centro(smpc_ctx * c, int ndx, int nu, int ncmax, bool * found)
  {
    *found = true;
    if (ndx == 9 && nu == 12 && ncmax == 34)
      return setup_riccati4<12, 12, 34, 128, 9, true>(c, "two-column-padded");
    if (ndx == 9 && nu == 12 && ncmax == 8)
      return setup_riccati4<12, 12, 8, 128, 9, true>(c, "two-column-padded");
    *found = false;
    return SMPC_OK;
  }
} // namespace smpc
