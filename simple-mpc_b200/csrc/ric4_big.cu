// generic-shape two-column Riccati sweeps (riccati4.cuh) for the shapes that fill an SM with one instance: Go2 full dynamics
// with friction pyramids + landing rows (36, 12, 60) and Talos full dynamics with wrench cones (56, 22, 78) / without (56, 22, 44)
#include "riccati4_kernel.cuh"
namespace smpc
{
  int setup_ric4_big(smpc_ctx * c, int ndx, int nu, int ncmax, bool * found)
  {
    *found = true;
    if (ndx == 36 && nu == 12 && ncmax == 60)
      return setup_riccati4<36, 12, 60, 256>(c, "two-column-256");
    if (ndx == 56 && nu == 22 && ncmax == 78)
      return setup_riccati4<56, 22, 78, 256>(c, "two-column-256");
    if (ndx == 56 && nu == 22 && ncmax == 44)
      return setup_riccati4<56, 22, 44, 256>(c, "two-column-256");
    *found = false;
    return SMPC_OK;
  }
} // namespace smpc
