// generic-shape two-column Riccati sweeps (riccati4.cuh) for the shapes that fill an SM with one instance: Go2 full dynamics
// with friction pyramids + landing rows (36, 12, 60) and Talos full dynamics with wrench cones (56, 22, 78) / without (56, 22, 44), Talos kinodynamics (56, 34, 34 | 46)
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
    // Talos kinodynamics (nu = 22 joint accelerations + 2 x 6 contact wrenches): joint limits + 6-D contact velocity rows (the
    // reference example, 34 rows) / wrench cones + velocity rows without joint limits (46 rows). With every row family
    // (68 rows) the stage exceeds 227 KB of shared memory and runs on the run-time-shape sweep of riccati.cuh
    if (ndx == 56 && nu == 34 && ncmax == 34)
      return setup_riccati4<56, 34, 34, 256>(c, "two-column-256");
    if (ndx == 56 && nu == 34 && ncmax == 46)
      return setup_riccati4<56, 34, 46, 256>(c, "two-column-256");
    *found = false;
    return SMPC_OK;
  }
} // namespace smpc
