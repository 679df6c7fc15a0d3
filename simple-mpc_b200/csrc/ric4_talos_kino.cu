// generic-shape two-column Riccati sweeps (riccati4.cuh) for Talos kinodynamics (its own translation unit: each of these
// instantiations compiles for about a minute)
#include "riccati4_kernel.cuh"
namespace smpc
{
  int setup_ric4_talos_kino(smpc_ctx * c, int ndx, int nu, int ncmax, bool * found)
  {
    *found = true;
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
