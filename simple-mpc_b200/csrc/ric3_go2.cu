// TMA-staged Riccati sweeps for the Go2-kinodynamics shapes (ndx 36, nu 24; ncmax 24 without / 36 with cone + landing rows)
#include "riccati3_kernel.cuh"
namespace smpc
{
  int setup_riccati_tuned(smpc_ctx * c, int ndx, int nu, int ncmax, bool * found)
  {
    *found = true;
    if (ndx == 36 && nu == 24 && ncmax == 24)
      return setup_riccati3<36, 24, 24>(c);
    if (ndx == 36 && nu == 24 && ncmax == 36)
      return setup_riccati3<36, 24, 36>(c);
    *found = false;
    return SMPC_OK;
  }
} // namespace smpc
