// TMA-staged Riccati sweeps for the Go2-kinodynamics shapes (ndx 36, nu 24; ncmax 24 without / 36 with cone + landing rows)
#include "riccati3_kernel.cuh"
namespace smpc
{
  int setup_ric3_go2(smpc_ctx * c, int ndx, int nu, int ncmax, bool * found)
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

#ifdef SMPC_RIC_PROFILE
// debug builds only (make SMPC_DEFS=-DSMPC_RIC_PROFILE): per-phase clock64() totals of the Riccati kernel, see tools/ric_phases.py
extern "C" int smpc_debug_ric_profile(unsigned long long * out16, int reset)
{
  if (cudaMemcpyFromSymbol(out16, g_ric_prof, 16 * sizeof(unsigned long long)) != cudaSuccess)
    return -1;
  if (reset)
  {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_ric_prof, z, sizeof(z));
  }
  return 0;
}
#endif
