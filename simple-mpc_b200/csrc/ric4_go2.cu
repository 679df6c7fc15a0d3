// generic-shape two-column Riccati sweeps (riccati4.cuh) for the Go2 shapes: kinodynamics (ndx 36, nu 24; ncmax 24 without /
// 36 with cone + landing rows) and full dynamics (ndx 36, nu 12; ncmax 24 with torque + joint limits)
#include "riccati4_kernel.cuh"
namespace smpc
{
  // 128 threads: throughput variant (three CTAs per SM); 256: latency variant when no SM holds more than one instance
  template <int NDX, int NU, int NC>
  static int pick(smpc_ctx * c)
  {
    bool wide = c->hmodel_sms > 0 && c->A.dm.B <= c->hmodel_sms;
    if (const char * e = getenv("SMPC_RIC_TPB"))
      wide = atoi(e) == 256;
    return wide ? setup_riccati4<NDX, NU, NC, 256, NDX, (NU == 12)>(c, "two-column-256") : setup_riccati4<NDX, NU, NC, 128>(c, "two-column");
  }
  int setup_ric4_go2(smpc_ctx * c, int ndx, int nu, int ncmax, bool * found)
  {
    *found = true;
    if (ndx == 36 && nu == 24 && ncmax == 24)
      return pick<36, 24, 24>(c);
    if (ndx == 36 && nu == 24 && ncmax == 36)
      return pick<36, 24, 36>(c);
    if (ndx == 36 && nu == 12 && ncmax == 24)
      return pick<36, 12, 24>(c);
    *found = false;
    return SMPC_OK;
  }
} // namespace smpc

#ifdef SMPC_RIC_PROFILE
// debug builds only: g_ric_prof is a per-translation-unit symbol; this reads the copy the riccati4 Go2 kernels write
extern "C" int smpc_debug_ric4_profile(unsigned long long * out16, int reset)
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
