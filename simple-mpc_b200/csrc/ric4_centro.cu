// two-column Riccati sweeps (riccati4.cuh, padded mode) for the centroidal problem: state dimension 9 computed on 12 x 12 blocks
// with exact-zero padding. Shapes: biped with two quad feet and wrench cones (nu 12, 2 x 17 rows; config C2, reference
// examples/talos_centroidal.py) and quadruped with four point feet and friction cones (nu 12, 4 x 2 rows)
#include "riccati4_kernel.cuh"
namespace smpc
{
  // 128 threads: throughput variant; 256: latency variant when no SM holds more than one instance (SMPC_RIC_TPB overrides)
  template <int NC>
  static int pick(smpc_ctx * c)
  {
    bool wide = c->hmodel_sms > 0 && c->A.dm.B <= c->hmodel_sms;
    if (const char * e = getenv("SMPC_RIC_TPB"))
      wide = atoi(e) == 256;
    return wide ? setup_riccati4<12, 12, NC, 256, 9, true>(c, "two-column-padded-256") : setup_riccati4<12, 12, NC, 128, 9, true>(c, "two-column-padded");
  }
  int setup_ric4_centro(smpc_ctx * c, int ndx, int nu, int ncmax, bool * found)
  {
    *found = true;
    if (ndx == 9 && nu == 12 && ncmax == 34)
      return pick<34>(c);
    if (ndx == 9 && nu == 12 && ncmax == 8)
      return pick<8>(c);
    *found = false;
    return SMPC_OK;
  }
} // namespace smpc
