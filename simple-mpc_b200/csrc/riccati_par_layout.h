// riccati_par_layout.h — leg boundaries and buffer layouts of the PARALLEL (leg-condensed) Riccati sweep (riccati_par.cuh).
// Kept apart from the kernels so that every translation unit can see the Arrays fields without recompiling when the
// element / condensation kernels change.
#pragma once
#include "par.cuh"

namespace smpc
{
  // leg boundaries a_j = floor(j N / J)
  SMPC_HD int leg_begin(int N, int J, int j) { return (int)(((long long)j * N) / J); }
  SMPC_HD int riccati_legs_used(int N, int J) { return (J >= 2 && N >= 2 * J) ? J : 1; }
  // per-leg records in HBM: [P | p] blocks use the layout of P0out (g rows of g + 1 columns, p last)
  SMPC_HD int leg_pp_size(int g) { return g * (g + 1); }
  SMPC_HD int leg_elem_size(int g) { return 2 * g * g + g; } // Pi [g][g] | Lam [g][g] | psi [g]

  struct ParSmem
  {
    int oPi, oPit, oLam, oY, oX, oRh, oW, oSg, oG, ov, ophi, opsi, oinvd, osPt, osA, osB, osK, oflag, size;
  };
  inline ParSmem make_par_smem(int ndx, int nu, int ncmax)
  {
    ParSmem R;
    int o = 0;
    R.oPi = o, o += ndx * ndx;
    R.oPit = o, o += ndx * ndx;
    R.oLam = o, o += ndx * ndx;
    R.oY = o, o += nu * ndx;
    R.oX = o, o += (nu * ndx > ndx * nu ? nu * ndx : ndx * nu); // P~ B (ndx x nu) before X (nu x ndx) exists
    R.oRh = o, o += nu * nu;
    R.oW = o, o += nu * ncmax;
    R.oSg = o, o += ncmax * ncmax;
    R.oG = o, o += ncmax * ndx;
    R.ov = o, o += ndx;
    R.ophi = o, o += ndx;
    R.opsi = o, o += ndx;
    R.oinvd = o, o += nu + ncmax; // inverse pivots of R^ and of Sigma
    o += o & 1;
    // the stage's operands from HBM / L2 (P~, A, B, K), staged once per stage: every product then reads shared memory only
    R.osPt = o, o += ndx * ndx;
    R.osA = o, o += ndx * ndx;
    R.osB = o, o += ndx * nu;
    R.osK = o, o += nu * ndx;
    R.oflag = o, o += 2;
    R.size = o;
    return R;
  }
  // the condensation step works on (ndx x ndx) blocks only
  struct CoarseSmem
  {
    int oC, oLC, oS, oR, oU, oPn, owv, oflag, size;
  };
  inline CoarseSmem make_coarse_smem(int ndx)
  {
    CoarseSmem R;
    int o = 0;
    R.oC = o, o += ndx * ndx;
    R.oLC = o, o += ndx * ndx;
    R.oS = o, o += ndx * ndx;
    R.oR = o, o += ndx * (ndx + 1);
    R.oU = o, o += ndx * (ndx + 1);
    R.oPn = o, o += ndx * (ndx + 1);
    R.owv = o, o += ndx + (ndx & 1);
    R.oflag = o, o += 2;
    R.size = o;
    return R;
  }
} // namespace smpc
