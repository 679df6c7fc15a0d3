// dev_types.h — POD tables shared by the host side of libsmpc_cuda and its kernels.
#pragma once
#include "../../include/smpc.h"

namespace smpc
{

  // smpc_model_desc plus the derived tables the kernels index (built once on the host).
  struct DevModel
  {
    smpc_model_desc d;
    int nlevels;
    int level_start[SMPC_MAX_JOINTS + 1]; // joints of depth l: level_joints[level_start[l] .. level_start[l+1])
    int level_joints[SMPC_MAX_JOINTS];
    int joint_of_dof[SMPC_MAX_NV];
    unsigned anc_mask[SMPC_MAX_JOINTS]; // bit j: joint j is an ancestor-or-self of joint i
    double total_mass;
  };

  // The problem recipe with owned storage.
  struct DevOcp
  {
    int kind, N, fs, nk;
    int nx, ndx, nu, ncmax, nct; // nct: terminal constraint rows
    int kinematics_limits, torque_limits, force_cone, land_cstr, terminal_constraint;
    double dt, gravity[3], mu, Lfoot, Wfoot, mass;
    double w_x[SMPC_MAX_NDX]; // diagonal of w_x (non-diagonal w_x is rejected at create)
    double w_u[SMPC_MAX_NU];  // diagonal of w_u
    double w_cent[36], w_centder[36], w_frame[36];
    double w_com[9], w_lmom[9], w_amom[9], w_lacc[9], w_aacc[9];
    double qmin[SMPC_MAX_NV], qmax[SMPC_MAX_NV];
    double wcone[17 * 6]; // wrench-cone rows A w <= 0 for 6-D contacts (mu, Lfoot, Wfoot)
    // full dynamics (reference src/fulldynamics.cpp)
    double w_forces[36];
    double umin[SMPC_MAX_NV], umax[SMPC_MAX_NV];
    double Kp[6], Kd[6];  // Baumgarte corrector of the rigid contacts
    double pyramid[5 * 3]; // friction pyramid rows A lam <= 0 (aligator MultibodyFrictionConeResidual [UPSTREAM])
  };

  // Dense LQ knot layout in HBM (offsets in doubles from the knot base), row-major blocks.
  struct KnotLayout
  {
    int ndx, nu, ncmax;
    int oQ, oS, oR, oq, or_, oA, oB, of, oC, oD, od, size;
  };
  inline KnotLayout make_knot_layout(int ndx, int nu, int ncmax)
  {
    KnotLayout L;
    L.ndx = ndx;
    L.nu = nu;
    L.ncmax = ncmax;
    int o = 0;
    L.oQ = o, o += ndx * ndx;
    L.oS = o, o += ndx * nu;
    L.oR = o, o += nu * nu;
    L.oq = o, o += ndx;
    L.or_ = o, o += nu;
    L.oA = o, o += ndx * ndx;
    L.oB = o, o += ndx * nu;
    L.of = o, o += ndx;
    L.oC = o, o += ncmax * ndx;
    L.oD = o, o += ncmax * nu;
    L.od = o, o += ncmax;
    L.size = (o + 1) & ~1; // keep every knot 16-byte aligned
    return L;
  }

  // Riccati solution layout per stage: K nu*ndx, k nu, Z ncmax*ndx, z ncmax, Pt ndx*ndx, pt ndx
  struct SolLayout
  {
    int oK, ok, oZ, oz, oPt, opt, size;
  };
  inline SolLayout make_sol_layout(int ndx, int nu, int ncmax)
  {
    SolLayout L;
    int o = 0;
    L.oK = o, o += nu * ndx;
    L.ok = o, o += nu;
    L.oZ = o, o += ncmax * ndx;
    L.oz = o, o += ncmax;
    L.oPt = o, o += ndx * ndx;
    L.opt = o, o += ndx;
    L.size = (o + 1) & ~1;
    return L;
  }

  // Per-knot scalars produced by the linearisation / trial evaluation kernels.
  struct KnotScalars
  {
    double cost;   // stage cost
    double pen;    // AL penalty terms of this knot (dynamics + constraints)
    double pinf;   // primal infeasibility (inf-norm pieces)
    double dinf;   // dual infeasibility (|Lx|, |Lu|)
    double crit;   // inner criterion pieces
    double dphi;   // unused by the knot kernels (filled by forward)
    double pad[2];
  };

  // Per-instance solver state (aligator results_/workspace_ scalars that persist between runs).
  struct InstState
  {
    double mu, preg, inner_tol, prim_tol;
    double ls_mov_avg, ls_avg_weight;
    double phi0, dphi0, alpha, phi_trial;
    double cost0;
    int phase;     // see solver_kernels.cu
    int num_iters, al_iters, conv, ls_trials, flags;
    int max_iters;
    int head;      // ring-buffer head of the slot table
    smpc_solve_info info;
  };

} // namespace smpc
