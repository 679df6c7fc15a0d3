///////////////////////////////////////////////////////////////////////////////
// solver-api.hpp — the table of C-ABI entry points (include/smpc.h) the host classes call.
//
// SolverApi::cuda() binds the table to libsmpc_cuda (the hand-written sm_100a kernels); that is the only
// table the product ever constructs, and smpc_create fails with SMPC_ERR_CUDA when there is no device —
// there is no CPU fallback. The table exists so that the TEST harness can drive the very same host code
// against the CPU oracle (identical signatures, `orc_` prefix) and compare the two end to end.
///////////////////////////////////////////////////////////////////////////////
#pragma once
#include "smpc.h"

namespace simple_mpc
{

  struct SolverApi
  {
    const char * (*last_error)(void);
    void (*default_solver_opts)(smpc_solver_opts *);
    int (*problem_dims)(const smpc_model_desc *, const smpc_ocp_desc *, int *, int *, int *);
    /// create(out, device, batch, model, ocp, opts)
    int (*create)(smpc_ctx **, int, int, const smpc_model_desc *, const smpc_ocp_desc *, const smpc_solver_opts *);
    void (*destroy)(smpc_ctx *);
    int (*set_slots)(smpc_ctx *, int, int, const smpc_slot *);
    int (*set_terminal)(smpc_ctx *, int, int, const smpc_terminal *);
    int (*cycle)(smpc_ctx *, int, int, const smpc_slot *);
    int (*set_warmstart)(smpc_ctx *, int, int, const double *, const double *);
    int (*solve)(smpc_ctx *, const double *, int);
    int (*get_results)(smpc_ctx *, int, int, double *, double *, double *, double *);
    int (*get_info)(smpc_ctx *, int, int, smpc_solve_info *);
    int (*get_xdot)(smpc_ctx *, int, int, int, double *);
    int (*get_contact_forces)(smpc_ctx *, int, int, int, double *);
    const char * name;

    /// the CUDA library (the product path)
    static const SolverApi & cuda();
  };

} // namespace simple_mpc
