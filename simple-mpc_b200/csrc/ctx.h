// ctx.h — what every translation unit of libsmpc_cuda shares: the device array table passed to the kernels, the
// context struct behind the opaque smpc_ctx handle, and the per-shape setup entry points. The library is built from
// several .cu files (one per kernel family and robot shape) so that they compile in parallel; no device code is
// called across translation units.
#pragma once
#include "kino_knot.cuh" // KnotIO
#include "riccati.cuh"
#include "riccati_par_layout.h"
#include "riccati_fwd.cuh"
#include "mpc_driver.cuh"
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace smpc
{
  std::string & err(); // text of the last error of the calling thread (smpc_last_error)


  enum Phase
  {
    PH_DONE = 0,
    PH_ACTIVE = 1,    // needs a linearisation
    PH_NEED_STEP = 2, // linearised, Newton step wanted
    PH_LS = 3         // step computed, linesearch pending
  };

  struct Dims
  {
    int B, N, nx, ndx, nu, ncmax, nk;
  };

  // device array pointers, passed by value to the kernels
  struct Arrays
  {
    Dims dm;
    KnotLayout L;
    SolLayout SL;
    RicSmem RS;
    FwdSmem FS;
    const DevModel * model;
    const DevOcp * ocp;
    smpc_solver_opts opt;
    smpc_slot * slots;   // [B][N] ring buffer, head in InstState
    smpc_slot * table;   // stage table (the cycle horizon of MPC::generateCycleHorizon), ntable entries
    int ntable;
    smpc_terminal * term; // [B]
    double *x0, *xs, *us, *vs, *lams, *prev_vs, *prev_lams, *txs, *tus;
    double * knots;
    int * ncs;
    double *sol, *P0;
    double *gx, *gu, *lplus;
    double *dxs, *dus, *vnew, *lnew;
    KnotScalars *sc, *sct;
    double *xdot, *xdott;
    double *fc, *fct; // contact forces of the last evaluation / trial [B][N][nk * 3] (full dynamics; else null)
    InstState * inst;
    int * counters; // [0] linesearch pending, [1] instances that need another round, [2] driver queue overflows
    // batched MPC driver (mpc_driver.cuh); null until smpc_mpc_setup
    MpcShared * mpcS;
    MpcInst * mpcI;
    int * mpc_ids;  // [B] stage-table index of the tail each instance appends this tick
    double * xmeas; // [B][nq + nv] measured multibody states
    int nxm, ncycle;
    // PARALLEL (leg-condensed) Riccati sweep, riccati_par.cuh; legs <= 1: serial sweep, pointers null
    int legs;
    ParSmem PS;
    CoarseSmem CS;
    double *legP0, *legT; // [B][legs][ndx (ndx + 1)]: (P_a, p_a) of the zero-terminal sweeps / true cost-to-go at the leg starts
    double * legE;        // [B][legs][leg_elem_size]: Pi, Lam, psi of the inner legs
    const double * legZero; // ndx (ndx + 1) zeros: the cost-to-go of pass 1
    int * legFail;          // [B]: a factorisation of pass 2 / 3 met a non-positive pivot -> pass 4 runs the serial sweep instead
    int leg_force_fail;     // test hook (SMPC_LEGS_FORCE_FALLBACK): every instance takes that fallback
  };

  __device__ __forceinline__ const smpc_slot & slot_at(const Arrays & A, const InstState & st, int b, int t)
  {
    return A.slots[(size_t)b * A.dm.N + (st.head + t) % A.dm.N];
  }

  // global-memory view of knot (b, t) for the linearisation (TRIAL = false) or a linesearch trial
  template <bool TRIAL>
  __device__ __forceinline__ void make_knot_io(const Arrays & A, const InstState & st, int b, int t, KnotIO & io, int * nc_dummy)
  {
    const Dims & d = A.dm;
    const size_t xb = (size_t)b * (d.N + 1);
    io.x = A.xs + (xb + t) * d.nx;
    io.u = (t < d.N) ? A.us + ((size_t)b * d.N + t) * d.nu : nullptr;
    io.xn = (t < d.N) ? A.xs + (xb + t + 1) * d.nx : nullptr;
    if (TRIAL)
    {
      io.dx = A.dxs + (xb + t) * d.ndx;
      io.du = (t < d.N) ? A.dus + ((size_t)b * d.N + t) * d.nu : nullptr;
      io.dxn = (t < d.N) ? A.dxs + (xb + t + 1) * d.ndx : nullptr;
      io.alpha = st.alpha;
      io.tx = A.txs + (xb + t) * d.nx;
      io.tu = (t < d.N) ? A.tus + ((size_t)b * d.N + t) * d.nu : nullptr;
    }
    else
    {
      io.dx = io.du = io.dxn = nullptr;
      io.alpha = 0;
      io.tx = io.tu = nullptr;
    }
    io.slot = (t < d.N) ? &slot_at(A, st, b, t) : nullptr;
    io.term = A.term + b;
    io.vs = A.vs + (xb + t) * d.ncmax;
    io.prev_vs = A.prev_vs + (xb + t) * d.ncmax;
    io.lam_t = A.lams + (xb + t) * d.ndx;
    io.lam_n = (t < d.N) ? A.lams + (xb + t + 1) * d.ndx : nullptr;
    io.prev_lam_n = (t < d.N) ? A.prev_lams + (xb + t + 1) * d.ndx : nullptr;
    io.mu = st.mu;
    io.preg = st.preg;
    io.knot = A.knots + (xb + t) * A.L.size;
    io.nc_out = TRIAL ? nc_dummy : A.ncs + xb + t;
    io.gx = A.gx + (xb + t) * d.ndx;
    io.gu = (t < d.N) ? A.gu + ((size_t)b * d.N + t) * d.nu : nullptr;
    io.lplus_n = (t < d.N) ? A.lplus + (xb + t + 1) * d.ndx : nullptr;
    io.xdot = (t < d.N) ? (TRIAL ? A.xdott : A.xdot) + ((size_t)b * d.N + t) * d.ndx : nullptr;
    io.fc = (t < d.N && A.fc) ? (TRIAL ? A.fct : A.fc) + ((size_t)b * d.N + t) * d.nk * 3 : nullptr;
    io.sc = (TRIAL ? A.sct : A.sc) + xb + t;
  }
} // namespace smpc

struct smpc_ctx
{
  typedef smpc::Arrays Arrays;
  typedef smpc::DevModel DevModel;
  typedef smpc::DevOcp DevOcp;
  typedef smpc::MpcShared MpcShared;
  int device = 0;
  cudaStream_t stream = nullptr;
  Arrays A{};
  DevModel hmodel;
  DevOcp hocp;
  int hmodel_sms = 0; // SM count of the device
  bool forward_staged = true;
  std::vector<void *> allocs;
  int * h_counters = nullptr; // pinned
  void * staging = nullptr;   // device staging for host<->device conversions
  size_t staging_bytes = 0;
  void * h_staging = nullptr; // pinned host staging
  size_t h_staging_bytes = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evr0 = nullptr, evr1 = nullptr;
  cudaEvent_t evk[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // first round of a solve: lin | decide+riccati | forward | first trial
  double acc_kernel_ms[4] = {0, 0, 0, 0};                                      // linearisation, Riccati backward, forward sweep, first trial (+ ls_decide)
  float last_ms = 0, last_ric_ms = 0;
  int ric_generic_tpb = 128; // CTA width of the run-time-shape Riccati sweep (riccati.cuh)
  double acc_ric_ms = 0, acc_solve_ms = 0; // accumulated since smpc_timer_reset
  int64_t acc_ric_launches = 0, acc_solves = 0;
  cudaEvent_t evu0 = nullptr, evu1 = nullptr; // user timer on the context stream
  int64_t launches = 0;
  size_t knot_smem = 0, ric_smem = 0;
  int knot_threads = 32;
  void (*launch_knots)(smpc_ctx *, bool trial) = nullptr;
  void (*launch_riccati)(smpc_ctx *) = nullptr;
  // pass 1 / pass 4 of the leg-parallel sweep with the tuned kernel of this shape (null: the shape has no tuned sweep)
  void (*launch_riccati_legs)(smpc_ctx *, int pass) = nullptr;
  const char * riccati_variant = "generic";
  std::vector<smpc_slot> h_table; // host copy of the stage table (contact masks drive the driver's bookkeeping)
  MpcShared h_mpc;
  bool mpc_ready = false;
  // CUDA Graph of the normal MPC tick (smpc_mpc_use_graph): driver kernels + one round + its first linesearch trial
  int use_graph = 0;
  cudaGraphExec_t tick_graph = nullptr;
  Arrays tick_graph_A{};   // kernel arguments the graph was captured with (re-captured when they change)
  int tick_graph_iters = 0; // max_iters baked into k_run_begin's node
  int tick_graph_fwd = 0;   // forward-sweep variant baked in
  int64_t graph_ticks = 0;
  // small batches: the graph also copies the head of the plan (x1, u0, u1, K0 of every instance) into this pinned buffer, so
  // the smpc_get_head that follows a tick is a host memcpy without a second synchronisation
  double * h_head = nullptr;
  bool head_cached = false;
};

#define CU(call)                                                                                                       \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t e_ = (call);                                                                                           \
    if (e_ != cudaSuccess)                                                                                             \
    {                                                                                                                  \
      smpc::err() = std::string(#call) + ": " + cudaGetErrorString(e_);                                                      \
      return SMPC_ERR_CUDA;                                                                                            \
    }                                                                                                                  \
  } while (0)


// kernel instantiations live in their own translation units; each returns SMPC_OK after binding the context's
// launch_knots / launch_riccati function pointers, SMPC_ERR_UNSUPPORTED (via *found = false) if the shape is not its own
namespace smpc
{
  int setup_knots_kino_go2(smpc_ctx * c);    // KinoCfg<13, 18, 4, 3>            knots_kino_go2.cu
  int setup_knots_kino_talos(smpc_ctx * c);  // KinoCfg<23, 28, 2, 6>            knots_kino_talos.cu
  int setup_knots_full_go2(smpc_ctx * c);    // FullCfg<13, 18, 4>               knots_full_go2.cu
  int setup_knots_full_talos(smpc_ctx * c);  // FullCfg<23, 28, 2, 6>            knots_full_talos.cu
  int setup_knots_centro(smpc_ctx * c);      // CentroCfg<2, 6>, CentroCfg<4, 3> knots_centro.cu
  // shape-specialised Riccati sweeps: *found = false when the shape is not instantiated in that translation unit
  int setup_ric3_go2(smpc_ctx * c, int ndx, int nu, int ncmax, bool * found); // riccati3.cuh, ric3_go2.cu
  int setup_ric4_go2(smpc_ctx * c, int ndx, int nu, int ncmax, bool * found); // riccati4.cuh, ric4_go2.cu
  int setup_ric4_big(smpc_ctx * c, int ndx, int nu, int ncmax, bool * found); // riccati4.cuh, ric4_big.cu
  int setup_ric4_talos_kino(smpc_ctx * c, int ndx, int nu, int ncmax, bool * found); // riccati4.cuh, ric4_talos_kino.cu
  int setup_ric4_centro(smpc_ctx * c, int ndx, int nu, int ncmax, bool * found); // riccati4.cuh padded mode, ric4_centro.cu
} // namespace smpc
