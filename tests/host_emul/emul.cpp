// TEST HARNESS — not part of the product. Compiles the per-knot / per-instance device routines of
// simple-mpc_b200/csrc/*.cuh with g++ (par.cuh turns each parallel phase into a serial loop over the
// thread index) so that kernel LOGIC can be debugged against the oracle on the GPU-less development
// box. The shipped library (libsmpc_cuda.so) never contains or calls this code.
#include "../../simple-mpc_b200/csrc/kino_knot.cuh"
#include "../../simple-mpc_b200/csrc/centro_knot.cuh"
#include "../../simple-mpc_b200/csrc/full_knot.cuh"
#include "../../simple-mpc_b200/csrc/riccati.cuh"
#include "../../simple-mpc_b200/csrc/riccati3.cuh"
#include "../../simple-mpc_b200/csrc/riccati4.cuh"
#include "../../simple-mpc_b200/csrc/riccati_par.cuh"
#include "../../simple-mpc_b200/csrc/riccati_fwd.cuh"
#include "../../simple-mpc_b200/csrc/mpc_driver.cuh"
#include "../../simple-mpc_b200/csrc/host_build.h"
#include <memory>
#include <vector>
#include <string>

using namespace smpc;
static std::string g_err;
typedef KinoCfg<13, 18, 4, 3> Go2Cfg;

// one centroidal knot (2 quad feet or 4 point feet); same argument meaning as emul_kino_knot
template <class Cfg>
static void centro_knot_any(const DevOcp & O, const KnotLayout & L, const KnotIO & io, bool terminal, bool derivs)
{
  auto s = std::make_unique<CentroScratch<Cfg>>();
  memset(s.get(), 0, sizeof(*s));
  if (terminal)
  {
    if (derivs)
      centro_terminal_knot<Cfg, true>(O, L, io, *s);
    else
      centro_terminal_knot<Cfg, false>(O, L, io, *s);
  }
  else
  {
    if (derivs)
      centro_stage_knot<Cfg, true>(O, L, io, *s);
    else
      centro_stage_knot<Cfg, false>(O, L, io, *s);
  }
}

// one full-dynamics knot of shape FCfg (terminal node: the kinodynamics one of shape TCfg)
template <class FCfg, class TCfg>
static void full_knot_any(const DevModel & M, const DevOcp & O, const KnotLayout & L, const KnotIO & io, bool terminal, bool derivs)
{
  if (terminal)
  {
    auto s = std::make_unique<KinoScratch<TCfg>>();
    memset(s.get(), 0, sizeof(*s));
    if (derivs)
      kino_terminal_knot<TCfg, true>(M, O, L, io, *s);
    else
      kino_terminal_knot<TCfg, false>(M, O, L, io, *s);
  }
  else
  {
    auto s = std::make_unique<FullScratch<FCfg>>();
    memset(s.get(), 0, sizeof(*s));
    if (derivs)
      full_stage_knot<FCfg, true>(M, O, L, io, *s);
    else
      full_stage_knot<FCfg, false>(M, O, L, io, *s);
  }
}

// one kinodynamics knot of shape Cfg
template <class Cfg>
static void kino_knot_any(const DevModel & M, const DevOcp & O, const KnotLayout & L, const KnotIO & io, bool terminal, bool derivs)
{
  auto s = std::make_unique<KinoScratch<Cfg>>();
  memset(s.get(), 0, sizeof(*s));
  if (terminal)
  {
    if (derivs)
      kino_terminal_knot<Cfg, true>(M, O, L, io, *s);
    else
      kino_terminal_knot<Cfg, false>(M, O, L, io, *s);
  }
  else
  {
    if (derivs)
      kino_stage_knot<Cfg, true>(M, O, L, io, *s);
    else
      kino_stage_knot<Cfg, false>(M, O, L, io, *s);
  }
}

template <int NDX, int NU, int NC, int GN = NDX>
static void ric4_any(const KnotLayout & L, const SolLayout & SL, int N, const double * knots, double * sol, double mudyn, double mueq, int * flag, double * P0)
{
  std::vector<double> sm(Ric4<NDX, NU, NC>::SIZE + 16, 0.0);
  riccati4_backward_inst<NDX, NU, NC, SMPC_EMUL_NT, GN>(L, SL, N, knots, sol, mudyn, mueq, sm.data(), flag, P0);
}

// the TMA-staged sweep of riccati3.cuh (the one the library launches); SNV = 18: the kinodynamics-structure products
template <int SNV, bool DZ = false>
static int emul_riccati3_impl(
  int N, int ndx, int nu, int ncmax, const double * knots, const int * ncs, double mudyn, double mueq, double * sol,
  const double * gx, const double * gu, const double * lplus, double * dxs, double * dus, double * vnew, double * lnew,
  double * dphi, int * flag)
{
  KnotLayout L = make_knot_layout(ndx, nu, ncmax);
  SolLayout SL = make_sol_layout(ndx, nu, ncmax);
  std::vector<double> P0((size_t)ndx * (ndx + 1), 0.0), sm2(4 * ndx + nu + 80, 0.0);
  if (ndx == 36 && nu == 24 && ncmax == 24)
  {
    std::vector<double> sm(Ric3<36, 24, 24>::SIZE + 8, 0.0);
    riccati3_backward_inst<36, 24, 24, SMPC_EMUL_NT, SNV, DZ>(L, SL, N, knots, sol, mudyn, mueq, sm.data(), flag, P0.data());
  }
  else if (ndx == 36 && nu == 24 && ncmax == 36)
  {
    std::vector<double> sm(Ric3<36, 24, 36>::SIZE + 8, 0.0);
    riccati3_backward_inst<36, 24, 36, SMPC_EMUL_NT, SNV, DZ>(L, SL, N, knots, sol, mudyn, mueq, sm.data(), flag, P0.data());
  }
  else
  {
    g_err = "emul_riccati3: size not instantiated";
    return -1;
  }
  // the hot path's forward sweep: stage matrices staged through (emulated) shared memory
  FwdSmem FS = make_fwd_smem(ndx, nu, ncmax, getenv("SMPC_EMUL_FWD_NBUF") ? atoi(getenv("SMPC_EMUL_FWD_NBUF")) : 2);
  std::vector<double> smf((size_t)FS.size + 8, 0.0);
  riccati_forward_staged(L, SL, FS, N, knots, ncs, sol, P0.data(), mudyn, gx, gu, lplus, dxs, dus, vnew, lnew, dphi, smf.data());
  return 0;
}


// The PARALLEL (leg-condensed) sweep of riccati_par.cuh with J legs: pass 1 / pass 4 through the tuned sweeps (riccati3 for
// the Go2 kinodynamics shapes, riccati4 otherwise), leg elements, coarse condensation, then the ordinary forward sweep.
// The legs of a pass run one after the other here; on the device they are concurrent CTAs.
template <int NDX, int NU, int NC, bool R3>
static void leg_sweep(const KnotLayout & L, const SolLayout & SL, int n, const double * knots, double * sol, double mudyn, double mueq, int * flag, double * P0, const double * Pterm)
{
  if constexpr (R3)
  {
    std::vector<double> sm(Ric3<NDX, NU, NC>::SIZE + 8, 0.0);
    riccati3_backward_inst<NDX, NU, NC, SMPC_EMUL_NT, 0, false>(L, SL, n, knots, sol, mudyn, mueq, sm.data(), flag, P0, Pterm);
  }
  else
  {
    std::vector<double> sm(Ric4<NDX, NU, NC>::SIZE + 16, 0.0);
    riccati4_backward_inst<NDX, NU, NC, SMPC_EMUL_NT, NDX>(L, SL, n, knots, sol, mudyn, mueq, sm.data(), flag, P0, Pterm);
  }
}
template <int NDX, int NU, int NC, bool R3>
static int parallel_any(
  int N, int J, const double * knots, const int * ncs, double mudyn, double mueq, double * sol, const double * gx, const double * gu,
  const double * lplus, double * dxs, double * dus, double * vnew, double * lnew, double * dphi, int * flag)
{
  const int ndx = NDX, nu = NU, ncmax = NC;
  KnotLayout L = make_knot_layout(ndx, nu, ncmax);
  SolLayout SL = make_sol_layout(ndx, nu, ncmax);
  J = riccati_legs_used(N, J);
  if (J < 2)
  {
    g_err = "emul_riccati_parallel: fewer than two legs";
    return -1;
  }
  const size_t pp = leg_pp_size(ndx), es = leg_elem_size(ndx);
  std::vector<double> legP0(J * pp, 0.0), legT(J * pp, 0.0), legE(J * es, 0.0), zero(pp, 0.0), P0(pp, 0.0);
  int fl = 0;
  *flag = 0;
  for (int j = 1; j < J; j++) // pass 1
  {
    const int a0 = leg_begin(N, J, j), a1 = leg_begin(N, J, j + 1);
    leg_sweep<NDX, NU, NC, R3>(L, SL, a1 - a0, knots + (size_t)a0 * L.size, sol + (size_t)a0 * SL.size, mudyn, mueq, &fl, &legP0[j * pp], j == J - 1 ? nullptr : zero.data());
    *flag |= fl;
  }
  ParSmem PS = make_par_smem(ndx, nu, ncmax);
  for (int j = 1; j < J - 1; j++) // pass 2
  {
    std::vector<double> sm(PS.size + 8, 0.0);
    leg_element_inst(L, SL, PS, leg_begin(N, J, j), leg_begin(N, J, j + 1), knots, ncs, sol, mudyn, mueq, sm.data(), &fl, &legE[j * es]);
    *flag |= fl;
  }
  {
    CoarseSmem CS = make_coarse_smem(ndx); // pass 3
    std::vector<double> sm(CS.size + 8, 0.0);
    leg_coarse_inst(ndx, J, CS, legP0.data(), legE.data(), legT.data(), sm.data(), &fl);
    *flag |= fl;
  }
  for (int j = 0; j < J - 1; j++) // pass 4
  {
    const int a0 = leg_begin(N, J, j), a1 = leg_begin(N, J, j + 1);
    leg_sweep<NDX, NU, NC, R3>(L, SL, a1 - a0, knots + (size_t)a0 * L.size, sol + (size_t)a0 * SL.size, mudyn, mueq, &fl, j == 0 ? P0.data() : &legP0[j * pp], &legT[(j + 1) * pp]);
    *flag |= fl;
  }
  std::vector<double> sm2(4 * ndx + nu + 80, 0.0);
  riccati_forward_inst(L, SL, N, knots, ncs, sol, P0.data(), mudyn, gx, gu, lplus, dxs, dus, vnew, lnew, dphi, sm2.data());
  return 0;
}

extern "C" {
const char * emul_last_error() { return g_err.c_str(); }

// one kinodynamics knot (Go2 sizes with point feet, or talos-like sizes with quad feet). knot buffer must hold KnotLayout(ndx,nu,ncmax).size doubles.
int emul_kino_knot(
  const smpc_model_desc * model, const smpc_ocp_desc * ocp, const smpc_slot * slot, const smpc_terminal * term, int terminal,
  int derivs, const double * x, const double * u, const double * xn, const double * vs, const double * prev_vs,
  const double * lam_t, const double * lam_n, const double * prev_lam_n, double mu, double preg, double * knot, int * nc,
  double * gx, double * gu, double * lplus, double * xdot, double * scalars /*8*/)
{
  try
  {
    auto M = std::make_unique<DevModel>();
    auto O = std::make_unique<DevOcp>();
    build_dev_model(*model, *M);
    build_dev_ocp(*model, *ocp, *O);
    const bool go2 = model->njoints == 13 && model->nv == 18 && model->nfeet == 4 && O->fs == 3;
    const bool talos = model->njoints == 23 && model->nv == 28 && model->nfeet == 2 && O->fs == 6;
    if (!go2 && !talos)
      throw std::runtime_error("emul: kinodynamics instantiations: Go2-sized (point feet) and talos-sized (quad feet)");
    KnotLayout L = make_knot_layout(O->ndx, O->nu, O->ncmax);
    KnotIO io;
    memset(&io, 0, sizeof(io));
    io.x = x, io.u = u, io.xn = xn;
    io.slot = slot, io.term = term;
    io.vs = vs, io.prev_vs = prev_vs, io.lam_t = lam_t, io.lam_n = lam_n, io.prev_lam_n = prev_lam_n;
    io.mu = mu, io.preg = preg;
    io.knot = knot, io.nc_out = nc, io.gx = gx, io.gu = gu, io.lplus_n = lplus, io.xdot = xdot;
    KnotScalars sc;
    memset(&sc, 0, sizeof(sc));
    io.sc = &sc;
    if (go2)
      kino_knot_any<Go2Cfg>(*M, *O, L, io, terminal != 0, derivs != 0);
    else
      kino_knot_any<KinoCfg<23, 28, 2, 6>>(*M, *O, L, io, terminal != 0, derivs != 0);
    memcpy(scalars, &sc, sizeof(sc));
  }
  catch (const std::exception & e)
  {
    g_err = e.what();
    return -1;
  }
  return 0;
}

// one full-dynamics knot (Go2 sizes with point feet, or talos-like sizes with quad feet); same argument meaning as
// emul_kino_knot. The terminal node is the kinodynamics one (reference src/fulldynamics.cpp:418-445 builds the same
// terminal cost / constraint).
int emul_full_knot(
  const smpc_model_desc * model, const smpc_ocp_desc * ocp, const smpc_slot * slot, const smpc_terminal * term, int terminal,
  int derivs, const double * x, const double * u, const double * xn, const double * vs, const double * prev_vs,
  const double * lam_t, const double * lam_n, const double * prev_lam_n, double mu, double preg, double * knot, int * nc,
  double * gx, double * gu, double * lplus, double * xdot, double * scalars /*8*/)
{
  try
  {
    auto M = std::make_unique<DevModel>();
    auto O = std::make_unique<DevOcp>();
    build_dev_model(*model, *M);
    build_dev_ocp(*model, *ocp, *O);
    const bool go2 = model->njoints == 13 && model->nv == 18 && model->nfeet == 4 && O->fs == 3;
    const bool talos = model->njoints == 23 && model->nv == 28 && model->nfeet == 2 && O->fs == 6;
    if (O->kind != SMPC_OCP_FULLDYNAMICS || !(go2 || talos))
      throw std::runtime_error("emul: only the Go2-sized (point feet) and talos-sized (quad feet) full-dynamics instantiations are built");
    KnotLayout L = make_knot_layout(O->ndx, O->nu, O->ncmax);
    KnotIO io;
    memset(&io, 0, sizeof(io));
    io.x = x, io.u = u, io.xn = xn;
    io.slot = slot, io.term = term;
    io.vs = vs, io.prev_vs = prev_vs, io.lam_t = lam_t, io.lam_n = lam_n, io.prev_lam_n = prev_lam_n;
    io.mu = mu, io.preg = preg;
    io.knot = knot, io.nc_out = nc, io.gx = gx, io.gu = gu, io.lplus_n = lplus, io.xdot = xdot;
    KnotScalars sc;
    memset(&sc, 0, sizeof(sc));
    io.sc = &sc;
    if (go2)
      full_knot_any<FullCfg<13, 18, 4>, Go2Cfg>(*M, *O, L, io, terminal != 0, derivs != 0);
    else
      full_knot_any<FullCfg<23, 28, 2, 6>, KinoCfg<23, 28, 2, 6>>(*M, *O, L, io, terminal != 0, derivs != 0);
    memcpy(scalars, &sc, sizeof(sc));
  }
  catch (const std::exception & e)
  {
    g_err = e.what();
    return -1;
  }
  return 0;
}
int emul_full_scratch_bytes() { return (int)sizeof(FullScratch<FullCfg<13, 18, 4>>); }

int emul_centro_knot(
  const smpc_model_desc * model, const smpc_ocp_desc * ocp, const smpc_slot * slot, const smpc_terminal * term, int terminal,
  int derivs, const double * x, const double * u, const double * xn, const double * vs, const double * prev_vs,
  const double * lam_t, const double * lam_n, const double * prev_lam_n, double mu, double preg, double * knot, int * nc,
  double * gx, double * gu, double * lplus, double * xdot, double * scalars /*8*/)
{
  try
  {
    auto O = std::make_unique<DevOcp>();
    build_dev_ocp(*model, *ocp, *O);
    KnotLayout L = make_knot_layout(O->ndx, O->nu, O->ncmax);
    KnotIO io;
    memset(&io, 0, sizeof(io));
    io.x = x, io.u = u, io.xn = xn;
    io.slot = slot, io.term = term;
    io.vs = vs, io.prev_vs = prev_vs, io.lam_t = lam_t, io.lam_n = lam_n, io.prev_lam_n = prev_lam_n;
    io.mu = mu, io.preg = preg;
    io.knot = knot, io.nc_out = nc, io.gx = gx, io.gu = gu, io.lplus_n = lplus, io.xdot = xdot;
    KnotScalars sc;
    memset(&sc, 0, sizeof(sc));
    io.sc = &sc;
    if (O->kind != SMPC_OCP_CENTROIDAL)
      throw std::runtime_error("emul_centro_knot: not a centroidal problem");
    if (O->nk == 2 && O->fs == 6)
      centro_knot_any<CentroCfg<2, 6>>(*O, L, io, terminal, derivs);
    else if (O->nk == 4 && O->fs == 3)
      centro_knot_any<CentroCfg<4, 3>>(*O, L, io, terminal, derivs);
    else
      throw std::runtime_error("emul_centro_knot: shape not instantiated");
    memcpy(scalars, &sc, sizeof(sc));
  }
  catch (const std::exception & e)
  {
    g_err = e.what();
    return -1;
  }
  return 0;
}

int emul_knot_size(int ndx, int nu, int ncmax) { return make_knot_layout(ndx, nu, ncmax).size; }
int emul_sol_size(int ndx, int nu, int ncmax) { return make_sol_layout(ndx, nu, ncmax).size; }
void emul_knot_layout(int ndx, int nu, int ncmax, int * out /*12*/)
{
  KnotLayout L = make_knot_layout(ndx, nu, ncmax);
  int v[12] = {L.oQ, L.oS, L.oR, L.oq, L.or_, L.oA, L.oB, L.of, L.oC, L.oD, L.od, L.size};
  memcpy(out, v, sizeof(v));
}
void emul_sol_layout(int ndx, int nu, int ncmax, int * out /*7*/)
{
  SolLayout L = make_sol_layout(ndx, nu, ncmax);
  int v[7] = {L.oK, L.ok, L.oZ, L.oz, L.oPt, L.opt, L.size};
  memcpy(out, v, sizeof(v));
}

// same, through the TMA-staged sweep of riccati3.cuh, dense products / kinodynamics-structure products (the one the library launches)
#define EMUL_RIC_ARGS                                                                                                  \
  int N, int ndx, int nu, int ncmax, const double * knots, const int * ncs, double mudyn, double mueq, double * sol,   \
    const double * gx, const double * gu, const double * lplus, double * dxs, double * dus, double * vnew, double * lnew, double * dphi, int * flag
int emul_riccati3(EMUL_RIC_ARGS) { return emul_riccati3_impl<0>(N, ndx, nu, ncmax, knots, ncs, mudyn, mueq, sol, gx, gu, lplus, dxs, dus, vnew, lnew, dphi, flag); }
int emul_riccati3_kino_d0(EMUL_RIC_ARGS) { return emul_riccati3_impl<18, true>(N, ndx, nu, ncmax, knots, ncs, mudyn, mueq, sol, gx, gu, lplus, dxs, dus, vnew, lnew, dphi, flag); }
int emul_riccati3_kino(EMUL_RIC_ARGS) { return emul_riccati3_impl<18>(N, ndx, nu, ncmax, knots, ncs, mudyn, mueq, sol, gx, gu, lplus, dxs, dus, vnew, lnew, dphi, flag); }

int emul_riccati_parallel(
  int N, int ndx, int nu, int ncmax, const double * knots, const int * ncs, double mudyn, double mueq, double * sol,
  const double * gx, const double * gu, const double * lplus, double * dxs, double * dus, double * vnew, double * lnew,
  double * dphi, int * flag, int J)
{
#define PAR_ARGS N, J, knots, ncs, mudyn, mueq, sol, gx, gu, lplus, dxs, dus, vnew, lnew, dphi, flag
  if (ndx == 36 && nu == 24 && ncmax == 24)
    return parallel_any<36, 24, 24, true>(PAR_ARGS);
  if (ndx == 36 && nu == 24 && ncmax == 36)
    return parallel_any<36, 24, 36, true>(PAR_ARGS);
  if (ndx == 56 && nu == 34 && ncmax == 34)
    return parallel_any<56, 34, 34, false>(PAR_ARGS);
  if (ndx == 56 && nu == 34 && ncmax == 46)
    return parallel_any<56, 34, 46, false>(PAR_ARGS);
#undef PAR_ARGS
  g_err = "emul_riccati_parallel: size not instantiated";
  return -1;
}

// same, through the generic-shape two-column sweep of riccati4.cuh
int emul_riccati4(
  int N, int ndx, int nu, int ncmax, const double * knots, const int * ncs, double mudyn, double mueq, double * sol,
  const double * gx, const double * gu, const double * lplus, double * dxs, double * dus, double * vnew, double * lnew,
  double * dphi, int * flag)
{
  KnotLayout L = make_knot_layout(ndx, nu, ncmax);
  SolLayout SL = make_sol_layout(ndx, nu, ncmax);
  std::vector<double> P0((size_t)ndx * (ndx + 1), 0.0), sm2(4 * ndx + nu + 80, 0.0);
  if (ndx == 36 && nu == 24 && ncmax == 24)
    ric4_any<36, 24, 24>(L, SL, N, knots, sol, mudyn, mueq, flag, P0.data());
  else if (ndx == 36 && nu == 24 && ncmax == 36)
    ric4_any<36, 24, 36>(L, SL, N, knots, sol, mudyn, mueq, flag, P0.data());
  else if (ndx == 36 && nu == 12 && ncmax == 24)
    ric4_any<36, 12, 24>(L, SL, N, knots, sol, mudyn, mueq, flag, P0.data());
  else if (ndx == 36 && nu == 12 && ncmax == 60)
    ric4_any<36, 12, 60>(L, SL, N, knots, sol, mudyn, mueq, flag, P0.data());
  else if (ndx == 56 && nu == 22 && ncmax == 78)
    ric4_any<56, 22, 78>(L, SL, N, knots, sol, mudyn, mueq, flag, P0.data());
  else if (ndx == 56 && nu == 22 && ncmax == 90)
    ric4_any<56, 22, 90>(L, SL, N, knots, sol, mudyn, mueq, flag, P0.data());
  else if (ndx == 9 && nu == 12 && ncmax == 34)
    ric4_any<12, 12, 34, 9>(L, SL, N, knots, sol, mudyn, mueq, flag, P0.data());
  else if (ndx == 9 && nu == 12 && ncmax == 8)
    ric4_any<12, 12, 8, 9>(L, SL, N, knots, sol, mudyn, mueq, flag, P0.data());
  else if (ndx == 56 && nu == 34 && ncmax == 34)
    ric4_any<56, 34, 34>(L, SL, N, knots, sol, mudyn, mueq, flag, P0.data());
  else if (ndx == 56 && nu == 34 && ncmax == 46)
    ric4_any<56, 34, 46>(L, SL, N, knots, sol, mudyn, mueq, flag, P0.data());
  else
  {
    g_err = "emul_riccati4: size not instantiated";
    return -1;
  }
  riccati_forward_inst(L, SL, N, knots, ncs, sol, P0.data(), mudyn, gx, gu, lplus, dxs, dus, vnew, lnew, dphi, sm2.data());
  return 0;
}

// backward + forward Riccati sweep of one instance on dense knots in the device layout
int emul_riccati(
  int N, int ndx, int nu, int ncmax, const double * knots, const int * ncs, double mudyn, double mueq, double * sol,
  const double * gx, const double * gu, const double * lplus, double * dxs, double * dus, double * vnew, double * lnew,
  double * dphi, int * flag)
{
  KnotLayout L = make_knot_layout(ndx, nu, ncmax);
  SolLayout SL = make_sol_layout(ndx, nu, ncmax);
  RicSmem RS = make_ric_smem(ndx, nu, ncmax);
  std::vector<double> sm(RS.size + 8, 0.0), P0((size_t)ndx * (ndx + 1), 0.0), sm2(4 * ndx + nu + 80, 0.0);
  riccati_backward_inst(L, SL, RS, N, knots, ncs, sol, mudyn, mueq, sm.data(), flag, P0.data());
  riccati_forward_inst(L, SL, N, knots, ncs, sol, P0.data(), mudyn, gx, gu, lplus, dxs, dus, vnew, lnew, dphi, sm2.data());
  return 0;
}

// ---- batched MPC driver (mpc_driver.cuh) for ONE instance: the slot ring lives in host memory ----
struct EmulMpc
{
  DevModel M;
  DevOcp O;
  MpcShared S;
  MpcInst I;
  MpcScratch scratch;
  std::vector<smpc_slot> table, ring;
  std::vector<double> x0;
  smpc_terminal term;
  int head, ncycle;
};

void * emul_mpc_create(
  const smpc_model_desc * model, const smpc_ocp_desc * ocp, const smpc_slot * table, int ntable, const smpc_mpc_settings * st,
  int phase, const double * x_reference, const smpc_slot * initial_slots /*[T]*/, const smpc_terminal * term)
{
  auto * e = new EmulMpc();
  build_dev_model(*model, e->M);
  build_dev_ocp(*model, *ocp, e->O);
  e->table.assign(table, table + ntable);
  e->ring.assign(initial_slots, initial_slots + e->O.N);
  e->term = *term;
  e->head = 0;
  e->ncycle = st->ncycle;
  e->x0.assign((size_t)e->O.nx, 0.0);
  memset(&e->S, 0, sizeof(e->S));
  e->S.T_fly = st->T_fly;
  e->S.T_contact = st->T_contact;
  e->S.swing_apex = st->swing_apex;
  e->S.timestep = st->timestep;
  e->S.standing_id = st->standing_id;
  memset(&e->I, 0, sizeof(e->I));
  std::vector<unsigned> masks((size_t)st->ncycle);
  for (int k = 0; k < st->ncycle; k++)
    masks[(size_t)k] = table[k].contact_mask;
  mpc_init_queues(masks.data(), st->ncycle, e->O.nk, e->O.N, phase, e->I);
  e->I.walking = 1;
  for (int k = 0; k < e->O.nx; k++)
    e->I.x_reference[k] = x_reference[k];
  mpc_init_instance(e->M, model->q_ref, &e->S, e->I, e->scratch);
  return e;
}
void emul_mpc_destroy(void * h) { delete static_cast<EmulMpc *>(h); }
void emul_mpc_set_velocity_base(void * h, const double * vb)
{
  for (int k = 0; k < 6; k++)
    static_cast<EmulMpc *>(h)->I.velocity_base[k] = vb[k];
}
void emul_mpc_set_walking(void * h, int walking)
{
  auto * e = static_cast<EmulMpc *>(h);
  e->I.walking = walking;
  if (!walking)
    for (int k = 0; k < 6; k++)
      e->I.velocity_base[k] = 0.0;
}
// recede + cycle + references of one tick; outputs in horizon order
int emul_mpc_tick(void * h, const double * xmeas, smpc_slot * slots_out, smpc_terminal * term_out, double * x0_out, int * takeoff, int * land)
{
  auto * e = static_cast<EmulMpc *>(h);
  const int N = e->O.N, nk = e->O.nk;
  const unsigned tail_mask = e->ring[(size_t)((e->head + N - 1) % N)].contact_mask;
  const int id = mpc_recede(e->S, e->I, e->table.data(), e->ncycle, nk, N, tail_mask);
  e->ring[(size_t)e->head] = e->table[(size_t)id];
  e->head = (e->head + 1) % N;
  MpcRing ring;
  ring.base = e->ring.data();
  ring.head = e->head;
  ring.N = N;
  mpc_refs(e->M, e->O, e->S, e->I, xmeas, ring, e->term, e->x0.data(), e->scratch);
  for (int t = 0; t < N; t++)
    slots_out[t] = ring.at(t);
  *term_out = e->term;
  for (int k = 0; k < e->O.nx; k++)
    x0_out[k] = e->x0[(size_t)k];
  for (int f = 0; f < nk; f++)
  {
    takeoff[f] = e->I.ntake[f] > 0 ? e->I.take[f][0] : -1;
    land[f] = e->I.nland[f] > 0 ? e->I.land[f][0] : -1;
  }
  return e->I.overflow;
}
}
