// TEST HARNESS — not part of the product. Python module `_smpc_hooks`: builds a simple_mpc.MPC whose solver
// table (simple-mpc/solver-api.hpp) is bound to the CPU ORACLE (oracle/liborc.so, `orc_` entry points with
// the signatures of include/smpc.h) instead of libsmpc_cuda, so that
//   * the host driver logic (horizon cycling, foot timings, swing references) can be tested on the GPU-less box;
//   * on a GPU box the same Python script drives both and compares them end to end.
// The shipped module `simple_mpc` has no such entry point: it always binds SolverApi::cuda().
#include "../../simple-mpc_b200/bindings/py-convert.hpp"
#include <dlfcn.h>

namespace py = pybind11;
using namespace simple_mpc;

namespace
{
  typedef int (*orc_create_t)(smpc_ctx **, int, const smpc_model_desc *, const smpc_ocp_desc *, const smpc_solver_opts *);
  orc_create_t g_orc_create = nullptr;
  int create_adapter(smpc_ctx ** out, int /*device*/, int batch, const smpc_model_desc * m, const smpc_ocp_desc * o, const smpc_solver_opts * s)
  {
    return g_orc_create(out, batch, m, o, s);
  }
  SolverApi g_api;
  bool g_loaded = false;

  template <class F>
  void sym(void * h, const char * name, F & f)
  {
    void * p = dlsym(h, name);
    if (!p)
      throw std::runtime_error(std::string("missing symbol ") + name);
    f = reinterpret_cast<F>(p);
  }

  const SolverApi & oracleApi(const std::string & path)
  {
    if (g_loaded)
      return g_api;
    void * h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h)
      throw std::runtime_error(std::string("cannot load ") + path + ": " + dlerror());
    sym(h, "orc_last_error", g_api.last_error);
    sym(h, "orc_default_solver_opts", g_api.default_solver_opts);
    sym(h, "orc_problem_dims", g_api.problem_dims);
    sym(h, "orc_create", g_orc_create);
    g_api.create = &create_adapter;
    sym(h, "orc_destroy", g_api.destroy);
    sym(h, "orc_set_slots", g_api.set_slots);
    sym(h, "orc_set_terminal", g_api.set_terminal);
    sym(h, "orc_cycle", g_api.cycle);
    sym(h, "orc_set_warmstart", g_api.set_warmstart);
    sym(h, "orc_solve", g_api.solve);
    sym(h, "orc_get_results", g_api.get_results);
    sym(h, "orc_get_info", g_api.get_info);
    sym(h, "orc_get_xdot", g_api.get_xdot);
    sym(h, "orc_get_contact_forces", g_api.get_contact_forces);
    g_api.name = "CPU oracle (test harness)";
    g_loaded = true;
    return g_api;
  }

  struct HookMPC : MPC
  {
    HookMPC(const MPCSettings & s, std::shared_ptr<OCPHandler> p, const SolverApi & api) : MPC(s, p, api) {}
  };
} // namespace

PYBIND11_MODULE(_smpc_hooks, m)
{
  py::module_::import("simple_mpc");
  m.def(
    "make_oracle_mpc",
    [](const py::dict & settings, std::shared_ptr<OCPHandler> ocp, const std::string & liborc_path) -> std::unique_ptr<MPC> {
      return std::unique_ptr<MPC>(new HookMPC(python::mpcSettingsFromDict(settings), ocp, oracleApi(liborc_path)));
    },
    "simple_mpc.MPC driven by the CPU oracle (tests only)");
}
