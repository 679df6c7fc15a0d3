"""Parity harness shared by the tests and __graft_entry__.smoke(): drives the CPU oracle and the CUDA library through
the same C-ABI call sequence (smpc_scenarios.Backend) and compares the results."""
import ctypes as C
import numpy as np
from smpc_ctypes import SolverOpts, load_cuda
from smpc_scenarios import (P, Backend, Go2Scenario, Go2FullScenario, TalosCentroidalScenario, TalosFullScenario,  # noqa: F401
                            TalosKinoScenario)
from oracle_libs import load_oracle


def relerr(a, b):
    den = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / den)


def relerr_elem(a, b, floor_frac=1e-3):
    """ELEMENTWISE relative error with an absolute floor: max_i |a_i - b_i| / max(|b_i|, floor_frac * max|b|). Unlike the
    normwise figure above it does not hide the small entries of u0 / K0 - the numbers a robot consumes - behind the
    largest one; entries below floor_frac of the array's scale are compared against that floor."""
    b = np.asarray(b)
    den = np.maximum(np.abs(b), floor_frac * max(np.abs(b).max(), 1e-300))
    return float((np.abs(np.asarray(a) - b) / den).max())


def run_parity(batch=2, ticks=3, seed=0, ctor_iters=100, verbose=False, scenario=None, cuda_threads=1, **ocp_kw):
    """Drive the oracle and the CUDA library through the same constructor solve + `ticks` MPC ticks and return
    the worst normwise relative differences."""
    orc, cu = load_oracle(), load_cuda()
    sc = (scenario or Go2Scenario)(batch, seed, **ocp_kw)
    opts_o, opts_c = SolverOpts(), SolverOpts()
    orc.orc_default_solver_opts(C.byref(opts_o))
    cu.smpc_default_solver_opts(C.byref(opts_c))
    assert bytes(opts_o) == bytes(opts_c), "default solver options differ"
    opts_c.num_threads = cuda_threads  # > 1: the CUDA side runs the PARALLEL (leg-condensed) sweep, the oracle stays serial
    bo = Backend(orc, "orc", sc.model, sc.ocpb, opts_o, batch)
    bc = Backend(cu, "smpc", sc.model, sc.ocpb, opts_c, batch)
    worst = dict(xs=0.0, us=0.0, K=0.0, k=0.0, iters_equal=True, alphas_equal=True, prim=0.0, dual=0.0, u0_elem=0.0, K0_elem=0.0)
    try:
        for b in (bo, bc):
            b.set_slots(sc.initial_slots())
            b.set_terminal(sc.terminals())
            xs, us = sc.warmstart()
            b.set_warmstart(xs, us)
        # constructor solve at the reference state (reference src/mpc.cpp:24-26,85); disturbances enter at the ticks
        x = np.tile(sc.xref, (batch, 1))

        def compare(tag):
            xo, uo, Ko, ko = bo.results()
            xc, uc, Kc, kc = bc.results()
            io, ic = bo.info(), bc.info()
            for i in range(batch):
                worst["xs"] = max(worst["xs"], relerr(xc[i], xo[i]))
                worst["us"] = max(worst["us"], relerr(uc[i], uo[i]))
                worst["K"] = max(worst["K"], max(relerr(Kc[i, t], Ko[i, t]) for t in range(sc.T)))
                worst["k"] = max(worst["k"], relerr(kc[i], ko[i]))
                # what the controller consumes each tick: u0 and K0, element by element
                worst["u0_elem"] = max(worst["u0_elem"], relerr_elem(uc[i, 0], uo[i, 0]))
                worst["K0_elem"] = max(worst["K0_elem"], relerr_elem(Kc[i, 0], Ko[i, 0]))
                if io[i].num_iters != ic[i].num_iters or io[i].conv != ic[i].conv:
                    worst["iters_equal"] = False
                if io[i].alpha != ic[i].alpha or io[i].ls_trials != ic[i].ls_trials:
                    worst["alphas_equal"] = False
                worst["prim"] = max(worst["prim"], abs(io[i].prim_infeas - ic[i].prim_infeas) / max(abs(io[i].prim_infeas), 1e-300))
                worst["dual"] = max(worst["dual"], abs(io[i].dual_infeas - ic[i].dual_infeas) / max(abs(io[i].dual_infeas), 1e-300))
            if verbose:
                print(tag, "iters", [(io[i].num_iters, ic[i].num_iters) for i in range(batch)], "alpha", [(io[i].alpha, ic[i].alpha) for i in range(batch)],
                      {k: ("%.2e" % v if isinstance(v, float) else v) for k, v in worst.items()})
            return xo

        bo.solve(x, ctor_iters)
        bc.solve(x, ctor_iters)
        xo = compare("ctor")
        for tick in range(ticks):
            tails = sc.tails(tick)
            # measured state: the oracle's predicted next state plus a small disturbance (same for both backends)
            x = sc.disturb(xo[:, 1, :].copy(), tick)
            for b in (bo, bc):
                b.cycle(tails)
                b.shift()
                b.solve(x, 1)
            xo = compare("tick %d" % tick)
    finally:
        bo.close()
        bc.close()
    return worst
