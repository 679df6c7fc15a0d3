"""PARALLEL (leg-condensed) Riccati sweep, reference src/mpc.cpp:46-50 (num_threads > 1 -> LQSolverChoice::PARALLEL).
CPU part: the oracle's leg-parallel sweep (oracle/orc_solver.hpp: riccati_backward_parallel) against its serial sweep on
whole MPC runs - the two must agree to the amplified-rounding floor of the serial recursion itself."""
import ctypes as C
import numpy as np
import pytest
import parity_utils as pu
from smpc_ctypes import SolverOpts
from oracle_libs import load_oracle


def _run(orc, sc, legs, ticks, batch):
    opts = SolverOpts()
    orc.orc_default_solver_opts(C.byref(opts))
    opts.num_threads = legs
    b = pu.Backend(orc, "orc", sc.model, sc.ocpb, opts, batch)
    out = []
    try:
        b.set_slots(sc.initial_slots())
        b.set_terminal(sc.terminals())
        b.set_warmstart(*sc.warmstart())
        x = np.tile(sc.xref, (batch, 1))
        b.solve(x, 100)
        out.append(tuple(np.array(a) for a in b.results()) + ([(i.num_iters, i.ls_trials, i.alpha) for i in b.info()],))
        for tick in range(ticks):
            x = sc.disturb(out[-1][0][:, 1, :].copy(), tick)
            b.cycle(sc.tails(tick))
            b.shift()
            b.solve(x, 1)
            out.append(tuple(np.array(a) for a in b.results()) + ([(i.num_iters, i.ls_trials, i.alpha) for i in b.info()],))
    finally:
        b.close()
    return out


@pytest.mark.parametrize("legs", [2, 3, 5, 10])
def test_oracle_parallel_sweep_equals_the_serial_sweep(legs):
    orc = load_oracle()
    sc = pu.Go2Scenario(1, 3)
    ser = _run(orc, pu.Go2Scenario(1, 3), 1, 3, 1)  # a scenario draws its disturbances from its own generator: one per run
    par = _run(orc, pu.Go2Scenario(1, 3), legs, 3, 1)
    worst = dict(xs=0.0, us=0.0, K=0.0)
    for (xs, us, Ks, ks, inf), (xp, up, Kp, kp, infp) in zip(ser, par):
        assert inf == infp
        worst["xs"] = max(worst["xs"], pu.relerr(xp, xs))
        worst["us"] = max(worst["us"], pu.relerr(up, us))
        worst["K"] = max(worst["K"], max(pu.relerr(Kp[0, t], Ks[0, t]) for t in range(sc.T)))
    if legs == 2:
        # two legs are the serial sweep cut in two: the hand-over of the cost-to-go is exact
        assert worst == dict(xs=0.0, us=0.0, K=0.0), worst
    assert worst["xs"] <= 1e-8 and worst["us"] <= 1e-8 and worst["K"] <= 1e-7, worst


@pytest.mark.gpu
@pytest.mark.parametrize("legs,cone", [(5, False), (10, False), (5, True)])
def test_cuda_parallel_sweep_against_the_serial_oracle(legs, cone, monkeypatch):
    """The leg-parallel sweep on the Go2 kinodynamics problem (riccati3 legs; forced with SMPC_LEGS because the serial sweep
    is the faster one for this shape and therefore the default): four launches per backward sweep - tuned sweep per leg from
    a zero cost-to-go, leg elements, condensation, tuned sweep per leg from the true cost-to-go - against the SERIAL oracle:
    constructor solve + 4 MPC ticks, same iteration counts / step sizes, xs, us, K within the tolerances of the serial test."""
    from smpc_ctypes import load_cuda
    cu = load_cuda()
    monkeypatch.setenv("SMPC_LEGS", str(legs))
    kw = dict(force_cone=True, land_cstr=True, terminal_constraint=True) if cone else {}
    # the context must really run the leg-parallel sweep
    sc = pu.Go2Scenario(2, 0, **kw)
    opts = SolverOpts()
    cu.smpc_default_solver_opts(C.byref(opts))
    b = pu.Backend(cu, "smpc", sc.model, sc.ocpb, opts, 2)
    assert cu.smpc_riccati_legs(b.ctx) == legs
    b.close()
    res = pu.run_parity(batch=2, ticks=4, seed=0, **kw)
    assert res["iters_equal"] and res["alphas_equal"], res
    assert res["xs"] <= 1e-7 and res["us"] <= 1e-7 and res["K"] <= (1e-5 if cone else 1e-6), res


@pytest.mark.gpu
def test_num_threads_selects_the_parallel_sweep_for_the_centroidal_problem():
    """MPCSettings::num_threads > 1 -> PARALLEL (reference src/mpc.cpp:46-50): for the centroidal problem (config C2, T = 100)
    the leg-parallel sweep is the measured winner and num_threads alone selects it; results against the SERIAL oracle."""
    from smpc_ctypes import load_cuda
    cu = load_cuda()
    sc = pu.TalosCentroidalScenario(2, 0)
    opts = SolverOpts()
    cu.smpc_default_solver_opts(C.byref(opts))
    opts.num_threads = 8
    b = pu.Backend(cu, "smpc", sc.model, sc.ocpb, opts, 2)
    assert cu.smpc_riccati_legs(b.ctx) == 8
    b.close()
    res = pu.run_parity(batch=2, ticks=6, seed=0, cuda_threads=8, scenario=pu.TalosCentroidalScenario)
    assert res["iters_equal"] and res["alphas_equal"], res
    assert res["xs"] <= 1e-8 and res["us"] <= 1e-8 and res["K"] <= 1e-7, res
    # Go2 kinodynamics: the serial sweep is faster, num_threads does not change what runs
    sc = pu.Go2Scenario(1, 0)
    b = pu.Backend(cu, "smpc", sc.model, sc.ocpb, opts, 1)
    assert cu.smpc_riccati_legs(b.ctx) == 1
    b.close()


@pytest.mark.gpu
def test_cuda_parallel_sweep_is_not_used_for_throughput_bound_batches(monkeypatch):
    """batch x (legs - 1) CTAs must fit the SMs once: a 64-instance context asked for 5 legs runs the serial sweep."""
    from smpc_ctypes import load_cuda
    cu = load_cuda()
    monkeypatch.setenv("SMPC_LEGS", "5")
    sc = pu.Go2Scenario(64, 0)
    opts = SolverOpts()
    cu.smpc_default_solver_opts(C.byref(opts))
    b = pu.Backend(cu, "smpc", sc.model, sc.ocpb, opts, 64)
    assert cu.smpc_riccati_legs(b.ctx) == 1
    b.close()


@pytest.mark.gpu
def test_serial_fallback_of_the_parallel_sweep_is_the_serial_sweep(monkeypatch):
    """When the condensed system of the leg-parallel sweep cannot be factorised (a cost-to-go that is not positive definite)
    pass 4 runs the serial sweep for that instance. SMPC_LEGS_FORCE_FALLBACK takes that branch for every instance: results
    are bitwise those of a context that runs the serial sweep, and smpc_solve_info.flags carries bit 2."""
    from smpc_ctypes import load_cuda
    cu = load_cuda()

    def run(legs, force):
        if legs:
            monkeypatch.setenv("SMPC_LEGS", str(legs))
        else:
            monkeypatch.delenv("SMPC_LEGS", raising=False)
        if force:
            monkeypatch.setenv("SMPC_LEGS_FORCE_FALLBACK", "1")
        else:
            monkeypatch.delenv("SMPC_LEGS_FORCE_FALLBACK", raising=False)
        sc = pu.Go2Scenario(2, 4)
        opts = SolverOpts()
        cu.smpc_default_solver_opts(C.byref(opts))
        b = pu.Backend(cu, "smpc", sc.model, sc.ocpb, opts, 2)
        assert cu.smpc_riccati_legs(b.ctx) == (legs or 1)
        b.set_slots(sc.initial_slots())
        b.set_terminal(sc.terminals())
        b.set_warmstart(*sc.warmstart())
        b.solve(np.tile(sc.xref, (2, 1)), 100)
        out = [np.array(a) for a in b.results()] + [[i.flags for i in b.info()]]
        b.close()
        return out

    ser, fb = run(0, False), run(5, True)
    assert all(np.array_equal(a, c) for a, c in zip(ser[:4], fb[:4]))
    assert all(f & 4 for f in fb[4]) and not any(f & 4 for f in ser[4])
