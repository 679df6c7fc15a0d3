"""The consumer of the MPC plan between two ticks: Riccati-feedback control law
u = us[0] - Ks[0] difference(x_meas, xs[0]) and the Interpolator applied to the plan
(reference examples/talos_fulldynamics.py:203-205, src/robot-handler.cpp:81-96, src/interpolator.cpp:5-78).

CPU part: the oracle's restatement against the defining properties (values at the knots, clamping beyond the horizon,
linearity in the velocity part, u(x_meas = xs[0]) = us[0], first-order agreement with the gain).
GPU part (-m gpu): the CUDA entry points against the oracle on the same plan."""
import ctypes as C
import numpy as np
import pytest

import parity_utils as pu
from smpc_ctypes import SolverOpts, load_cuda
from oracle_libs import load_oracle

DT = 0.01


def make(lib, prefix, sc, B):
    opts = SolverOpts()
    getattr(lib, prefix + "_default_solver_opts")(C.byref(opts))
    b = pu.Backend(lib, prefix, sc.model, sc.ocpb, opts, B)
    b.set_slots(sc.initial_slots())
    b.set_terminal(sc.terminals())
    b.set_warmstart(*sc.warmstart())
    b.solve(np.tile(sc.xref, (B, 1)), 100)
    return b


def advance(backends, sc, ticks):
    x = None
    for tick in range(ticks):
        tails = sc.tails(tick)
        xo = backends[0].results(gains=False)[0]
        x = sc.disturb(xo[:, 1, :].copy(), tick)
        for b in backends:
            b.cycle(tails)
            b.shift()
            b.solve(x, 1)
    return x


def feedback(lib, prefix, b, xm):
    xm = np.ascontiguousarray(xm)
    u = np.zeros((b.B, b.nu))
    rc = getattr(lib, prefix + "_feedback_control")(b.ctx, 0, b.B, pu.P(xm), pu.P(u))
    assert rc == 0, getattr(lib, prefix + "_last_error")()
    return u


def interp(lib, prefix, b, delay, nk=4):
    x, u = np.zeros((b.B, b.nx)), np.zeros((b.B, b.nu))
    c = np.zeros((b.B, nk), dtype=np.int32)
    rc = getattr(lib, prefix + "_interpolate_plan")(b.ctx, 0, b.B, C.c_double(delay), pu.P(x), pu.P(u), c.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0, getattr(lib, prefix + "_last_error")()
    return x, u, c


def test_oracle_control_law_and_interpolation_properties():
    orc = load_oracle()
    sc = pu.Go2Scenario(2, seed=11)
    b = make(orc, "orc", sc, 2)
    advance([b], sc, 12)  # the first swing phases have entered the horizon
    xs, us, Ks, _ = b.results()
    # control law: exact at the plan, first order in the deviation, sign of the reference formula
    assert np.array_equal(feedback(orc, "orc", b, xs[:, 0]), us[:, 0])
    xm = xs[:, 0].copy()
    xm[:, 19:] += 1e-3  # velocity deviation: difference(x_meas, xs0) = -1e-3 on the velocity rows
    u = feedback(orc, "orc", b, xm)
    for i in range(2):
        assert np.allclose(u[i], us[i, 0] + Ks[i, 0][:, 18:] @ (1e-3 * np.ones(18)), rtol=0, atol=1e-12 * max(1.0, np.abs(Ks[i, 0]).max()))
    # interpolation: knots, clamping, linearity
    for k in (0, 1, 17, 49):
        x, u, c = interp(orc, "orc", b, k * DT + 1e-12)
        assert np.allclose(x, xs[:, k], atol=1e-9) and np.allclose(u, us[:, min(k, 49)], atol=1e-7 * np.abs(us).max())
    x, u, c = interp(orc, "orc", b, 0.7)
    assert np.array_equal(x, xs[:, 50]) and np.array_equal(u, us[:, 49])
    x, u, c = interp(orc, "orc", b, 3.5 * DT)
    assert np.allclose(x[:, 19:], 0.5 * (xs[:, 3, 19:] + xs[:, 4, 19:]), atol=1e-13)
    assert np.allclose(x[:, 7:19], 0.5 * (xs[:, 3, 7:19] + xs[:, 4, 7:19]), atol=1e-13)  # revolute joints: linear
    assert np.allclose(np.linalg.norm(x[:, 3:7], axis=1), 1.0, atol=1e-14)
    assert np.allclose(u, 0.5 * (us[:, 3] + us[:, 4]), atol=1e-12 * np.abs(us).max())
    # contacts: the mask of the stage the delay falls into
    for k in (0, 30, 49, 80):
        _, _, c = interp(orc, "orc", b, k * DT + 1e-12)
        for i in range(2):
            mask = sc.GAIT[(int(sc.phase[i]) + 12 - 50 + min(k, 49)) % 80] if 12 - 50 + min(k, 49) >= 0 else 0b1111
            assert [int(v) for v in c[i]] == [(mask >> f) & 1 for f in range(4)], (k, i)
    b.close()


@pytest.mark.gpu
def test_cuda_control_law_and_interpolation_match_the_oracle():
    orc, cu = load_oracle(), load_cuda()
    sc = pu.Go2Scenario(3, seed=12)
    bo, bc = make(orc, "orc", sc, 3), make(cu, "smpc", sc, 3)
    x = advance([bo, bc], sc, 15)
    rng = np.random.default_rng(0)
    xm = x.copy()
    xm[:, 19:] += rng.normal(size=(3, 18)) * 0.05
    xm[:, 7:19] += rng.normal(size=(3, 12)) * 0.02
    xm[:, 0:3] += rng.normal(size=(3, 3)) * 0.01
    uo, uc = feedback(orc, "orc", bo, xm), feedback(cu, "smpc", bc, xm)
    assert pu.relerr(uc, uo) <= 1e-7, pu.relerr(uc, uo)  # limited by the parity of K (tests/test_gpu_parity.py: TOL_K)
    for delay in (0.0, 0.0037, 0.2512, 0.4899, 0.4951, 0.5, 0.7):
        (xo, u_o, co), (xc, u_c, cc) = interp(orc, "orc", bo, delay), interp(cu, "smpc", bc, delay)
        assert pu.relerr(xc, xo) <= 1e-8 and pu.relerr(u_c, u_o) <= 1e-8, (delay, pu.relerr(xc, xo), pu.relerr(u_c, u_o))
        assert np.array_equal(co, cc)
    bo.close()
    bc.close()
