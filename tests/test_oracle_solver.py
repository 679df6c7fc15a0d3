"""Known-answer and property checks of the oracle's ProxDDP / Riccati restatement.

  * the proximal Riccati sweep solves its KKT system: stationarity, dynamics and constraint rows of the
    dual-regularised LQ problem vanish at the returned (dx, du, nu, lambda) — checked on random problems;
  * an unconstrained LQ problem matches a dense numpy solve of the same KKT system;
  * the MPC scenario: the constructor solve converges from the reference state, ticks accept full steps,
    contact forces carry the robot's weight, multipliers of inactive inequality rows are zero;
  * API edge cases: index errors mirror the reference's std::runtime_error paths.
"""
import ctypes as C
import numpy as np
import pytest
import parity_utils as pu
from smpc_ctypes import SolverOpts, Slot, dbl_p
from oracle_libs import load_oracle


def P(a):
    return a.ctypes.data_as(dbl_p)


def random_lq(rng, N, ndx, nu, ncs, nct):
    stages, dense = [], []
    for t in range(N):
        H = rng.normal(size=(ndx + nu, ndx + nu))
        H = H @ H.T + np.eye(ndx + nu)
        Q, S, R = H[:ndx, :ndx], H[:ndx, ndx:], H[ndx:, ndx:]
        q, r = rng.normal(size=ndx), rng.normal(size=nu)
        A = np.eye(ndx) + 0.1 * rng.normal(size=(ndx, ndx))
        B = rng.normal(size=(ndx, nu))
        f = 0.1 * rng.normal(size=ndx)
        Cm, D, d = rng.normal(size=(ncs, ndx)), rng.normal(size=(ncs, nu)), rng.normal(size=ncs)
        dense.append(dict(Q=Q, S=S, R=R, q=q, r=r, A=A, B=B, f=f, C=Cm, D=D, d=d))
        stages.append(np.concatenate([Q.ravel(), S.ravel(), R.ravel(), q, r, A.ravel(), B.ravel(), f, Cm.ravel(), D.ravel(), d]))
    H = rng.normal(size=(ndx, ndx))
    QN = H @ H.T + np.eye(ndx)
    qN = rng.normal(size=ndx)
    CN, dN = rng.normal(size=(nct, ndx)), rng.normal(size=nct)
    term = np.concatenate([QN.ravel(), qN, CN.ravel(), dN])
    return np.concatenate(stages), term, dense, dict(Q=QN, q=qN, C=CN, d=dN)


@pytest.mark.parametrize("ncs,nct,mudyn", [(0, 0, 0.0), (3, 2, 1e-6), (4, 0, 1e-3)])
def test_riccati_solves_its_kkt_system(ncs, nct, mudyn):
    orc = load_oracle()
    rng = np.random.default_rng(0)
    N, ndx, nu, mueq = 7, 5, 3, 1e-4
    stages, term, dense, tn = random_lq(rng, N, ndx, nu, ncs, nct)
    K, k = np.zeros((N, nu, ndx)), np.zeros((N, nu))
    P0, p0 = np.zeros((ndx, ndx)), np.zeros(ndx)
    dxs, dus = np.zeros((N + 1, ndx)), np.zeros((N, nu))
    rc = orc.orc_riccati(N, ndx, nu, ncs, nct, C.c_double(mudyn), C.c_double(mueq), P(stages), P(term), P(K), P(k), P(P0), P(p0),
                         P(dxs), P(dus))
    assert rc == 0, orc.orc_last_error()
    # dense KKT of the same dual-regularised problem with dx0 = 0
    nz = N * (ndx + nu) + ndx  # (x1..xN interleaved later) use explicit variable ordering: u0, x1, u1, ..., x_N
    nvar = N * nu + N * ndx
    ncon = N * ndx + N * ncs + nct
    ix = lambda t: N * nu + (t - 1) * ndx  # x_t, t >= 1
    iu = lambda t: t * nu
    Hm, g = np.zeros((nvar, nvar)), np.zeros(nvar)
    J, c = np.zeros((ncon, nvar)), np.zeros(ncon)
    reg = np.zeros(ncon)
    row = 0
    for t in range(N):
        d = dense[t]
        Hm[iu(t):iu(t) + nu, iu(t):iu(t) + nu] += d['R']
        g[iu(t):iu(t) + nu] += d['r']
        if t >= 1:
            Hm[ix(t):ix(t) + ndx, ix(t):ix(t) + ndx] += d['Q']
            Hm[ix(t):ix(t) + ndx, iu(t):iu(t) + nu] += d['S']
            Hm[iu(t):iu(t) + nu, ix(t):ix(t) + ndx] += d['S'].T
            g[ix(t):ix(t) + ndx] += d['q']
        # dynamics A x + B u - x+ + f - mudyn lam = 0
        if t >= 1:
            J[row:row + ndx, ix(t):ix(t) + ndx] = d['A']
        J[row:row + ndx, iu(t):iu(t) + nu] = d['B']
        J[row:row + ndx, ix(t + 1):ix(t + 1) + ndx] = -np.eye(ndx)
        c[row:row + ndx] = d['f']
        reg[row:row + ndx] = mudyn
        row += ndx
        if t >= 1:
            J[row:row + ncs, ix(t):ix(t) + ndx] = d['C']
        J[row:row + ncs, iu(t):iu(t) + nu] = d['D']
        c[row:row + ncs] = d['d']
        reg[row:row + ncs] = mueq
        row += ncs
    Hm[ix(N):ix(N) + ndx, ix(N):ix(N) + ndx] += tn['Q']
    g[ix(N):ix(N) + ndx] += tn['q']
    J[row:row + nct, ix(N):ix(N) + ndx] = tn['C']
    c[row:row + nct] = tn['d']
    reg[row:row + nct] = mueq
    KKT = np.block([[Hm, J.T], [J, -np.diag(reg)]])
    sol = np.linalg.solve(KKT, -np.concatenate([g, c]))
    for t in range(N):
        assert np.abs(sol[iu(t):iu(t) + nu] - dus[t]).max() < 1e-8 * max(1.0, np.abs(sol).max())
        assert np.abs(sol[ix(t + 1):ix(t + 1) + ndx] - dxs[t + 1]).max() < 1e-8 * max(1.0, np.abs(sol).max())
    # feedback gains reproduce the open-loop step
    assert np.abs(k[0] - dus[0]).max() < 1e-12
    for t in range(1, N):
        assert np.abs(k[t] + K[t] @ dxs[t] - dus[t]).max() < 1e-10 * max(1.0, np.abs(dus).max())


@pytest.fixture(scope="module")
def mpc_run():
    orc = load_oracle()
    sc = pu.Go2Scenario(2, seed=0)
    o = SolverOpts()
    orc.orc_default_solver_opts(C.byref(o))
    b = pu.Backend(orc, "orc", sc.model, sc.ocpb, o, 2)
    b.set_slots(sc.initial_slots())
    b.set_terminal(sc.terminals())
    b.set_warmstart(*sc.warmstart())
    b.solve(np.tile(sc.xref, (2, 1)), 100)
    ctor = [(i.num_iters, i.conv, i.prim_infeas, i.dual_infeas) for i in b.info()]
    ticks = []
    xs, us, _, _ = b.results()
    for t in range(12):
        b.cycle(sc.tails(t))
        b.shift()
        b.solve(xs[:, 1, :].copy(), 1)
        xs, us, Ks, ks = b.results()
        ticks.append(dict(info=[(i.num_iters, i.alpha, i.ls_trials, i.flags) for i in b.info()], us=us.copy(), duals=b.duals()))
    yield sc, b, ctor, ticks
    b.close()


def test_constructor_solve_converges(mpc_run):
    sc, b, ctor, ticks = mpc_run
    for iters, conv, prim, dual in ctor:
        assert conv == 1 and iters < 20
        assert max(prim, dual) <= 1e-4  # MPCSettings::TOL of reference examples/go2_kinodynamics.py:98


def test_ticks_take_one_full_newton_step(mpc_run):
    sc, b, ctor, ticks = mpc_run
    for tk in ticks:
        for iters, alpha, trials, flags in tk['info']:
            assert iters == 1 and alpha == 1.0 and trials == 1 and flags == 0


def test_contact_forces_carry_the_weight(mpc_run):
    sc, b, ctor, ticks = mpc_run
    us = ticks[-1]['us']
    fz = us[:, 0, 2:12:3].sum(axis=1)
    assert np.all(np.abs(fz - sc.mass * 9.81) < 0.1 * sc.mass * 9.81)


def test_inactive_joint_limit_multipliers_are_zero(mpc_run):
    sc, b, ctor, ticks = mpc_run
    vs, lams = ticks[-1]['duals']
    # rows 0..11 of every stage are the joint-limit box; the nominal gait stays inside the limits
    assert np.abs(vs[:, :-1, :12]).max() == 0.0


def test_index_errors_mirror_reference_throws():
    orc = load_oracle()
    sc = pu.Go2Scenario(1, seed=0)
    o = SolverOpts()
    orc.orc_default_solver_opts(C.byref(o))
    b = pu.Backend(orc, "orc", sc.model, sc.ocpb, o, 1)
    s = (Slot * 1)(sc.gait_slot(0))
    assert orc.orc_set_slot(b.ctx, 0, 1, sc.T, s) != 0  # reference src/ocp-handler.cpp:74-77
    assert b"Stage index" in orc.orc_last_error()
    assert orc.orc_set_slot(b.ctx, 0, 2, 0, s) != 0
    out = np.zeros(sc.ndx)
    assert orc.orc_get_xdot(b.ctx, 0, 1, sc.T, P(out)) != 0
    b.close()
