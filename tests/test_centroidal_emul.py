"""Centroidal problem: kernel LOGIC vs the oracle on the CPU-only box.

The device code of simple-mpc_b200/csrc/centro_knot.cuh + dense_knot.cuh is compiled by g++ through the
par.cuh emulation (tests/host_emul, a test harness) and must reproduce the oracle's LQ knots, per-knot
scalars and merit-gradient pieces at a non-trivial iterate (wrench-cone rows active and inactive, single
and double support); the generic Riccati sweep fed with the oracle's knots must return bitwise the
oracle's gains.
"""
import ctypes as C
import numpy as np
import pytest
import parity_utils as pu
from smpc_ctypes import default_opts, Slot, SolveInfo, dbl_p
from oracle_libs import load_oracle
from test_emul_vs_oracle import load_emul, P


@pytest.fixture(scope="module")
def case():
    orc, emu = load_oracle(), load_emul()
    sc = pu.TalosCentroidalScenario(1, seed=5)
    m, ocp, T, nx, ndx, nu = sc.model, sc.ocpb.desc, sc.T, 9, 9, 12
    opts = default_opts(orc)
    ctx = C.c_void_p()
    assert orc.orc_create(C.byref(ctx), 1, C.byref(m), C.byref(ocp), C.byref(opts)) == 0
    ncmax = orc.orc_get_ncmax(ctx)
    assert ncmax == 34
    slots = (Slot * T)(*[sc.gait_slot(i) for i in range(T)])  # 20 DS, 80 SS: both stage types in the horizon
    orc.orc_set_slots(ctx, 0, 1, slots)
    term = sc.terminals()
    orc.orc_set_terminal(ctx, 0, 1, term)
    xs, us = sc.warmstart()
    orc.orc_set_warmstart(ctx, 0, 1, P(xs), P(us))
    x0 = sc.x_start()[0]
    x0[3:9] += [10.0, 25.0, 0.0, 5.0, 5.0, 1.0]  # a push large enough to put forces on the wrench-cone faces
    info = SolveInfo()
    assert orc.orc_solve(ctx, P(x0), 2) == 0
    orc.orc_get_info(ctx, 0, 1, C.byref(info))
    preg, mu = info.preg, 1e-8
    xs0, us0 = np.zeros((T + 1, nx)), np.zeros((T, nu))
    orc.orc_get_results(ctx, 0, 1, P(xs0), P(us0), None, None)
    vs0, lams0 = np.zeros((T + 1, ncmax)), np.zeros((T + 1, ndx))
    orc.orc_get_duals(ctx, 0, 1, P(vs0), P(lams0))
    assert orc.orc_solve(ctx, P(xs0[0].copy()), 1) == 0
    orc.orc_get_info(ctx, 0, 1, C.byref(info))
    lay = np.zeros(12, np.int32)
    emu.emul_knot_layout(ndx, nu, ncmax, lay.ctypes.data_as(C.POINTER(C.c_int)))
    oQ, oS, oR, oq, or_, oA, oB, of, oC, oD, od, size = [int(v) for v in lay]
    names = ['Q', 'S', 'R', 'q', 'r', 'A', 'B', 'f', 'C', 'D', 'd']
    worst, pinf, dinf, cost = {}, 0.0, 0.0, 0.0
    orc_knots = np.zeros((T + 1, size))
    ncs = np.zeros(T + 1, np.int32)
    gxs, gus, lpl = np.zeros((T + 1, ndx)), np.zeros((T, nu)), np.zeros((T + 1, ndx))
    n_active = 0
    for t in range(T + 1):
        knot, nc = np.zeros(size), C.c_int(0)
        gx, gu, lplus, xdot, scal = np.zeros(ndx), np.zeros(nu), np.zeros(ndx), np.zeros(ndx), np.zeros(8)
        u = us0[t] if t < T else np.zeros(nu)
        xn = xs0[t + 1] if t < T else xs0[t]
        lam_n = lams0[t + 1] if t < T else np.zeros(ndx)
        r = emu.emul_centro_knot(C.byref(m), C.byref(ocp), C.byref(slots[min(t, T - 1)]), term, int(t == T), 1, P(xs0[t]), P(u),
                                 P(xn), P(vs0[t]), P(vs0[t]), P(lams0[t]), P(lam_n), P(lam_n), C.c_double(mu), C.c_double(preg),
                                 P(knot), C.byref(nc), P(gx), P(gu), P(lplus), P(xdot), P(scal))
        assert r == 0, emu.emul_last_error()
        ncs[t], gxs[t] = nc.value, gx
        if t < T:
            gus[t], lpl[t + 1] = gu, lplus
        cost += scal[0]
        pinf = max(pinf, scal[2])
        dinf = max(dinf, scal[7] if t < T else 0, scal[6] if t > 0 else 0)
        o = dict(Q=np.zeros((ndx, ndx)), S=np.zeros((ndx, nu)), R=np.zeros((nu, nu)), q=np.zeros(ndx), r=np.zeros(nu),
                 A=np.zeros((ndx, ndx)), B=np.zeros((ndx, nu)), f=np.zeros(ndx), C=np.zeros((ncmax, ndx)), D=np.zeros((ncmax, nu)),
                 d=np.zeros(ncmax))
        onc = C.c_int(0)
        assert orc.orc_get_knot(ctx, 0, t, C.byref(onc), *[P(o[k]) for k in names]) == 0
        n = onc.value
        assert n == nc.value
        assert n == (0 if t == T else 17 * bin(slots[t].contact_mask).count("1"))
        kn = orc_knots[t]
        Cm, Dm, dm = np.zeros((ncmax, ndx)), np.zeros((ncmax, nu)), np.zeros(ncmax)
        Cm[:n] = o['C'].ravel()[:n * ndx].reshape(n, ndx)
        dm[:n] = o['d'][:n]
        kn[oQ:oQ + ndx * ndx], kn[oq:oq + ndx] = o['Q'].ravel(), o['q']
        kn[oC:oC + ncmax * ndx], kn[od:od + ncmax] = Cm.ravel(), dm
        blocks = dict(Q=(oQ, ndx * ndx), q=(oq, ndx), C=(oC, ncmax * ndx), d=(od, ncmax))
        if t < T:
            Dm[:n] = o['D'].ravel()[:n * nu].reshape(n, nu)
            n_active += int((np.abs(Dm).sum(axis=1) > 0).sum())
            kn[oS:oS + ndx * nu], kn[oR:oR + nu * nu], kn[or_:or_ + nu] = o['S'].ravel(), o['R'].ravel(), o['r']
            kn[oA:oA + ndx * ndx], kn[oB:oB + ndx * nu], kn[of:of + ndx] = o['A'].ravel(), o['B'].ravel(), o['f']
            kn[oD:oD + ncmax * nu] = Dm.ravel()
            blocks.update(S=(oS, ndx * nu), R=(oR, nu * nu), r=(or_, nu), A=(oA, ndx * ndx), B=(oB, ndx * nu), f=(of, ndx),
                          D=(oD, ncmax * nu))
        for k, (o0, sz) in blocks.items():
            a, b = knot[o0:o0 + sz], kn[o0:o0 + sz]
            den = np.abs(b).max()
            err = np.abs(a - b).max() / den if den > 0 else np.abs(a).max()
            worst[k] = max(worst.get(k, 0.0), err)
    odx, odu = np.zeros((T + 1, ndx)), np.zeros((T, nu))
    orc.orc_get_step(ctx, 0, P(odx), P(odu))
    Ko, ko, xs1, us1 = np.zeros((T, nu, ndx)), np.zeros((T, nu)), np.zeros((T + 1, nx)), np.zeros((T, nu))
    orc.orc_get_results(ctx, 0, 1, P(xs1), P(us1), P(Ko), P(ko))
    slay = np.zeros(7, np.int32)
    emu.emul_sol_layout(ndx, nu, ncmax, slay.ctypes.data_as(C.POINTER(C.c_int)))
    oK, ssize = int(slay[0]), int(slay[6])
    sol = np.zeros((T + 1, ssize))
    dxs, dus = np.zeros((T + 1, ndx)), np.zeros((T, nu))
    vnew, lnew = np.zeros((T + 1, ncmax)), np.zeros((T + 1, ndx))
    dphi, flag = C.c_double(0), C.c_int(0)
    rc = emu.emul_riccati(T, ndx, nu, ncmax, P(orc_knots), ncs.ctypes.data_as(C.POINTER(C.c_int)), C.c_double(mu), C.c_double(mu),
                          P(sol), P(gxs), P(gus), P(lpl), P(dxs), P(dus), P(vnew), P(lnew), C.byref(dphi), C.byref(flag))
    assert rc == 0 and flag.value == 0
    K = np.array([sol[t][oK:oK + nu * ndx].reshape(nu, ndx) for t in range(T)])
    # the two-column sweep in padded mode (riccati4.cuh: state dimension 9 computed on 12 x 12 blocks with exact-zero padding)
    sol4 = np.zeros((T + 1, ssize))
    dxs4, dus4 = np.zeros((T + 1, ndx)), np.zeros((T, nu))
    rc = emu.emul_riccati4(T, ndx, nu, ncmax, P(orc_knots), ncs.ctypes.data_as(C.POINTER(C.c_int)), C.c_double(mu), C.c_double(mu),
                           P(sol4), P(gxs), P(gus), P(lpl), P(dxs4), P(dus4), P(vnew), P(lnew), C.byref(dphi), C.byref(flag))
    assert rc == 0 and flag.value == 0, emu.emul_last_error()
    K4 = np.array([sol4[t][oK:oK + nu * ndx].reshape(nu, ndx) for t in range(T)])
    orc.orc_destroy(ctx)
    return dict(K4=K4, dxs4=dxs4, dus4=dus4, worst=worst, pinf=pinf, dinf=dinf, cost=cost, info=info, K=K, Ko=Ko, dxs=dxs, dus=dus, odx=odx, odu=odu,
                n_active=n_active)


def test_centroidal_knots_match_oracle(case):
    for k in ['Q', 'S', 'R', 'q', 'r', 'A', 'B', 'C', 'D', 'd']:
        assert case['worst'][k] < 1e-12, (k, case['worst'][k])
    assert case['worst']['f'] < 1e-10
    assert case['n_active'] > 0  # the iterate exercises active wrench-cone rows


def test_centroidal_scalars_match_oracle(case):
    assert abs(case['pinf'] - case['info'].prim_infeas) <= 1e-12 * max(1.0, case['info'].prim_infeas)
    assert abs(case['dinf'] - case['info'].dual_infeas) <= 1e-9 * max(1.0, case['info'].dual_infeas)


def test_centroidal_riccati_bitwise(case):
    assert np.array_equal(case['K'], case['Ko'])
    assert np.array_equal(case['dxs'], case['odx'])
    assert np.array_equal(case['dus'], case['odu'])


def test_centroidal_padded_two_column_riccati_bitwise(case):
    assert np.array_equal(case['K4'], case['Ko'])
    assert np.array_equal(case['dxs4'], case['odx'])
    assert np.array_equal(case['dus4'], case['odu'])


def test_centroidal_oracle_solve_converges():
    """standing double support from a perturbed centroidal state: the cold solve (<= 100 iterations, reference
    src/mpc.cpp:17,85) must converge with forces inside the wrench cones and summing to the weight"""
    orc = load_oracle()
    sc = pu.TalosCentroidalScenario(1, seed=2)
    b = pu.Backend(orc, "orc", sc.model, sc.ocpb, default_opts(orc), 1)
    T = sc.T
    s = sc.gait_slot(0)
    b.set_slots((Slot * T)(*([s] * T)))
    b.set_terminal(sc.terminals())
    b.set_warmstart(*sc.warmstart())
    x0 = sc.xref.copy()
    x0[3:6] += [2.0, -1.0, 0.5]
    b.solve(x0[None, :], 100)
    info = b.info()[0]
    xs, us, Ks, ks = b.results()
    assert info.prim_infeas < 1e-3, (info.prim_infeas, info.dual_infeas, info.num_iters)
    fz = us[0, :, 2] + us[0, :, 8]
    assert np.all(us[0, :, 2] > 0) and np.all(us[0, :, 8] > 0)
    assert abs(fz[-1] - sc.mass * 9.81) < 0.05 * sc.mass * 9.81
    b.close()
