"""Full-dynamics kernel LOGIC vs the oracle on the CPU-only box (see test_emul_vs_oracle.py for the harness):
the per-knot linearisation of csrc/full_knot.cuh (rigid-contact forward dynamics, its derivatives, costs, constraints,
multiplier projection) must reproduce the oracle's LQ knots to rounding, and the generic Riccati sweep fed with those
knots must return the oracle's gains within the conditioning of the penalised recursion."""
import ctypes as C
import numpy as np
import pytest
from smpc_ctypes import (go2_like_model, go2_full_ocp, full_slot, fk_numpy, x_ref_of, default_opts, Slot, Terminal, SolveInfo, dbl_p,
                         talos_like_model, talos_full_ocp)
from oracle_libs import load_oracle
from test_emul_vs_oracle import load_emul, GAIT


def P(a):
    return a.ctypes.data_as(dbl_p)


TALOS_GAIT = [0b11] * 6 + [0b01] * 10 + [0b11] * 6 + [0b10] * 10  # short cycle: every phase and both landing edges inside T


def run_case(cone, kp, robot="go2"):
    orc, emu = load_oracle(), load_emul()
    if robot == "go2":
        m, lim = go2_like_model()
        T, gait = 50, GAIT
        nq, nv = m.nq, m.nv
        ocpb = go2_full_ocp(m, lim, T=T, force_cone=cone, land_cstr=cone, terminal_constraint=cone,
                            Kp=(kp, kp, 10 * kp), Kd=(kp, kp, 2 * kp))
    else:
        # config C4: talos-like biped, 6-D contacts (CONTACT_6D, LOCAL_WORLD_ALIGNED), wrench cones, 6-D landing velocity
        m, lim = talos_like_model()
        T, gait = 24, TALOS_GAIT
        nq, nv = m.nq, m.nv
        ocpb = talos_full_ocp(m, lim, T=T, force_cone=cone, land_cstr=cone, terminal_constraint=cone,
                              Kp=(kp, 0.5 * kp, 10 * kp, 0.3 * kp, 0.2 * kp, kp), Kd=(20.0 * kp,) * 6 if kp else (100.0,) * 6)
    nx, ndx, nu = nq + nv, 2 * nv, nv - 6
    ocp = ocpb.desc
    opts = default_opts(orc)
    ctx = C.c_void_p()
    assert orc.orc_create(C.byref(ctx), 1, C.byref(m), C.byref(ocp), C.byref(opts)) == 0, orc.orc_last_error()
    ncmax = orc.orc_get_ncmax(ctx)
    xref = x_ref_of(m)
    feet = fk_numpy(m, xref[:nq])['feet']
    rng = np.random.default_rng(1)

    def gait_slot(i):
        mask, prev = gait[i % len(gait)], gait[(i - 1) % len(gait)]
        return full_slot(m, ocpb, feet, contact_mask=mask, land_mask=mask & ~prev)

    slots = (Slot * T)(*[gait_slot(i) for i in range(T)])
    orc.orc_set_slots(ctx, 0, 1, slots)
    term = Terminal()
    term.x_ref[:nx] = xref
    term.com_ref[:] = [0.0, 0.0, 0.3]
    term.tau = np.sqrt(0.3 / 9.81)
    orc.orc_set_terminal(ctx, 0, 1, C.byref(term))
    xs = np.tile(xref, (T + 1, 1)).copy()
    us = np.zeros((T, nu))
    orc.orc_set_warmstart(ctx, 0, 1, P(xs), P(us))
    x0 = xref.copy()
    x0[nq:] += rng.normal(size=nv) * (0.1 if robot == "go2" else 0.03)
    x0[7:nq] += rng.normal(size=nq - 7) * (0.05 if robot == "go2" else 0.01)
    info = SolveInfo()
    assert orc.orc_solve(ctx, P(x0), 2) == 0, orc.orc_last_error()  # two iterations: non-trivial multipliers
    orc.orc_get_info(ctx, 0, 1, C.byref(info))
    preg, mu = info.preg, 1e-8
    xs0, us0 = np.zeros((T + 1, nx)), np.zeros((T, nu))
    orc.orc_get_results(ctx, 0, 1, P(xs0), P(us0), None, None)
    vs0, lams0 = np.zeros((T + 1, ncmax)), np.zeros((T + 1, ndx))
    orc.orc_get_duals(ctx, 0, 1, P(vs0), P(lams0))
    orc.orc_solve(ctx, P(xs0[0].copy()), 1)  # the iteration whose LQ problem is compared
    orc.orc_get_info(ctx, 0, 1, C.byref(info))
    lay = np.zeros(12, np.int32)
    emu.emul_knot_layout(ndx, nu, ncmax, lay.ctypes.data_as(C.POINTER(C.c_int)))
    oQ, oS, oR, oq, or_, oA, oB, of, oC, oD, od, size = [int(v) for v in lay]
    names = ['Q', 'S', 'R', 'q', 'r', 'A', 'B', 'f', 'C', 'D', 'd']
    worst, pinf, dinf = {}, 0.0, 0.0
    emu_knots = np.zeros((T + 1, size))
    ncs = np.zeros(T + 1, np.int32)
    gxs, gus, lpl = np.zeros((T + 1, ndx)), np.zeros((T, nu)), np.zeros((T + 1, ndx))
    for t in range(T + 1):
        knot, nc = np.zeros(size), C.c_int(0)
        gx, gu, lplus, xdot, sc = np.zeros(ndx), np.zeros(nu), np.zeros(ndx), np.zeros(ndx), np.zeros(8)
        u = us0[t] if t < T else np.zeros(nu)
        xn = xs0[t + 1] if t < T else xs0[t]
        lam_n = lams0[t + 1] if t < T else np.zeros(ndx)
        r = emu.emul_full_knot(C.byref(m), C.byref(ocp), C.byref(slots[min(t, T - 1)]), C.byref(term), int(t == T), 1, P(xs0[t]), P(u),
                               P(xn), P(vs0[t]), P(vs0[t]), P(lams0[t]), P(lam_n), P(lam_n), C.c_double(mu), C.c_double(preg),
                               P(knot), C.byref(nc), P(gx), P(gu), P(lplus), P(xdot), P(sc))
        assert r == 0, emu.emul_last_error()
        emu_knots[t], ncs[t], gxs[t] = knot, nc.value, gx
        if t < T:
            gus[t], lpl[t + 1] = gu, lplus
        pinf = max(pinf, sc[2])
        dinf = max(dinf, sc[7] if t < T else 0, sc[6] if t > 0 else 0)
        o = dict(Q=np.zeros((ndx, ndx)), S=np.zeros((ndx, nu)), R=np.zeros((nu, nu)), q=np.zeros(ndx), r=np.zeros(nu),
                 A=np.zeros((ndx, ndx)), B=np.zeros((ndx, nu)), f=np.zeros(ndx), C=np.zeros((ncmax, ndx)), D=np.zeros((ncmax, nu)),
                 d=np.zeros(ncmax))
        onc = C.c_int(0)
        assert orc.orc_get_knot(ctx, 0, t, C.byref(onc), *[P(o[k]) for k in names]) == 0
        n = onc.value
        assert n == nc.value, (t, n, nc.value)
        kn = np.zeros(size)
        Cm, Dm, dm = np.zeros((ncmax, ndx)), np.zeros((ncmax, nu)), np.zeros(ncmax)
        Cm[:n] = o['C'].ravel()[:n * ndx].reshape(n, ndx)
        dm[:n] = o['d'][:n]
        kn[oQ:oQ + ndx * ndx], kn[oq:oq + ndx] = o['Q'].ravel(), o['q']
        kn[oC:oC + ncmax * ndx], kn[od:od + ncmax] = Cm.ravel(), dm
        blocks = dict(Q=(oQ, ndx * ndx), q=(oq, ndx), C=(oC, ncmax * ndx), d=(od, ncmax))
        if t < T:
            Dm[:n] = o['D'].ravel()[:n * nu].reshape(n, nu)
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
    rc = emu.emul_riccati(T, ndx, nu, ncmax, P(emu_knots), ncs.ctypes.data_as(C.POINTER(C.c_int)), C.c_double(mu), C.c_double(mu), P(sol),
                          P(gxs), P(gus), P(lpl), P(dxs), P(dus), P(vnew), P(lnew), C.byref(dphi), C.byref(flag))
    assert rc == 0 and flag.value == 0
    K = np.array([sol[t][oK:oK + nu * ndx].reshape(nu, ndx) for t in range(T)])
    # the shape-specialised two-column sweep (riccati4.cuh) on the ORACLE's knots: bitwise the oracle's gains and steps
    orc_knots = np.zeros((T + 1, size))
    for t in range(T + 1):
        o = dict(Q=np.zeros((ndx, ndx)), S=np.zeros((ndx, nu)), R=np.zeros((nu, nu)), q=np.zeros(ndx), r=np.zeros(nu),
                 A=np.zeros((ndx, ndx)), B=np.zeros((ndx, nu)), f=np.zeros(ndx), C=np.zeros((ncmax, ndx)), D=np.zeros((ncmax, nu)),
                 d=np.zeros(ncmax))
        onc = C.c_int(0)
        assert orc.orc_get_knot(ctx, 0, t, C.byref(onc), *[P(o[k]) for k in names]) == 0
        n = onc.value
        kn = orc_knots[t]
        kn[oQ:oQ + ndx * ndx], kn[oq:oq + ndx] = o['Q'].ravel(), o['q']
        kn[oC:oC + n * ndx], kn[od:od + n] = o['C'].ravel()[:n * ndx], o['d'][:n]
        if t < T:
            kn[oS:oS + ndx * nu], kn[oR:oR + nu * nu], kn[or_:or_ + nu] = o['S'].ravel(), o['R'].ravel(), o['r']
            kn[oA:oA + ndx * ndx], kn[oB:oB + ndx * nu], kn[of:of + ndx] = o['A'].ravel(), o['B'].ravel(), o['f']
            kn[oD:oD + n * nu] = o['D'].ravel()[:n * nu]
    sol4 = np.zeros((T + 1, ssize))
    dxs4, dus4 = np.zeros((T + 1, ndx)), np.zeros((T, nu))
    rc = emu.emul_riccati4(T, ndx, nu, ncmax, P(orc_knots), ncs.ctypes.data_as(C.POINTER(C.c_int)), C.c_double(mu), C.c_double(mu), P(sol4),
                           P(gxs), P(gus), P(lpl), P(dxs4), P(dus4), P(vnew), P(lnew), C.byref(dphi), C.byref(flag))
    assert rc == 0 and flag.value == 0, emu.emul_last_error()
    K4 = np.array([sol4[t][oK:oK + nu * ndx].reshape(nu, ndx) for t in range(T)])
    orc.orc_destroy(ctx)
    return dict(worst=worst, pinf=pinf, dinf=dinf, info=info, Ko=Ko, odx=odx, odu=odu, K=K, dxs=dxs, dus=dus, K4=K4, dxs4=dxs4, dus4=dus4)


@pytest.fixture(scope="module", params=[(False, 0.0), (True, 5.0), (True, 0.0, "talos"), (True, 5.0, "talos")],
                ids=["torque+joint-limits", "cone+landing+dcm+baumgarte", "talos-6d-wrench-cone", "talos-6d-cone+landing+dcm+baumgarte"])
def case(request):
    return run_case(*request.param)


def test_full_knots_match_oracle_to_rounding(case):
    w = case['worst']
    for k in ['Q', 'S', 'R', 'q', 'r', 'A', 'B', 'C', 'D', 'd']:
        assert w[k] < 1e-10, (k, w[k])
    assert w['f'] < 1e-9


def test_full_infeasibility_scalars_match(case):
    assert abs(case['pinf'] - case['info'].prim_infeas) <= 1e-10 * max(1.0, case['info'].prim_infeas)
    assert abs(case['dinf'] - case['info'].dual_infeas) <= 1e-8 * max(1.0, case['info'].dual_infeas)


def test_full_end_to_end_gains_within_conditioning(case):
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(case['K'], case['Ko']) < 1e-6
    assert rel(case['dus'], case['odu']) < 1e-6


def test_full_two_column_riccati_bitwise(case):
    """riccati4.cuh (generic shapes, two columns per factorisation step) fed with the ORACLE's knots returns the oracle's
    gains and Newton step bit for bit, for the Go2 full-dynamics shapes (36, 12, 24 | 60) and the Talos ones (56, 22, 78 | 90)."""
    assert np.array_equal(case['K4'], case['Ko'])
    assert np.array_equal(case['dxs4'], case['odx'])
    assert np.array_equal(case['dus4'], case['odu'])
