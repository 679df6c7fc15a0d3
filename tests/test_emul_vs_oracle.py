"""Kernel LOGIC vs the oracle on the CPU-only box.

tests/host_emul/libemul.so is the device code of simple-mpc_b200/csrc/*.cuh compiled by g++ with every
parallel phase serialised over the thread index (par.cuh). It is a test harness, not a product path.
  * knots: the per-knot linearisation (kino_knot.cuh) must reproduce the oracle's LQ knots to rounding
    (two independent implementations of the same derivatives: compact world-frame inertias vs dense 6x6);
  * Riccati: fed with the ORACLE's knots, both sweeps (generic riccati.cuh and the TMA-staged riccati3.cuh)
    must return gains and steps that are BITWISE equal to the oracle's (same canonical operation order).
"""
import ctypes as C
import os
import numpy as np
import pytest
from smpc_ctypes import (ROOT, go2_like_model, go2_kino_ocp, talos_like_model, talos_kino_ocp, standing_slot, fk_numpy, x_ref_of, total_mass, default_opts, Slot, Terminal, SolveInfo, dbl_p)
from oracle_libs import load_oracle


def P(a):
    return a.ctypes.data_as(dbl_p)


GAIT = [0b1111] * 10 + [0b0110] * 30 + [0b1111] * 10 + [0b1001] * 30
TALOS_GAIT = [0b11] * 6 + [0b01] * 10 + [0b11] * 6 + [0b10] * 10  # short cycle: every phase inside T


def load_emul():
    path = os.path.join(ROOT, "tests", "host_emul", "libemul.so")
    if not os.path.exists(path):
        pytest.skip("tests/host_emul/libemul.so not built (run __graft_entry__.build())")
    lib = C.CDLL(path)
    lib.emul_last_error.restype = C.c_char_p
    return lib


def run_case(cone, robot="go2"):
    orc, emu = load_oracle(), load_emul()
    if robot == "go2":
        m, lim = go2_like_model()
        nq, nv, T, gait = m.nq, m.nv, 50, GAIT
        ocpb = go2_kino_ocp(m, lim, T=T, force_cone=cone, land_cstr=cone, terminal_constraint=cone)
    else:
        # talos-like biped, quad feet: 6-D contact wrenches as controls, 6-D LOCAL contact velocity rows, FramePlacement pose
        # cost; cone = True adds the 17-row wrench cones and drops the joint limits (the (56, 34, 46) Riccati shape)
        m, lim = talos_like_model()
        nq, nv, T, gait = m.nq, m.nv, 24, TALOS_GAIT
        ocpb = talos_kino_ocp(m, lim, T=T, force_cone=cone, kinematics_limits=not cone, terminal_constraint=cone)
    nk, fs = m.nfeet, ocpb.desc.force_size
    nx, ndx, nu = nq + nv, 2 * nv, nv - 6 + nk * fs
    ocp = ocpb.desc
    opts = default_opts(orc)
    ctx = C.c_void_p()
    assert orc.orc_create(C.byref(ctx), 1, C.byref(m), C.byref(ocp), C.byref(opts)) == 0
    ncmax = orc.orc_get_ncmax(ctx)
    xref = x_ref_of(m)
    feet = fk_numpy(m, xref[:nq])['feet']
    rng = np.random.default_rng(1)

    def gait_slot(i):
        mask, prev = gait[i % len(gait)], gait[(i - 1) % len(gait)]
        return standing_slot(m, ocpb, feet, contact_mask=mask, land_mask=mask & ~prev,
                             force_z=total_mass(m) * 9.81 / max(bin(mask).count('1'), 1))

    slots = (Slot * T)(*[gait_slot(i) for i in range(T)])
    orc.orc_set_slots(ctx, 0, 1, slots)
    term = Terminal()
    term.x_ref[:nx] = xref
    term.com_ref[:] = [0.0, 0.0, 0.3]
    term.tau = np.sqrt(0.3 / 9.81)
    orc.orc_set_terminal(ctx, 0, 1, C.byref(term))
    xs = np.tile(xref, (T + 1, 1)).copy()
    us = np.tile(np.array(slots[0].u_ref[:nu]), (T, 1)).copy()
    orc.orc_set_warmstart(ctx, 0, 1, P(xs), P(us))
    x0 = xref.copy()
    x0[nq:] += rng.normal(size=nv) * (0.1 if robot == "go2" else 0.03)
    x0[7:nq] += rng.normal(size=nq - 7) * (0.05 if robot == "go2" else 0.01)
    info = SolveInfo()
    orc.orc_solve(ctx, P(x0), 2)  # two iterations: non-trivial multipliers
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
    emu_knots, orc_knots = np.zeros((T + 1, size)), np.zeros((T + 1, size))
    ncs = np.zeros(T + 1, np.int32)
    gxs, gus, lpl = np.zeros((T + 1, ndx)), np.zeros((T, nu)), np.zeros((T + 1, ndx))
    for t in range(T + 1):
        knot, nc = np.zeros(size), C.c_int(0)
        gx, gu, lplus, xdot, sc = np.zeros(ndx), np.zeros(nu), np.zeros(ndx), np.zeros(ndx), np.zeros(8)
        u = us0[t] if t < T else np.zeros(nu)
        xn = xs0[t + 1] if t < T else xs0[t]
        lam_n = lams0[t + 1] if t < T else np.zeros(ndx)
        r = emu.emul_kino_knot(C.byref(m), C.byref(ocp), C.byref(slots[min(t, T - 1)]), C.byref(term), int(t == T), 1, P(xs0[t]), P(u),
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
        assert n == nc.value
        kn = orc_knots[t]
        Cm, Dm, dm = np.zeros((ncmax, ndx)), np.zeros((ncmax, nu)), np.zeros(ncmax)
        Cm[:n] = o['C'].ravel()[:n * ndx].reshape(n, ndx)
        dm[:n] = o['d'][:n]
        kn[oQ:oQ + ndx * ndx], kn[oq:oq + ndx] = o['Q'].ravel(), o['q']
        kn[oC:oC + ncmax * ndx], kn[od:od + ncmax] = Cm.ravel(), dm
        if t < T:
            Dm[:n] = o['D'].ravel()[:n * nu].reshape(n, nu)
            kn[oS:oS + ndx * nu], kn[oR:oR + nu * nu], kn[or_:or_ + nu] = o['S'].ravel(), o['R'].ravel(), o['r']
            kn[oA:oA + ndx * ndx], kn[oB:oB + ndx * nu], kn[of:of + ndx] = o['A'].ravel(), o['B'].ravel(), o['f']
            kn[oD:oD + ncmax * nu] = Dm.ravel()
        blocks = dict(Q=(oQ, ndx * ndx), q=(oq, ndx), C=(oC, ncmax * ndx), d=(od, ncmax))
        if t < T:
            blocks.update(S=(oS, ndx * nu), R=(oR, nu * nu), r=(or_, nu), A=(oA, ndx * ndx), B=(oB, ndx * nu), f=(of, ndx),
                          D=(oD, ncmax * nu))
        for k, (o0, sz) in blocks.items():
            a, b = knot[o0:o0 + sz], kn[o0:o0 + sz]
            den = np.abs(b).max()
            err = np.abs(a - b).max() / den if den > 0 else np.abs(a).max()
            worst[k] = max(worst.get(k, 0.0), err)
    # Newton step and gains of the oracle for that iteration
    odx, odu = np.zeros((T + 1, ndx)), np.zeros((T, nu))
    orc.orc_get_step(ctx, 0, P(odx), P(odu))
    Ko, ko, xs1, us1 = np.zeros((T, nu, ndx)), np.zeros((T, nu)), np.zeros((T + 1, nx)), np.zeros((T, nu))
    orc.orc_get_results(ctx, 0, 1, P(xs1), P(us1), P(Ko), P(ko))
    slay = np.zeros(7, np.int32)
    emu.emul_sol_layout(ndx, nu, ncmax, slay.ctypes.data_as(C.POINTER(C.c_int)))
    oK, ok, oZ, oz, oPt, opt, ssize = [int(v) for v in slay]

    def sweep(fn, knots):
        sol = np.zeros((T + 1, ssize))
        dxs, dus = np.zeros((T + 1, ndx)), np.zeros((T, nu))
        vnew, lnew = np.zeros((T + 1, ncmax)), np.zeros((T + 1, ndx))
        dphi, flag = C.c_double(0), C.c_int(0)
        rc = fn(T, ndx, nu, ncmax, P(knots), ncs.ctypes.data_as(C.POINTER(C.c_int)), C.c_double(mu), C.c_double(mu), P(sol), P(gxs),
                P(gus), P(lpl), P(dxs), P(dus), P(vnew), P(lnew), C.byref(dphi), C.byref(flag))
        assert rc == 0 and flag.value == 0
        K = np.array([sol[t][oK:oK + nu * ndx].reshape(nu, ndx) for t in range(T)])
        return K, dxs, dus

    out = dict(worst=worst, pinf=pinf, dinf=dinf, info=info, Ko=Ko, odx=odx, odu=odu)
    tuned = emu.emul_riccati3 if robot == "go2" else emu.emul_riccati4  # riccati3.cuh is instantiated for the Go2 shapes only
    out['generic_on_oracle'] = sweep(emu.emul_riccati, orc_knots)
    out['tma_on_oracle'] = sweep(tuned, orc_knots)
    out['two_column_on_oracle'] = sweep(emu.emul_riccati4, orc_knots)
    if robot == "go2":
        # the products that skip the structural zeros of the kinodynamics A, B (riccati3.cuh KinoRows): same operations in the same order
        out['kino_structure_on_oracle'] = sweep(emu.emul_riccati3_kino, orc_knots)
        if not cone:
            # no row depends on the control (D = 0): Sigma = mu I, the W / Sigma phases drop out (riccati3.cuh DZ)
            out['kino_structure_d0_on_oracle'] = sweep(emu.emul_riccati3_kino_d0, orc_knots)
    # the PARALLEL (leg-condensed) sweep of riccati_par.cuh on the same knots: tuned sweeps per leg + leg elements + condensation
    for legs in (3, 5, 10):
        out['parallel%d_on_oracle' % legs] = sweep(lambda *a, J=legs: emu.emul_riccati_parallel(*a, J), orc_knots)
    out['special_on_emul'] = sweep(tuned, emu_knots)
    if robot == "go2":
        # the forward sweep with ONE landing buffer (riccati_fwd.cuh, nbuf = 1: the variant large batches run)
        os.environ["SMPC_EMUL_FWD_NBUF"] = "1"
        out['tma_one_buffer_on_oracle'] = sweep(emu.emul_riccati3, orc_knots)
        del os.environ["SMPC_EMUL_FWD_NBUF"]
    # conditioning of the penalised recursion, measured: the oracle's own knots with every entry moved by at most ONE ULP
    # (relative 2^-52, the size of the difference between two correct evaluations of the same Jacobians)
    spread = []
    for seed in range(4):
        r = np.random.default_rng(100 + seed)
        pert = orc_knots * (1.0 + (2.0 ** -52) * r.uniform(-1, 1, size=orc_knots.shape))
        Kp, dxp, dup = sweep(tuned, pert)
        rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
        spread.append(dict(K=max(rel(Kp[t], Ko[t]) for t in range(T)), K0=rel(Kp[0], Ko[0]), du=rel(dup, odu), dx=rel(dxp, odx)))
    out['ulp_spread'] = spread
    orc.orc_destroy(ctx)
    return out


@pytest.fixture(scope="module", params=[(False, "go2"), (True, "go2"), (False, "talos"), (True, "talos")],
                ids=["joint-limits+contact-velocity", "cone+landing+dcm", "talos-6d-limits+contact-velocity", "talos-6d-wrench-cone+dcm"])
def case(request):
    return run_case(*request.param)


def test_knots_match_oracle_to_rounding(case):
    w = case['worst']
    for k in ['Q', 'S', 'R', 'q', 'r', 'A', 'B', 'C', 'D', 'd']:
        assert w[k] < 1e-12, (k, w[k])
    assert w['f'] < 1e-10  # f = slack + mu lam: a difference of nearly equal states


def test_infeasibility_scalars_match(case):
    assert abs(case['pinf'] - case['info'].prim_infeas) <= 1e-12 * max(1.0, case['info'].prim_infeas)
    assert abs(case['dinf'] - case['info'].dual_infeas) <= 1e-9 * max(1.0, case['info'].dual_infeas)


def test_riccati_bitwise(case):
    for name in ['generic_on_oracle', 'tma_on_oracle', 'two_column_on_oracle', 'kino_structure_on_oracle', 'kino_structure_d0_on_oracle', 'tma_one_buffer_on_oracle']:
        if name not in case:
            continue
        K, dxs, dus = case[name]
        assert np.array_equal(K, case['Ko']), name
        assert np.array_equal(dxs, case['odx']), name
        assert np.array_equal(dus, case['odu']), name


def test_parallel_sweep_matches_the_serial_sweep(case):
    """riccati_par.cuh (CPU emulation of the four passes) against the oracle's SERIAL sweep on identical knots: the hand-over
    of the cost-to-go between legs goes through the condensed system, so equality is to rounding amplified by the penalised
    recursion - the same floor two correct evaluations of the knots differ by (test_one_ulp_...)."""
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    for legs in (3, 5, 10):
        K, dxs, dus = case['parallel%d_on_oracle' % legs]
        T = K.shape[0]
        eK = max(rel(K[t], case['Ko'][t]) for t in range(T))
        print("legs %d: K %.2e dx %.2e du %.2e ; one-ulp spread of the serial sweep: K %.2e dx %.2e du %.2e" % (
            legs, eK, rel(dxs, case['odx']), rel(dus, case['odu']), max(s['K'] for s in case['ulp_spread']),
            max(s['dx'] for s in case['ulp_spread']), max(s['du'] for s in case['ulp_spread'])))
        # measured: 3 legs within the one-ulp spread of the serial sweep, 10 legs up to 20 x above it (K 2e-8, du 1e-7 at worst):
        # every hand-over adds a condensation step. Gains at the tolerance of the GPU parity tests (TOL_K), the step at 3 x that
        assert eK <= 1e-7 and rel(dxs, case['odx']) <= 3e-7 and rel(dus, case['odu']) <= 3e-7, (legs, eK, rel(dxs, case['odx']), rel(dus, case['odu']))


def test_end_to_end_gains_within_conditioning(case):
    K, dxs, dus = case['special_on_emul']
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(K, case['Ko']) < 1e-7
    assert rel(dus, case['odu']) < 1e-7


def test_one_ulp_on_the_knots_is_amplified_to_the_parity_floor(case):
    """The claim behind the GPU-parity tolerances (tests/test_gpu_parity.py): the Riccati sweep itself is
    bitwise reproducible, and what separates the CUDA gains from the oracle's is the 1-ulp difference between two independent
    evaluations of the knot Jacobians, amplified by the 1/mu = 1e8 conditioning of the penalised recursion. Measured here
    instead of asserted: moving every entry of the ORACLE's knots by at most one ulp (relative 2^-52 = 2.2e-16) and
    re-running the SAME sweep changes the gains by 1e-12 .. 1e-7 relative (normwise per stage) - four to nine orders of
    magnitude above the perturbation, and the very band in which the CUDA-vs-oracle differences sit."""
    eps = 2.0 ** -52
    ks = [sp['K'] for sp in case['ulp_spread']]
    dus = [sp['du'] for sp in case['ulp_spread']]
    print("1-ulp knot perturbation -> gain spread", ["%.2e" % v for v in ks], "step spread", ["%.2e" % v for v in dus])
    assert min(ks) > 1e3 * eps, ks            # the amplification is real: >= 1000 x the perturbation ...
    assert max(ks) < 1e-6 and max(dus) < 1e-6  # ... and bounded: the 1e-7 gain tolerance of the GPU tests covers it
