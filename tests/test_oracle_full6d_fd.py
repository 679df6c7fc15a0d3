"""Pins the oracle's 6-D rigid contacts (CONTACT_6D, LOCAL_WORLD_ALIGNED; oracle/orc_full.hpp) and the full-dynamics
stage of the talos-like biped (config C4 of SURVEY.md §8d; reference src/fulldynamics.cpp:56-64,97-114,163-181) with the
ground truth available here (Pinocchio/Aligator absent, no golden vector):
  * the contact equation itself, re-evaluated from finite differences of plain numpy forward kinematics along the
    trajectory q(t) = q (+) (v t + a t^2 / 2): spatial acceleration of the contact frame + Kd o velocity + Kp o log6(pose) = 0;
  * da/dq, da/dv, da/dtau, dlam/d. against central finite differences through the manifold retraction (this is what
    checks the extra RNEA term of a force held in world-aligned axes);
  * A, B, cost gradients and constraint Jacobians (wrench cone on lam, 6-D landing velocity, FramePlacement cost,
    6-D contact-force cost) of the stage against finite differences.
"""
import ctypes as C
import numpy as np
import pytest
from smpc_ctypes import talos_like_model, talos_full_ocp, full_slot, fk_numpy, x_ref_of, dbl_p
from oracle_libs import load_oracle

KP = (3.0, 5.0, 50.0, 2.0, 4.0, 6.0)
KD = (1.0, 2.0, 10.0, 3.0, 1.5, 2.5)


def P(a):
    return a.ctypes.data_as(dbl_p)


@pytest.fixture(scope="module")
def env():
    orc = load_oracle()
    m, lim = talos_like_model()
    ocpb = talos_full_ocp(m, lim, force_cone=True, land_cstr=True, Kp=KP, Kd=KD)
    return orc, m, lim, ocpb


def integrate(orc, m, ocp, x, dx):
    xo = np.zeros(m.nq + m.nv)
    assert orc.orc_state_integrate(C.byref(m), C.byref(ocp), P(x), P(np.ascontiguousarray(dx)), P(xo)) == 0
    return xo


def difference(orc, m, ocp, x0, x1):
    d = np.zeros(2 * m.nv)
    assert orc.orc_state_difference(C.byref(m), C.byref(ocp), P(x0), P(x1), P(d)) == 0
    return d


def random_state(orc, m, ocp, rng, scale=1.0):
    x = x_ref_of(m)
    nj = m.nv - 6
    return integrate(orc, m, ocp, x, scale * np.concatenate([rng.normal(size=6) * 0.2, rng.normal(size=nj) * 0.2, rng.normal(size=m.nv) * 0.5]))


def fdyn(orc, m, ocp, x, tau, mask, derivs=1, Kp=KP, Kd=KD):
    nq, nv = m.nq, m.nv
    ncon = 6 * bin(mask).count("1")
    o = dict(a=np.zeros(nv), lam=np.zeros(max(ncon, 1)), da_dq=np.zeros((nv, nv)), da_dv=np.zeros((nv, nv)), da_dtau=np.zeros((nv, nv)),
             dl_dq=np.zeros((max(ncon, 1), nv)), dl_dv=np.zeros((max(ncon, 1), nv)), dl_dtau=np.zeros((max(ncon, 1), nv)))
    g = np.array(ocp.gravity[:])
    Kp, Kd = np.asarray(Kp, float), np.asarray(Kd, float)
    q, v, tau = [np.ascontiguousarray(z) for z in (x[:nq], x[nq:], tau)]
    r = orc.orc_full_dynamics(C.byref(m), P(q), P(v), P(tau), C.c_uint(mask), P(g), P(Kp), P(Kd), derivs,
                              *[P(o[k]) for k in ['a', 'lam', 'da_dq', 'da_dv', 'da_dtau', 'dl_dq', 'dl_dv', 'dl_dtau']], 6)
    assert r == 0, orc.orc_last_error()
    o['lam'] = o['lam'][:ncon]
    for k in ['dl_dq', 'dl_dv', 'dl_dtau']:
        o[k] = o[k][:ncon]
    return o


def log6(orc, R, p):
    pose = np.concatenate([np.asarray(R).ravel(), np.asarray(p)])
    nu = np.zeros(6)
    orc.orc_log6(P(pose), P(nu))
    return nu


@pytest.mark.parametrize("mask", [0b11, 0b01, 0b10])
def test_contact6d_dynamics_satisfies_the_constraint(env, mask):
    orc, m, lim, ocpb = env
    ocp = ocpb.desc
    nq, nv = m.nq, m.nv
    rng = np.random.default_rng(40 + mask)
    x = random_state(orc, m, ocp, rng, 0.5)
    tau = np.concatenate([np.zeros(6), rng.normal(size=nv - 6) * 20])
    o = fdyn(orc, m, ocp, x, tau, mask, derivs=0)
    a = o['a']
    h = 1e-4

    def pose(f, t):
        xt = integrate(orc, m, ocp, x, np.concatenate([x[nq:] * t + 0.5 * a * t * t, np.zeros(nv)]))
        return fk_numpy(m, xt[:nq])['feet'][f]

    def local_twist(f, t):
        (Rm, pm), (Rp, pp) = pose(f, t - h), pose(f, t + h)
        return log6(orc, Rm.T @ Rp, Rm.T @ (pp - pm)) / (2 * h)

    for f in range(2):
        if not mask & (1 << f):
            continue
        R, p = pose(f, 0.0)
        v_loc = local_twist(f, 0.0)
        a_loc = (local_twist(f, h) - local_twist(f, -h)) / (2 * h)
        e = log6(orc, R, p)
        c_loc = a_loc + np.array(KD) * v_loc + np.array(KP) * e
        # O(h) consistency of the doubly finite-differenced twist and of the second-order retraction
        assert np.abs(c_loc).max() < 2e-3 * max(1.0, np.abs(a_loc).max()), c_loc


def test_contact6d_forces_balance_gravity(env):
    """Momentum balance with 6-D contacts: hdot = sum of the world-aligned contact forces + m g (linear part)."""
    orc, m, lim, ocpb = env
    ocp = ocpb.desc
    rng = np.random.default_rng(3)
    x = random_state(orc, m, ocp, rng, 0.3)
    tau = np.concatenate([np.zeros(6), rng.normal(size=m.nv - 6) * 10])
    o = fdyn(orc, m, ocp, x, tau, 0b11, derivs=0)
    nv = m.nv
    out = dict(foot=np.zeros(48), fref=np.zeros(48), base=np.zeros(12), com=np.zeros(3), hg=np.zeros(6), hdg=np.zeros(6),
               Ag=np.zeros(6 * nv), dh_dq=np.zeros(6 * nv), dhd_dq=np.zeros(6 * nv), dhd_dv=np.zeros(6 * nv), Jcom=np.zeros(3 * nv))
    keys = ['foot', 'fref', 'base', 'com', 'hg', 'hdg', 'Ag', 'dh_dq', 'dhd_dq', 'dhd_dv', 'Jcom']
    assert orc.orc_kinematics(C.byref(m), P(x[:m.nq].copy()), P(x[m.nq:].copy()), P(o['a']), *[P(out[k]) for k in keys]) == 0
    mass = sum(m.mass[i] for i in range(m.njoints))
    feet = fk_numpy(m, x[:m.nq])['feet']
    lin = sum(o['lam'][6 * f:6 * f + 3] for f in range(2))
    ang = sum(o['lam'][6 * f + 3:6 * f + 6] + np.cross(feet[f][1] - out['com'], o['lam'][6 * f:6 * f + 3]) for f in range(2))
    assert np.abs(out['hdg'][:3] - (lin + mass * np.array(ocp.gravity[:]))).max() < 1e-8 * mass * 9.81
    assert np.abs(out['hdg'][3:] - ang).max() < 1e-8 * mass * 9.81


@pytest.mark.parametrize("mask", [0b11, 0b01, 0b10])
def test_contact6d_dynamics_derivatives(env, mask):
    orc, m, lim, ocpb = env
    ocp = ocpb.desc
    nq, nv = m.nq, m.nv
    rng = np.random.default_rng(50 + mask)
    x = random_state(orc, m, ocp, rng)
    tau = np.concatenate([rng.normal(size=6), rng.normal(size=nv - 6) * 20])
    o = fdyn(orc, m, ocp, x, tau, mask)
    h = 1e-6
    for key, blk in [('q', 0), ('v', nv)]:
        da, dl = np.zeros((nv, nv)), np.zeros((o['lam'].size, nv))
        for j in range(nv):
            e = np.zeros(2 * nv)
            e[blk + j] = h
            op = fdyn(orc, m, ocp, integrate(orc, m, ocp, x, e), tau, mask, 0)
            om = fdyn(orc, m, ocp, integrate(orc, m, ocp, x, -e), tau, mask, 0)
            da[:, j] = (op['a'] - om['a']) / (2 * h)
            dl[:, j] = (op['lam'] - om['lam']) / (2 * h)
        sa, sl = max(1.0, np.abs(da).max()), max(1.0, np.abs(dl).max(initial=0.0))
        assert np.abs(da - o['da_d' + key]).max() < 5e-7 * sa, (key, np.abs(da - o['da_d' + key]).max(), sa)
        assert np.abs(dl - o['dl_d' + key]).max(initial=0.0) < 5e-7 * sl, (key, np.abs(dl - o['dl_d' + key]).max(), sl)
    da, dl = np.zeros((nv, nv)), np.zeros((o['lam'].size, nv))
    for j in range(nv):
        e = np.zeros(nv)
        e[j] = 1e-3
        op, om = fdyn(orc, m, ocp, x, tau + e, mask, 0), fdyn(orc, m, ocp, x, tau - e, mask, 0)
        da[:, j] = (op['a'] - om['a']) / 2e-3
        dl[:, j] = (op['lam'] - om['lam']) / 2e-3
    assert np.abs(da - o['da_dtau']).max() < 1e-8 * max(1.0, np.abs(da).max())
    assert np.abs(dl - o['dl_dtau']).max(initial=0.0) < 1e-8 * max(1.0, np.abs(dl).max(initial=0.0))


def stage(orc, m, ocp, slot, x, u, derivs=1):
    nx, ndx, nu, ncmax = m.nq + m.nv, 2 * m.nv, m.nv - 6, 2 * (m.nv - 6) + 2 * (17 + 6)
    o = dict(xnext=np.zeros(nx), xdot=np.zeros(ndx), c=np.zeros(ncmax), ctype=np.zeros(ncmax, np.int32), A=np.zeros((ndx, ndx)),
             B=np.zeros((ndx, nu)), lx=np.zeros(ndx), lu=np.zeros(nu), lxx=np.zeros((ndx, ndx)), lxu=np.zeros((ndx, nu)),
             luu=np.zeros((nu, nu)), Cx=np.zeros((ncmax, ndx)), Cu=np.zeros((ncmax, nu)))
    cost, nc = C.c_double(0), C.c_int(0)
    x, u = np.ascontiguousarray(x), np.ascontiguousarray(u)
    r = orc.orc_stage_eval(C.byref(m), C.byref(ocp), C.byref(slot), P(x), P(u), derivs, C.byref(nc), P(o['xnext']), P(o['xdot']),
                           C.byref(cost), P(o['c']), o['ctype'].ctypes.data_as(C.POINTER(C.c_int)), P(o['A']), P(o['B']), P(o['lx']),
                           P(o['lu']), P(o['lxx']), P(o['lxu']), P(o['luu']), P(o['Cx']), P(o['Cu']))
    assert r == 0, orc.orc_last_error()
    n = nc.value
    o['nc'], o['cost'] = n, cost.value
    o['Cx'] = o['Cx'].ravel()[:n * ndx].reshape(n, ndx)
    o['Cu'] = o['Cu'].ravel()[:n * nu].reshape(n, nu)
    o['c'] = o['c'][:n]
    return o


def rot(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


@pytest.mark.parametrize("mask,land", [(0b11, 0), (0b11, 0b10), (0b01, 0b01), (0b10, 0)])
def test_full6d_stage_derivatives(env, mask, land):
    orc, m, lim, ocpb = env
    ocp = ocpb.desc
    nq, nv = m.nq, m.nv
    ndx, nu = 2 * nv, nv - 6
    rng = np.random.default_rng(60 + mask + 4 * land)
    x = random_state(orc, m, ocp, rng)
    ref = fk_numpy(m, np.array(m.q_ref[:nq]))['feet']
    feet = [(rot(rng.normal(size=3), 0.05) @ ref[f][0], ref[f][1] + rng.normal(size=3) * 0.01) for f in range(2)]
    slot = full_slot(m, ocpb, feet, contact_mask=mask, land_mask=land)
    for f in range(2):
        if mask & (1 << f):
            slot.f_ref[f][3:6] = [1.0, -2.0, 0.5]
    u = rng.normal(size=nu) * 20
    s0 = stage(orc, m, ocp, slot, x, u)
    nstance = bin(mask).count("1")
    assert s0['nc'] == 2 * nu + nstance * 17 + 6 * bin(land & mask).count("1")
    h = 1e-6
    Afd, lxfd, Cxfd = np.zeros((ndx, ndx)), np.zeros(ndx), np.zeros((s0['nc'], ndx))
    for j in range(ndx):
        e = np.zeros(ndx)
        e[j] = h
        sp = stage(orc, m, ocp, slot, integrate(orc, m, ocp, x, e), u, 0)
        sm = stage(orc, m, ocp, slot, integrate(orc, m, ocp, x, -e), u, 0)
        Afd[:, j] = difference(orc, m, ocp, sm['xnext'], sp['xnext']) / (2 * h)
        lxfd[j] = (sp['cost'] - sm['cost']) / (2 * h)
        Cxfd[:, j] = (sp['c'] - sm['c']) / (2 * h)
    Bfd, lufd, Cufd = np.zeros((ndx, nu)), np.zeros(nu), np.zeros((s0['nc'], nu))
    for j in range(nu):
        e = np.zeros(nu)
        e[j] = 1e-3
        sp, sm = stage(orc, m, ocp, slot, x, u + e, 0), stage(orc, m, ocp, slot, x, u - e, 0)
        Bfd[:, j] = difference(orc, m, ocp, sm['xnext'], sp['xnext']) / 2e-3
        lufd[j] = (sp['cost'] - sm['cost']) / 2e-3
        Cufd[:, j] = (sp['c'] - sm['c']) / 2e-3
    assert np.abs(Afd - s0['A']).max() < 2e-7 * max(1.0, np.abs(s0['A']).max())
    assert np.abs(Bfd - s0['B']).max() < 1e-8 * max(1.0, np.abs(s0['B']).max())
    assert np.abs(lxfd - s0['lx']).max() < 2e-6 * max(1.0, np.abs(s0['lx']).max())
    assert np.abs(lufd - s0['lu']).max() < 2e-6 * max(1.0, np.abs(s0['lu']).max())
    assert np.abs(Cxfd - s0['Cx']).max() < 5e-7 * max(1.0, np.abs(s0['Cx']).max())
    assert np.abs(Cufd - s0['Cu']).max() < 1e-7 * max(1.0, np.abs(s0['Cu']).max())
    H = np.block([[s0['lxx'], s0['lxu']], [s0['lxu'].T, s0['luu']]])
    assert np.abs(H - H.T).max() < 1e-9 * np.abs(H).max()
    assert np.linalg.eigvalsh(H).min() > -1e-8 * np.abs(H).max()
