"""Pins the oracle's rigid-contact forward dynamics and the full-dynamics stage (oracle/orc_full.hpp,
full_stage_eval) with the ground truth available here (Pinocchio/Aligator absent, no golden vector):
  * the KKT equations  M a - Jc^T lam = tau - b,  Jc a + gamma = 0  re-checked with an independent numpy
    inverse-dynamics (numerical M, b from finite differences of the centroidal/energy identities is not available,
    so the check is: contact-point acceleration is -Kd v - Kp p, and a without contacts equals the free dynamics),
  * every analytic derivative (da/dq, da/dv, da/dtau, dlam/d.) against central finite differences through the
    manifold retraction,
  * A, B, cost gradients and constraint Jacobians of the full-dynamics stage against finite differences.
"""
import ctypes as C
import numpy as np
import pytest
from smpc_ctypes import go2_like_model, go2_full_ocp, full_slot, fk_numpy, x_ref_of, dbl_p
from oracle_libs import load_oracle


def P(a):
    return a.ctypes.data_as(dbl_p)


@pytest.fixture(scope="module")
def env():
    orc = load_oracle()
    m, lim = go2_like_model()
    ocpb = go2_full_ocp(m, lim, force_cone=True, land_cstr=True, Kp=(3.0, 5.0, 50.0), Kd=(1.0, 2.0, 10.0))
    return orc, m, lim, ocpb


def integrate(orc, m, ocp, x, dx):
    xo = np.zeros(m.nq + m.nv)
    assert orc.orc_state_integrate(C.byref(m), C.byref(ocp), P(x), P(np.ascontiguousarray(dx)), P(xo)) == 0
    return xo


def difference(orc, m, ocp, x0, x1):
    d = np.zeros(2 * m.nv)
    assert orc.orc_state_difference(C.byref(m), C.byref(ocp), P(x0), P(x1), P(d)) == 0
    return d


def random_state(orc, m, ocp, rng):
    x = x_ref_of(m)
    return integrate(orc, m, ocp, x, np.concatenate([rng.normal(size=6) * 0.2, rng.normal(size=12) * 0.2, rng.normal(size=18) * 0.5]))


def fdyn(orc, m, ocp, x, tau, mask, derivs=1, Kp=None, Kd=None):
    nq, nv = m.nq, m.nv
    ncon = 3 * bin(mask).count("1")
    o = dict(a=np.zeros(nv), lam=np.zeros(max(ncon, 1)), da_dq=np.zeros((nv, nv)), da_dv=np.zeros((nv, nv)), da_dtau=np.zeros((nv, nv)),
             dl_dq=np.zeros((max(ncon, 1), nv)), dl_dv=np.zeros((max(ncon, 1), nv)), dl_dtau=np.zeros((max(ncon, 1), nv)))
    g = np.array(ocp.gravity[:])
    Kp = np.array(ocp.Kp_correction[:3]) if Kp is None else np.asarray(Kp, float)
    Kd = np.array(ocp.Kd_correction[:3]) if Kd is None else np.asarray(Kd, float)
    q, v, tau = [np.ascontiguousarray(z) for z in (x[:nq], x[nq:], tau)]
    r = orc.orc_full_dynamics(C.byref(m), P(q), P(v), P(tau), C.c_uint(mask), P(g), P(Kp), P(Kd), derivs,
                              *[P(o[k]) for k in ['a', 'lam', 'da_dq', 'da_dv', 'da_dtau', 'dl_dq', 'dl_dv', 'dl_dtau']], 3)
    assert r == 0, orc.orc_last_error()
    o['lam'] = o['lam'][:ncon]
    for k in ['dl_dq', 'dl_dv', 'dl_dtau']:
        o[k] = o[k][:ncon]
    return o


def foot_point_motion(orc, m, ocp, x, a, f, h=1e-5):
    """position, velocity, acceleration of foot f along the trajectory (q, v, a) by finite differences of FK."""
    nq, nv = m.nq, m.nv

    def pos(t):
        xt = integrate(orc, m, ocp, x, np.concatenate([x[nq:] * t + 0.5 * a * t * t, np.zeros(nv)]))
        return fk_numpy(m, xt[:nq])['feet'][f]
    (R0, p0), (_, pp), (_, pm) = pos(0.0), pos(h), pos(-h)
    return R0, p0, (pp - pm) / (2 * h), (pp - 2 * p0 + pm) / (h * h)


@pytest.mark.parametrize("mask", [0b1111, 0b0110, 0b1001, 0])
def test_contact_dynamics_satisfies_the_constraint(env, mask):
    orc, m, lim, ocpb = env
    ocp = ocpb.desc
    rng = np.random.default_rng(10 + mask)
    x = random_state(orc, m, ocp, rng)
    tau = np.concatenate([np.zeros(6), rng.normal(size=12) * 5])
    o = fdyn(orc, m, ocp, x, tau, mask, derivs=0)
    Kp, Kd = np.array(ocp.Kp_correction[:3]), np.array(ocp.Kd_correction[:3])
    for f in range(4):
        if not mask & (1 << f):
            continue
        # the trajectory q(t) = q (+) (v t + a t^2/2) is exact to second order only for constant body-frame velocity of the
        # base; the O(t^2) error of the retraction is what the tolerance covers
        R, p, pd, pdd = foot_point_motion(orc, m, ocp, x, o['a'], f)
        c = R.T @ pdd + Kd * (R.T @ pd) + Kp * (R.T @ p)
        assert np.abs(c).max() < 5e-4 * max(1.0, np.abs(pdd).max()), c
    if mask == 0:
        assert o['lam'].size == 0


def test_contact_forces_balance_gravity_at_rest(env):
    """standing still with zero torque-free base: sum of world-frame contact forces = m (a_com - g)."""
    orc, m, lim, ocpb = env
    ocp = ocpb.desc
    x = x_ref_of(m)
    tau = np.zeros(18)
    o = fdyn(orc, m, ocp, x, tau, 0b1111, derivs=0, Kp=(0, 0, 0), Kd=(0, 0, 0))
    feet = fk_numpy(m, x[:m.nq])['feet']
    total = sum(feet[f][0] @ o['lam'][3 * f:3 * f + 3] for f in range(4))
    mass = sum(m.mass[i] for i in range(m.njoints))
    # momentum balance: m * a_com = sum f + m g ; with all four feet pinned and zero joint torque the legs fold, so only check
    # the balance through the centroidal momentum rate returned by the kinematics entry point
    nv = m.nv
    out = dict(foot=np.zeros(48), fref=np.zeros(48), base=np.zeros(12), com=np.zeros(3), hg=np.zeros(6), hdg=np.zeros(6),
               Ag=np.zeros(6 * nv), dh_dq=np.zeros(6 * nv), dhd_dq=np.zeros(6 * nv), dhd_dv=np.zeros(6 * nv), Jcom=np.zeros(3 * nv))
    keys = ['foot', 'fref', 'base', 'com', 'hg', 'hdg', 'Ag', 'dh_dq', 'dhd_dq', 'dhd_dv', 'Jcom']
    assert orc.orc_kinematics(C.byref(m), P(x[:m.nq].copy()), P(x[m.nq:].copy()), P(o['a']), *[P(out[k]) for k in keys]) == 0
    assert np.abs(out['hdg'][:3] - (total + mass * np.array(ocp.gravity[:]))).max() < 1e-9 * mass * 9.81
    # angular part about the centre of mass
    ang = sum(np.cross(feet[f][1] - out['com'], feet[f][0] @ o['lam'][3 * f:3 * f + 3]) for f in range(4))
    assert np.abs(out['hdg'][3:] - ang).max() < 1e-9 * mass * 9.81


@pytest.mark.parametrize("mask", [0b1111, 0b0110, 0])
def test_contact_dynamics_derivatives(env, mask):
    orc, m, lim, ocpb = env
    ocp = ocpb.desc
    nq, nv = m.nq, m.nv
    rng = np.random.default_rng(20 + mask)
    x = random_state(orc, m, ocp, rng)
    tau = np.concatenate([rng.normal(size=6), rng.normal(size=12) * 5])
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
        assert np.abs(da - o['da_d' + key]).max() < 2e-7 * sa, key
        assert np.abs(dl - o['dl_d' + key]).max(initial=0.0) < 2e-7 * sl, key
    da, dl = np.zeros((nv, nv)), np.zeros((o['lam'].size, nv))
    for j in range(nv):
        e = np.zeros(nv)
        e[j] = 1e-4
        op, om = fdyn(orc, m, ocp, x, tau + e, mask, 0), fdyn(orc, m, ocp, x, tau - e, mask, 0)
        da[:, j] = (op['a'] - om['a']) / 2e-4
        dl[:, j] = (op['lam'] - om['lam']) / 2e-4
    assert np.abs(da - o['da_dtau']).max() < 1e-8 * max(1.0, np.abs(da).max())
    assert np.abs(dl - o['dl_dtau']).max(initial=0.0) < 1e-8 * max(1.0, np.abs(dl).max(initial=0.0))


def stage(orc, m, ocp, slot, x, u, derivs=1):
    nx, ndx, nu, ncmax = m.nq + m.nv, 2 * m.nv, 12, 24 + 4 * 9
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


@pytest.mark.parametrize("mask,land", [(0b1111, 0), (0b1011, 0b0010), (0b0110, 0b0100)])
def test_full_stage_derivatives(env, mask, land):
    orc, m, lim, ocpb = env
    ocp = ocpb.desc
    nq, nv = m.nq, m.nv
    ndx, nu = 2 * nv, 12
    rng = np.random.default_rng(30 + mask)
    x = random_state(orc, m, ocp, rng)
    feet = [(np.eye(3), fk_numpy(m, np.array(m.q_ref[:nq]))['feet'][f][1] + rng.normal(size=3) * 0.01) for f in range(4)]
    slot = full_slot(m, ocpb, feet, contact_mask=mask, land_mask=land)
    u = rng.normal(size=nu) * 5
    s0 = stage(orc, m, ocp, slot, x, u)
    nstance = bin(mask).count("1")
    assert s0['nc'] == 24 + nstance * 5 + 4 * bin(land & mask).count("1")
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
        e[j] = 1e-4
        sp, sm = stage(orc, m, ocp, slot, x, u + e, 0), stage(orc, m, ocp, slot, x, u - e, 0)
        Bfd[:, j] = difference(orc, m, ocp, sm['xnext'], sp['xnext']) / 2e-4
        lufd[j] = (sp['cost'] - sm['cost']) / 2e-4
        Cufd[:, j] = (sp['c'] - sm['c']) / 2e-4
    assert np.abs(Afd - s0['A']).max() < 1e-7 * max(1.0, np.abs(s0['A']).max())
    assert np.abs(Bfd - s0['B']).max() < 1e-8 * max(1.0, np.abs(s0['B']).max())
    assert np.abs(lxfd - s0['lx']).max() < 1e-6 * max(1.0, np.abs(s0['lx']).max())
    assert np.abs(lufd - s0['lu']).max() < 1e-6 * max(1.0, np.abs(s0['lu']).max())
    assert np.abs(Cxfd - s0['Cx']).max() < 2e-7 * max(1.0, np.abs(s0['Cx']).max())
    assert np.abs(Cufd - s0['Cu']).max() < 1e-7 * max(1.0, np.abs(s0['Cu']).max())
    H = np.block([[s0['lxx'], s0['lxu']], [s0['lxu'].T, s0['luu']]])
    assert np.abs(H - H.T).max() < 1e-9 * np.abs(H).max()
    assert np.linalg.eigvalsh(H).min() > -1e-8 * np.abs(H).max()
