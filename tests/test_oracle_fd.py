"""Pins the CPU oracle (oracle/) with the only ground truth available in this environment
(no golden vector exists in the reference for this path, Aligator/Pinocchio are absent):
  * closed-form Lie-group Jacobians against finite differences and against each other,
  * forward kinematics against an independent numpy implementation,
  * every analytic derivative of the kinodynamics stage (centroidal-dynamics derivatives, A, B, cost gradients,
    constraint Jacobians) against central finite differences taken through the manifold retraction,
  * centroidal identities  h = Ag v,  dh/dv = dhdot/da = Ag,  hdot = d/dt h  along a trajectory.
"""
import ctypes as C
import numpy as np
import pytest
from smpc_ctypes import go2_like_model, go2_kino_ocp, standing_slot, fk_numpy, x_ref_of, dbl_p
from oracle_libs import load_oracle


def P(a):
    return a.ctypes.data_as(dbl_p)


@pytest.fixture(scope="module")
def env():
    orc = load_oracle()
    m, lim = go2_like_model()
    ocpb = go2_kino_ocp(m, lim, force_cone=True, land_cstr=True)
    return orc, m, lim, ocpb


def integrate(orc, m, ocp, x, dx):
    xo = np.zeros(m.nq + m.nv)
    assert orc.orc_state_integrate(C.byref(m), C.byref(ocp), P(x), P(np.ascontiguousarray(dx)), P(xo)) == 0
    return xo


def difference(orc, m, ocp, x0, x1):
    d = np.zeros(2 * m.nv)
    assert orc.orc_state_difference(C.byref(m), C.byref(ocp), P(x0), P(x1), P(d)) == 0
    return d


def test_exp_log_roundtrip_and_jacobians(env):
    orc = env[0]
    rng = np.random.default_rng(0)
    for scale in [0.7, 1e-3, 1e-6]:
        nu = rng.normal(size=6) * scale
        M = np.zeros(12)
        orc.orc_exp6(P(nu), P(M))
        back = np.zeros(6)
        orc.orc_log6(P(M), P(back))
        assert np.abs(back - nu).max() < 1e-14
        J = np.zeros(36)
        Jl = np.zeros(36)
        orc.orc_Jexp6(P(nu), P(J))
        orc.orc_Jlog6(P(nu), P(Jl))
        J, Jl = J.reshape(6, 6), Jl.reshape(6, 6)
        assert np.abs(Jl @ J - np.eye(6)).max() < 1e-13

        def pose(v):
            o = np.zeros(12)
            orc.orc_exp6(P(np.ascontiguousarray(v)), P(o))
            return o[:9].reshape(3, 3), o[9:]

        R0, p0 = pose(nu)
        h = 1e-6
        Jfd = np.zeros((6, 6))
        for i in range(6):
            e = np.zeros(6)
            e[i] = h

            def loc(v):
                R, p = pose(v)
                Mr = np.concatenate([(R0.T @ R).ravel(), R0.T @ (p - p0)])
                o = np.zeros(6)
                orc.orc_log6(P(Mr), P(o))
                return o

            Jfd[:, i] = (loc(nu + e) - loc(nu - e)) / (2 * h)
        assert np.abs(J - Jfd).max() < 1e-8


def kin(orc, m, q, v, a):
    nv = m.nv
    out = dict(foot=np.zeros(48), fref=np.zeros(48), base=np.zeros(12), com=np.zeros(3), hg=np.zeros(6), hdg=np.zeros(6),
               Ag=np.zeros(6 * nv), dh_dq=np.zeros(6 * nv), dhd_dq=np.zeros(6 * nv), dhd_dv=np.zeros(6 * nv), Jcom=np.zeros(3 * nv))
    keys = ['foot', 'fref', 'base', 'com', 'hg', 'hdg', 'Ag', 'dh_dq', 'dhd_dq', 'dhd_dv', 'Jcom']
    q, v, a = [np.ascontiguousarray(z) for z in (q, v, a)]
    assert orc.orc_kinematics(C.byref(m), P(q), P(v), P(a), *[P(out[k]) for k in keys]) == 0
    return out


def random_state(orc, m, ocp, rng):
    x = x_ref_of(m)
    return integrate(orc, m, ocp, x, np.concatenate([rng.normal(size=6) * 0.2, rng.normal(size=12) * 0.2, rng.normal(size=18) * 0.5]))


def test_fk_matches_independent_numpy(env):
    orc, m, lim, ocpb = env
    rng = np.random.default_rng(1)
    x = random_state(orc, m, ocpb.desc, rng)
    k = kin(orc, m, x[:m.nq], x[m.nq:], np.zeros(m.nv))
    ref = fk_numpy(m, x[:m.nq])
    for f in range(4):
        assert np.abs(k['foot'][12 * f + 9:12 * f + 12] - ref['feet'][f][1]).max() < 1e-14
        assert np.abs(k['foot'][12 * f:12 * f + 9].reshape(3, 3) - ref['feet'][f][0]).max() < 1e-14
    # total mass and a hand-computed centre of mass
    masses = np.array([m.mass[i] for i in range(m.njoints)])
    coms = np.array([ref['joints'][i][0] @ np.array(m.lever[i][:]) + ref['joints'][i][1] for i in range(m.njoints)])
    assert np.abs(k['com'] - (masses[:, None] * coms).sum(0) / masses.sum()).max() < 1e-14


def test_centroidal_identities_and_derivatives(env):
    orc, m, lim, ocpb = env
    ocp = ocpb.desc
    nq, nv = m.nq, m.nv
    rng = np.random.default_rng(2)
    x = random_state(orc, m, ocp, rng)
    q, v = x[:nq].copy(), x[nq:].copy()
    a = rng.normal(size=nv)
    k0 = kin(orc, m, q, v, a)
    Ag = k0['Ag'].reshape(6, nv)
    assert np.abs(Ag @ v - k0['hg']).max() < 1e-12
    h = 1e-6

    def fd_q(key):
        cols = []
        for j in range(nv):
            e = np.zeros(2 * nv)
            e[j] = h
            xp, xm = integrate(orc, m, ocp, x, e), integrate(orc, m, ocp, x, -e)
            cols.append((kin(orc, m, xp[:nq], v, a)[key] - kin(orc, m, xm[:nq], v, a)[key]) / (2 * h))
        return np.array(cols).T

    def fd_vec(key, which):
        cols = []
        for j in range(nv):
            e = np.zeros(nv)
            e[j] = h
            if which == 'v':
                cols.append((kin(orc, m, q, v + e, a)[key] - kin(orc, m, q, v - e, a)[key]) / (2 * h))
            else:
                cols.append((kin(orc, m, q, v, a + e)[key] - kin(orc, m, q, v, a - e)[key]) / (2 * h))
        return np.array(cols).T

    assert np.abs(fd_vec('hg', 'v') - Ag).max() < 1e-7
    assert np.abs(fd_vec('hdg', 'a') - Ag).max() < 1e-7
    assert np.abs(fd_q('hg') - k0['dh_dq'].reshape(6, nv)).max() < 1e-7
    assert np.abs(fd_q('hdg') - k0['dhd_dq'].reshape(6, nv)).max() < 2e-7
    assert np.abs(fd_vec('hdg', 'v') - k0['dhd_dv'].reshape(6, nv)).max() < 2e-7
    assert np.abs(fd_q('com') - k0['Jcom'].reshape(3, nv)).max() < 1e-8
    # hdot is the time derivative of h along (q, v, a)
    dt = 1e-6
    xp = integrate(orc, m, ocp, x, np.concatenate([v * dt, a * dt]))
    xm = integrate(orc, m, ocp, x, np.concatenate([-v * dt, -a * dt]))
    hd = (kin(orc, m, xp[:nq], xp[nq:], a)['hg'] - kin(orc, m, xm[:nq], xm[nq:], a)['hg']) / (2 * dt)
    assert np.abs(hd - k0['hdg']).max() < 1e-7


def stage(orc, m, ocp, slot, x, u, derivs=1):
    nx, ndx, nu, ncmax = m.nq + m.nv, 2 * m.nv, 24, 12 + 4 * 6
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
def test_kino_stage_derivatives(env, mask, land):
    orc, m, lim, ocpb = env
    ocp = ocpb.desc
    nq, nv = m.nq, m.nv
    ndx, nu = 2 * nv, 24
    rng = np.random.default_rng(3 + mask)
    x = random_state(orc, m, ocp, rng)
    feet = [(np.eye(3), fk_numpy(m, np.array(m.q_ref[:nq]))['feet'][f][1] + rng.normal(size=3) * 0.01) for f in range(4)]
    slot = standing_slot(m, ocpb, feet, contact_mask=mask, land_mask=land)
    u = np.array(slot.u_ref[:nu]) + rng.normal(size=nu) * np.concatenate([np.ones(12) * 5, np.ones(12) * 2])
    s0 = stage(orc, m, ocp, slot, x, u)
    nstance = bin(mask).count("1")
    assert s0['nc'] == 12 + nstance * (2 + 3) + bin(land & mask).count("1")
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
        e[j] = h
        sp, sm = stage(orc, m, ocp, slot, x, u + e, 0), stage(orc, m, ocp, slot, x, u - e, 0)
        Bfd[:, j] = difference(orc, m, ocp, sm['xnext'], sp['xnext']) / (2 * h)
        lufd[j] = (sp['cost'] - sm['cost']) / (2 * h)
        Cufd[:, j] = (sp['c'] - sm['c']) / (2 * h)
    assert np.abs(Afd - s0['A']).max() < 1e-7
    assert np.abs(Bfd - s0['B']).max() < 1e-8
    assert np.abs(lxfd - s0['lx']).max() < 1e-6 * max(1.0, np.abs(s0['lx']).max())
    assert np.abs(lufd - s0['lu']).max() < 1e-6
    assert np.abs(Cxfd - s0['Cx']).max() < 1e-7
    assert np.abs(Cufd - s0['Cu']).max() < 1e-6 * max(1.0, np.abs(s0['Cu']).max())
    # Gauss-Newton Hessians are symmetric positive semi-definite
    H = np.block([[s0['lxx'], s0['lxu']], [s0['lxu'].T, s0['luu']]])
    assert np.abs(H - H.T).max() < 1e-9 * np.abs(H).max()
    assert np.linalg.eigvalsh(H).min() > -1e-8 * np.abs(H).max()


def test_integrate_difference_inverse(env):
    orc, m, lim, ocpb = env
    rng = np.random.default_rng(5)
    x = random_state(orc, m, ocpb.desc, rng)
    d = rng.normal(size=2 * m.nv) * 0.3
    y = integrate(orc, m, ocpb.desc, x, d)
    assert np.abs(difference(orc, m, ocpb.desc, x, y) - d).max() < 1e-13
