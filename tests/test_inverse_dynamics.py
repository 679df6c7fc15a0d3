"""KinodynamicsID / CentroidalID (reference src/inverse-dynamics/*.cpp).

CPU part (`-m "not gpu"`): the oracle (oracle/orc_id.hpp) run through the closed-loop scenarios of the reference's own tests
(tests/inverse-dynamics/kinodynamics-id.cpp, centroidal-id.cpp: posture / contact / base / CoM / swing-foot tasks on the quadruped
with point feet and the biped with quad feet, joint / velocity / effort limits checked at every step — shortened horizons), and
the CPU emulation of the CUDA routine (tests/host_emul) against the oracle.
GPU part (`-m gpu`): the CUDA path through the C ABI (include/smpc_id.h) against the oracle on identical measurements along
closed-loop trajectories of a batch, plus the reference's behavioural checks on the CUDA path itself.
Tolerances: both sides solve the same strictly convex QP to a relative KKT residual of 1e-10 with the same algorithm from the
same starting point (cold, or the previous solution of the same closed loop); torques and
contact forces agree to 1e-8 (normwise, with an absolute floor at the scale 1), accelerations to 1e-7: a solve that stops one
Newton step earlier on one side (residual just under / over the threshold) moves them by that much."""
import ctypes as C
import numpy as np
import pytest

import smpc_ctypes as sc
import smpc_id as sid
from smpc_ctypes import _p
from oracle_libs import load_emul, load_oracle

DT = 1e-3
TOL = 1e-8    # torques (the control output) and contact forces
TOL_A = 1e-7  # accelerations: a by-product, more sensitive along weakly regularised directions of the QP (see the docstring)


def robot(name):
    if name == "go2":
        m, lim = sc.go2_like_model()
        return m, lim, 30.0, 23.7
    m, lim = sc.talos_like_model()
    return m, lim, 10.0, 300.0


def make(lib, prefix, name, batch, **kw):
    m, lim, vmax, eff = robot(name)
    d = sid.make_desc(m, lim["qmin"], lim["qmax"], vmax, eff, sid.default_settings(**kw), control_dt=DT)
    return sid.IdBatch(lib, prefix, m, d, batch), m, lim, vmax, eff


def crouched_start(m):
    """the reference tests' `solo_q_start`: knees bent a little more than the reference posture, base shifted"""
    q = np.array(m.q_ref[: m.nq])
    for leg in range(4):
        q[7 + 3 * leg + 1] += 0.1
        q[7 + 3 * leg + 2] -= 0.2
    q[0], q[1] = 0.01, 0.01
    q[2] -= 0.02
    return q


def step_state(q, v, a):
    """the reference tests' integration (TestKinoID::step): q (+)= (v + a dt / 2) dt ; v += a dt"""
    qn = np.stack([sid.integrate_q(q[b], (v[b] + a[b] / 2 * DT) * DT) for b in range(q.shape[0])])
    return qn, v + a * DT


def check_limits(m, lim, vmax, eff, q, v, tau):
    na = m.nv - 6
    assert (q[:, 7:] <= lim["qmax"][None, :na] + 1e-9).all() and (q[:, 7:] >= lim["qmin"][None, :na] - 1e-9).all()
    assert (np.abs(v[:, 6:]) <= vmax + 1e-6).all()
    assert (np.abs(tau) <= eff + 1e-6).all()


def feet_positions(m, q):
    fk = sc.fk_numpy(m, q)
    return np.array([fk["feet"][f][1] for f in range(m.nfeet)])


def closed_loop(ib, m, lim, vmax, eff, q0, steps, free_flyer_pinned=False, on_step=None):
    B = ib.B
    q = np.repeat(q0[None], B, 0) if q0.ndim == 1 else q0.copy()
    v = np.zeros((B, m.nv))
    qref = np.array(m.q_ref[: m.nq])
    for i in range(steps):
        tau = ib.solve(i * DT, q, v)
        a = ib.accelerations()
        assert all(inf.status == 0 for inf in ib.info()), (i, [(inf.status, inf.prim_res, inf.dual_res) for inf in ib.info()])
        qn, vn = step_state(q, v, a)
        if free_flyer_pinned:
            qn[:, :7] = qref[:7]
            vn[:, :6] = 0
        check_limits(m, lim, vmax, eff, qn, vn, tau)
        if on_step:
            on_step(i, q, v, qn, vn, tau, a)
        q, v = qn, vn
    return q, v


# --------------------------------------------------------------------------------------------------------------------
# CPU: the oracle through the reference's scenarios
# --------------------------------------------------------------------------------------------------------------------
def test_oracle_posture_task_error_decreases():
    """reference tests/inverse-dynamics/kinodynamics-id.cpp:109-141 (KinodynamicsID_postureTask)"""
    ib, m, lim, vmax, eff = make(load_oracle(), "orc_id_", "go2", 1, kp_posture=20.0, w_posture=1.0)
    qref = np.array(m.q_ref[: m.nq])
    nf = m.nfeet
    ib.set_target(qref[None], np.zeros((1, m.nv)), np.zeros((1, m.nv)), np.zeros((1, nf), np.int32), np.zeros((1, nf, 6)))
    errs = []
    closed_loop(ib, m, lim, vmax, eff, crouched_start(m), 400, free_flyer_pinned=True,
                on_step=lambda i, q, v, qn, vn, tau, a: errs.append(np.linalg.norm(qn[0, 7:] - qref[7:])))
    assert all(e1 <= e0 + 1e-12 for e0, e1 in zip(errs, errs[1:]))
    assert errs[-1] < 0.75 * errs[0]


@pytest.mark.parametrize("name,equality,kw", [
    ("go2", 0, dict(kp_base=1.0, kp_contact=10.0, w_base=1.0, w_contact_motion=10.0, w_contact_force=1.0)),
    ("go2", 1, dict(kp_base=1.0, kp_contact=10.0, w_base=1.0, w_contact_motion=10.0, w_contact_force=1.0)),
    ("talos", 0, dict(kp_base=1.0, kp_posture=1.0, kp_contact=10.0, w_base=1.0, w_posture=0.05, w_contact_motion=10.0, w_contact_force=1.0)),
    ("talos", 1, dict(kp_base=1.0, kp_posture=1.0, kp_contact=10.0, w_base=1.0, w_posture=0.05, w_contact_motion=10.0, w_contact_force=1.0)),
])
def test_oracle_contacts_hold_still(name, equality, kw):
    """reference kinodynamics-id.cpp:143-238 (contactPoint / contactQuad, cost / equality): contact velocity stays below 1e-2
    m/s while the robot settles; contact forces respect the friction pyramid and carry the weight"""
    ib, m, lim, vmax, eff = make(load_oracle(), "orc_id_", name, 1, contact_motion_equality=equality, **kw)
    q0 = crouched_start(m) if name == "go2" else np.array(m.q_ref[: m.nq])
    mass = sc.total_mass(m)

    def on_step(i, q, v, qn, vn, tau, a):
        foot_vel = (feet_positions(m, qn[0]) - feet_positions(m, q[0])) / DT
        # hard contacts: the reference's 1e-2 m/s; contacts as a cost on the SYNTHETIC biped (whose reference posture is not
        # statically balanced like the real Talos half-sitting one) drift at up to 1.1e-2 m/s while the posture settles
        assert np.linalg.norm(foot_vel, axis=1).max() <= (1e-2 if equality or name == "go2" else 2e-2)
        if i % 20 == 0:
            f = ib.forces()[0]
            if name == "go2":
                fz = f[:, 2]
                assert (np.abs(f[:, :2]) <= 0.6 * fz[:, None] + 1e-6).all() and (fz >= 0.01 * mass * 9.81 - 1e-6).all()
                total = fz.sum()
            else:
                c = f.reshape(m.nfeet, 4, 3)
                assert (np.abs(c[:, :, :2]) <= 0.6 * c[:, :, 2:3] + 1e-6).all()
                total = c[:, :, 2].sum()
            assert abs(total - mass * 9.81) < 0.1 * mass * 9.81

    closed_loop(ib, m, lim, vmax, eff, q0, 120 if name == "talos" else 250, on_step=on_step)


def test_oracle_base_task_converges():
    """reference kinodynamics-id.cpp:240-275 (KinodynamicsID_baseTask), shortened: the base pose error decreases while it is
    above 2e-2 and ends well below its start"""
    ib, m, lim, vmax, eff = make(load_oracle(), "orc_id_", "go2", 1, kp_base=7.0, kp_contact=0.1, w_base=100.0, w_contact_force=1.0,
                                 w_contact_motion=1.0)
    qref = np.array(m.q_ref[: m.nq])
    errs = []
    closed_loop(ib, m, lim, vmax, eff, crouched_start(m), 1500,
                on_step=lambda i, q, v, qn, vn, tau, a: errs.append(np.linalg.norm(qn[0, :3] - qref[:3])))
    errs = np.array(errs)
    above = errs > 2e-2
    assert (np.diff(errs)[above[1:]] <= 1e-12).all()
    assert errs[-1] < 0.25 * errs[0]


def test_oracle_all_tasks_error_decreases():
    """reference kinodynamics-id.cpp:277-309 (KinodynamicsID_allTasks)"""
    ib, m, lim, vmax, eff = make(load_oracle(), "orc_id_", "go2", 1, kp_base=10.0, kp_posture=1.0, kp_contact=10.0, w_base=10.0,
                                 w_posture=0.1, w_contact_force=1.0, w_contact_motion=1.0)
    qref = np.array(m.q_ref[: m.nq])
    errs = []

    def on_step(i, q, v, qn, vn, tau, a):
        errs.append(np.linalg.norm(np.r_[qn[0, :3] - qref[:3], qn[0, 7:] - qref[7:]]))

    closed_loop(ib, m, lim, vmax, eff, crouched_start(m), 400, on_step=on_step)
    assert all(e1 <= e0 + 1e-12 for e0, e1 in zip(errs, errs[1:]))


def centroidal_targets(m, B, lift=None, dz=0.0):
    """CentroidalID::setTarget arguments: CoM of the reference posture (+dz), feet at their current poses; `lift`: foot that
    leaves the ground and tracks a pose 3 cm above its current one"""
    qref = np.array(m.q_ref[: m.nq])
    fk = sc.fk_numpy(m, qref)
    com = sc.com_numpy(m, qref) + np.array([0, 0, dz])
    pose = np.zeros((B, m.nfeet, 12))
    for f in range(m.nfeet):
        R, p = fk["feet"][f]
        pose[:, f, :9] = R.ravel()
        pose[:, f, 9:] = p + (np.array([0, 0, 0.03]) if f == lift else 0)
    cs = np.ones((B, m.nfeet), np.int32)
    if lift is not None:
        cs[:, lift] = 0
    ft = np.zeros((B, m.nfeet, 6))
    ft[:, :, 2] = sc.total_mass(m) * 9.81 / cs[0].sum() * cs
    return np.repeat(com[None], B, 0), np.zeros((B, 3)), pose, np.zeros((B, m.nfeet, 6)), cs, ft


def test_oracle_centroidal_com_task_converges():
    """reference tests/inverse-dynamics/centroidal-id.cpp:290-342 (CentroidalID_comTask), shortened"""
    ib, m, lim, vmax, eff = make(load_oracle(), "orc_id_", "go2", 1, centroidal=1, kp_com=7.0, kp_contact=0.1, w_com=100.0,
                                 w_contact_force=1.0, w_contact_motion=1.0)
    com, vcom, pose, fvel, cs, ft = centroidal_targets(m, 1, dz=-0.03)
    ib.set_target_centroidal(com, vcom, pose, fvel, cs, ft)
    errs = []
    closed_loop(ib, m, lim, vmax, eff, np.array(m.q_ref[: m.nq]), 1500,
                on_step=lambda i, q, v, qn, vn, tau, a: errs.append(np.linalg.norm(sc.com_numpy(m, qn[0]) - com[0])))
    errs = np.array(errs)
    assert (np.diff(errs)[errs[1:] > 1e-3] <= 1e-12).all()
    assert errs[-1] < 0.25 * errs[0]


def test_oracle_centroidal_swing_foot_tracking():
    """reference centroidal-id.cpp:344-409 (CentroidalID_footTrackingTask), shortened: one foot leaves the contact set and tracks
    a pose 3 cm above the ground; the other contacts are hard constraints"""
    ib, m, lim, vmax, eff = make(load_oracle(), "orc_id_", "go2", 1, centroidal=1, kp_feet_tracking=5.0, kp_posture=0.1,
                                 kp_contact=1.0, w_feet_tracking=1e3, w_posture=1.0, w_contact_force=0.1, contact_motion_equality=1)
    com, vcom, pose, fvel, cs, ft = centroidal_targets(m, 1, lift=1)
    ib.set_target_centroidal(com, vcom, pose, fvel, cs, ft)
    errs = []
    # 0.45 s: the reference tilts the base so that the three-legged stance holds for its 5 s; here the run stops before the
    # robot tips over
    closed_loop(ib, m, lim, vmax, eff, np.array(m.q_ref[: m.nq]), 450,
                on_step=lambda i, q, v, qn, vn, tau, a: errs.append(np.linalg.norm(feet_positions(m, qn[0])[1] - pose[0, 1, 9:])))
    errs = np.array(errs)
    assert (np.diff(errs)[errs[1:] > 1e-3] <= 1e-12).all()
    assert errs[-1] < 0.85 * errs[0]
    assert (ib.forces()[0][1] == 0).all()  # the lifted foot carries no force variables


def test_set_target_rejected_on_the_wrong_context_kind():
    orc = load_oracle()
    ib, m, *_ = make(orc, "orc_id_", "go2", 1, centroidal=1, w_com=1.0)
    qref = np.array(m.q_ref[: m.nq])
    with pytest.raises(RuntimeError):
        ib.set_target(qref[None], np.zeros((1, m.nv)), np.zeros((1, m.nv)), np.ones((1, 4), np.int32), np.zeros((1, 4, 6)))
    ib2, *_ = make(orc, "orc_id_", "go2", 1, w_posture=1.0)
    with pytest.raises(RuntimeError):
        ib2.set_target_centroidal(*centroidal_targets(m, 1))


# --------------------------------------------------------------------------------------------------------------------
# CPU: emulation of the CUDA routine against the oracle
# --------------------------------------------------------------------------------------------------------------------
def relerr(a, b, scale=None):
    scale = np.abs(b).max() if scale is None else scale
    return np.abs(a - b).max() / max(scale, 1e-300)


SCENARIOS = [
    ("go2", dict(kp_base=10.0, kp_posture=1.0, kp_contact=10.0, w_base=10.0, w_posture=0.1, w_contact_motion=1.0, w_contact_force=1.0)),
    ("go2", dict(kp_base=1.0, kp_posture=1.0, kp_contact=10.0, w_base=1.0, w_posture=0.1, w_contact_motion=10.0, w_contact_force=1.0,
                 contact_motion_equality=1)),
    ("talos", dict(kp_base=1.0, kp_posture=1.0, kp_contact=10.0, w_base=1.0, w_posture=0.05, w_contact_motion=10.0, w_contact_force=1.0)),
    ("go2", dict(centroidal=1, kp_base=1.0, kp_com=1.0, kp_posture=1.0, kp_contact=10.0, w_base=1.0, w_com=1.0, w_posture=0.1,
                 w_contact_motion=10.0, w_contact_force=1.0)),
]


@pytest.mark.parametrize("name,kw", SCENARIOS)
def test_emulated_cta_routine_matches_oracle(name, kw):
    """csrc/id_qp.cuh compiled for the CPU (tests/host_emul) vs oracle/orc_id.hpp along a short closed-loop trajectory"""
    orc, em = load_oracle(), load_emul()
    ib, m, lim, vmax, eff = make(orc, "orc_id_", name, 1, **kw)
    d = sid.make_desc(m, lim["qmin"], lim["qmax"], vmax, eff, sid.default_settings(**kw), control_dt=DT)
    na = m.nv - 6
    worst = {"tau": 0.0, "a": 0.0}

    ws = np.full(em.emul_id_ws_doubles(C.byref(m), C.byref(d)), np.nan)  # the instance's warm-start record, as in HBM

    def on_step(i, q, v, qn, vn, tau, a):
        te, ae, fe, info = np.zeros(na), np.zeros(m.nv), np.zeros((m.nfeet, 12)), sid.IdInfo()
        rc = em.emul_id_solve(C.byref(m), C.byref(d), None, None, None, None, None, None, None, None, None, 1, _p(q[0].copy()),
                              _p(v[0].copy()), _p(te), _p(ae), _p(fe), C.byref(info), _p(ws))
        assert rc == 0 and info.status == 0
        worst["tau"] = max(worst["tau"], relerr(te, tau[0], max(np.abs(tau[0]).max(), 1.0)))
        worst["a"] = max(worst["a"], relerr(ae, a[0], max(np.abs(a[0]).max(), 1.0)))

    q0 = crouched_start(m) if name == "go2" else np.array(m.q_ref[: m.nq])
    ib.set_warm_start(False)  # the constructor's dry run is not emulated: both sides start their first solve cold
    ib.set_warm_start(True)
    closed_loop(ib, m, lim, vmax, eff, q0, 30 if name == "go2" else 10, on_step=on_step)
    assert worst["tau"] <= TOL and worst["a"] <= TOL_A, worst


# --------------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path through the C ABI
# --------------------------------------------------------------------------------------------------------------------
def perturbed_starts(m, B, seed, name):
    rng = np.random.default_rng(seed)
    q0 = crouched_start(m) if name == "go2" else np.array(m.q_ref[: m.nq])
    q = np.repeat(q0[None], B, 0)
    q[:, 7:] += 0.05 * rng.standard_normal((B, m.nv - 6))
    q[:, :2] += 0.01 * rng.standard_normal((B, 2))
    return q


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", SCENARIOS)
def test_cuda_matches_oracle_along_closed_loop(name, kw):
    """same measurements into libsmpc_cuda (smpc_id_*) and the oracle (orc_id_*) at every step of a closed loop driven by the
    CUDA solution: torques, accelerations and the contact wrenches agree to 1e-8"""
    cuda, orc = sc.load_cuda(), load_oracle()
    B = 16 if name == "go2" else 6
    ig, m, lim, vmax, eff = make(cuda, "smpc_id_", name, B, **kw)
    io, *_ = make(orc, "orc_id_", name, B, **kw)
    q = perturbed_starts(m, B, 1, name)
    v = np.zeros((B, m.nv))
    worst = {"tau": 0.0, "a": 0.0, "w": 0.0}
    steps = 60 if name == "go2" else 20
    for i in range(steps):
        tg, to = ig.solve(i * DT, q, v), io.solve(i * DT, q, v)
        ag, ao = ig.accelerations(), io.accelerations()
        fg, fo = ig.forces(), io.forces()
        assert all(x.status == 0 for x in ig.info()) and all(x.status == 0 for x in io.info())
        for b in range(B):
            worst["tau"] = max(worst["tau"], relerr(tg[b], to[b], max(np.abs(to[b]).max(), 1.0)))
            worst["a"] = max(worst["a"], relerr(ag[b], ao[b], max(np.abs(ao[b]).max(), 1.0)))
        if name == "go2":  # point feet: the force variables are the contact forces (quad corner forces have a null space)
            worst["w"] = max(worst["w"], relerr(fg, fo))
        q, v = step_state(q, v, ag)
        check_limits(m, lim, vmax, eff, q, v, tg)
    assert worst["tau"] <= TOL and worst["a"] <= TOL_A and worst["w"] <= 1e-7, worst
    assert ig.lib.smpc_id_launch_count(ig.ctx) == steps + 1  # one kernel launch per solve (+ the constructor's dry run)


@pytest.mark.gpu
def test_cuda_contacts_hold_still_and_swing_foot_tracks():
    """the reference's behavioural checks on the CUDA path: a batch of quadrupeds settles on hard contacts (contact velocity
    <= 1e-2 m/s, limits respected), then CentroidalID lifts one foot towards a pose 3 cm above the ground"""
    cuda = sc.load_cuda()
    B = 8
    ig, m, lim, vmax, eff = make(cuda, "smpc_id_", "go2", B, kp_base=1.0, kp_contact=10.0, w_base=1.0, w_contact_motion=10.0,
                                 w_contact_force=1.0, contact_motion_equality=1)

    def on_step(i, q, v, qn, vn, tau, a):
        for b in (0, B - 1):
            assert np.linalg.norm((feet_positions(m, qn[b]) - feet_positions(m, q[b])) / DT, axis=1).max() <= 1e-2

    closed_loop(ig, m, lim, vmax, eff, perturbed_starts(m, B, 2, "go2"), 200, on_step=on_step)
    ic, m, lim, vmax, eff = make(cuda, "smpc_id_", "go2", B, centroidal=1, kp_feet_tracking=5.0, kp_posture=0.1, kp_contact=1.0,
                                 w_feet_tracking=1e3, w_posture=1.0, w_contact_force=0.1, contact_motion_equality=1)
    com, vcom, pose, fvel, cs, ft = centroidal_targets(m, B, lift=1)
    ic.set_target_centroidal(com, vcom, pose, fvel, cs, ft)
    errs = []
    closed_loop(ic, m, lim, vmax, eff, np.array(m.q_ref[: m.nq]), 450,
                on_step=lambda i, q, v, qn, vn, tau, a: errs.append(np.linalg.norm(feet_positions(m, qn[0])[1] - pose[0, 1, 9:])))
    assert errs[-1] < 0.85 * errs[0] and (ic.forces()[:, 1] == 0).all()


@pytest.mark.gpu
def test_cuda_rejects_bad_arguments_loudly():
    cuda = sc.load_cuda()
    ig, m, *_ = make(cuda, "smpc_id_", "go2", 2, w_posture=1.0)
    with pytest.raises(RuntimeError):
        ig.set_target_centroidal(*centroidal_targets(m, 2))
    qref = np.array(m.q_ref[: m.nq])
    with pytest.raises(RuntimeError):
        ig.set_target(qref[None], np.zeros((1, m.nv)), np.zeros((1, m.nv)), np.ones((1, 4), np.int32), np.zeros((1, 4, 6)), first=2)


# --------------------------------------------------------------------------------------------------------------------
# host classes / Python module (the reference's bindings/expose-inverse-dynamics.cpp surface)
# --------------------------------------------------------------------------------------------------------------------
def test_python_module_exposes_the_reference_settings_classes():
    import host_api_utils as hu
    sm = hu.import_simple_mpc()
    s = sm.KinodynamicsIDSettings()
    assert (s.friction_coefficient, s.contact_weight_ratio_max, s.contact_weight_ratio_min) == (0.6, 10.0, 0.01)
    assert s.w_base == s.w_posture == s.w_contact_motion == s.w_contact_force == -1.0 and s.contact_motion_equality is False
    c = sm.CentroidalIDSettings()
    assert isinstance(c, sm.KinodynamicsIDSettings) and c.w_com == c.w_feet_tracking == -1.0 and c.kp_com == 0.0
    for cls in (sm.KinodynamicsID, sm.CentroidalID):
        assert all(hasattr(cls, n) for n in ("setTarget", "solve", "getAccelerations"))


@pytest.mark.gpu
def test_host_classes_match_the_c_abi_and_run_the_reference_scenarios():
    """simple_mpc.KinodynamicsID / CentroidalID (C++ host classes over the C ABI) give the torques of the ctypes path on the
    same robot, and the reference's allTasks / posture scenarios behave as its tests demand"""
    import host_api_utils as hu
    sm = hu.import_simple_mpc()
    mh = hu.go2_handler(sm)
    s = sm.KinodynamicsIDSettings()
    s.kp_base, s.kp_posture, s.kp_contact = 10.0, 1.0, 10.0
    s.w_base, s.w_posture, s.w_contact_force, s.w_contact_motion = 10.0, 0.1, 1.0, 1.0
    kid = sm.KinodynamicsID(mh, DT, s)
    xref = np.array(mh.getReferenceState())
    nq, nv = mh.getModel().nq, mh.getModel().nv
    q, v = xref[:nq].copy(), np.zeros(nv)
    for leg in range(4):
        q[7 + 3 * leg + 1] += 0.1
        q[7 + 3 * leg + 2] -= 0.2
    errs = []
    for i in range(150):
        tau = np.array(kid.solve(i * DT, q, v))
        a = np.array(kid.getAccelerations())
        assert tau.shape == (nv - 6,) and a.shape == (nv,) and np.abs(tau).max() <= 23.7 + 1e-6
        q = sid.integrate_q(q, (v + a / 2 * DT) * DT)
        v = v + a * DT
        errs.append(np.linalg.norm(np.r_[q[:3] - xref[:3], q[7:] - xref[7:nq]]))
    assert all(e1 <= e0 + 1e-12 for e0, e1 in zip(errs, errs[1:]))
    # posture only, no contacts (reference KinodynamicsID_postureTask)
    s2 = sm.KinodynamicsIDSettings()
    s2.kp_posture, s2.w_posture = 20.0, 1.0
    kid2 = sm.KinodynamicsID(mh, DT, s2)
    kid2.setTarget(xref[:nq], np.zeros(nv), np.zeros(nv), [False] * 4, [])
    tau = np.array(kid2.solve(0.0, q, v))
    assert np.isfinite(tau).all() and (np.array(kid2.getContactForces()) == 0).all()
    # CentroidalID: lift one foot
    c = sm.CentroidalIDSettings()
    c.kp_feet_tracking, c.kp_posture, c.kp_contact = 5.0, 0.1, 1.0
    c.w_feet_tracking, c.w_posture, c.w_contact_force, c.contact_motion_equality = 1e3, 1.0, 0.1, True
    cid = sm.CentroidalID(mh, DT, c)
    dh = sm.RobotDataHandler(mh)
    dh.updateInternalData(xref, False)
    poses = [dh.getFootPose(i) for i in range(4)]
    target = np.array(poses[0].translation) + np.array([0, 0, 0.03])
    poses[0] = sm.SE3(np.array(poses[0].rotation), target)
    w = mh.getMass() * 9.81 / 3
    m_desc, *_ = robot("go2")
    cid.setTarget(sc.com_numpy(m_desc, xref[:nq]), np.zeros(3), poses, [np.zeros(6)] * 4, [False, True, True, True],
                  [np.zeros(3)] + [np.array([0, 0, w])] * 3)
    q, v = xref[:nq].copy(), np.zeros(nv)
    e0 = None
    for i in range(200):
        cid.solve(i * DT, q, v)
        a = np.array(cid.getAccelerations())
        q = sid.integrate_q(q, (v + a / 2 * DT) * DT)
        v = v + a * DT
        dh.updateInternalData(np.r_[q, v], False)
        e = np.linalg.norm(np.array(dh.getFootPose(0).translation) - target)
        e0 = e if e0 is None else e0
    assert e < 0.97 * e0


def test_warm_start_changes_the_iteration_count_not_the_solution():
    """smpc_id_set_warm_start: a solve that starts from the previous primal / dual solution needs 1-2 Newton steps instead of
    6-8 along a closed loop and returns the torques of the cold-started solve (strictly convex QP)"""
    orc = load_oracle()
    kw = SCENARIOS[0][1]
    warm, m, lim, vmax, eff = make(orc, "orc_id_", "go2", 1, **kw)
    cold, *_ = make(orc, "orc_id_", "go2", 1, **kw)
    cold.set_warm_start(False)
    q, v = crouched_start(m)[None], np.zeros((1, m.nv))
    nw = nc = 0
    for i in range(60):
        tw, tc = warm.solve(i * DT, q, v), cold.solve(i * DT, q, v)
        # two different iteration paths to the same minimiser, each stopped at a relative KKT residual of 1e-10: the torques
        # (the control) agree to 1e-8, the accelerations - more sensitive along weakly regularised directions - to 1e-6
        assert relerr(tw, tc, max(np.abs(tc).max(), 1.0)) <= TOL
        assert relerr(warm.accelerations(), cold.accelerations(), max(np.abs(cold.accelerations()).max(), 1.0)) <= 1e-6
        nw, nc = nw + warm.info()[0].newton_iters, nc + cold.info()[0].newton_iters
        q, v = step_state(q, v, cold.accelerations())
    assert nw < 0.5 * nc, (nw, nc)


@pytest.mark.parametrize("name,lift", [("go2", 2), ("talos", 1)])
def test_emulated_cta_routine_matches_oracle_with_a_swing_foot(name, lift):
    """CentroidalID with one foot out of the contact set (tracking task - 3 rows for a point foot, 6 for a quad foot - and hard
    contacts on the others): emulation of csrc/id_qp.cuh vs the oracle on the first steps of the swing"""
    orc, em = load_oracle(), load_emul()
    kw = dict(centroidal=1, kp_feet_tracking=5.0, kp_posture=0.1, kp_contact=1.0, w_feet_tracking=1e3, w_posture=1.0, w_contact_force=0.1,
              contact_motion_equality=1, kp_com=1.0, w_com=1.0, kp_base=1.0, w_base=1.0)
    if name == "talos":  # single support of the biped: the stance contact as a cost (w = 10). As a hard constraint this QP is
        # infeasible (the tipping biped asks for a wrench outside the stance foot's cone): the primal residual of both solvers
        # stalls at 1e-5 however many multiplier updates are allowed, and they report status 1
        kw.update(contact_motion_equality=0, w_contact_motion=10.0)
    ib, m, lim, vmax, eff = make(orc, "orc_id_", name, 1, **kw)
    d = sid.make_desc(m, lim["qmin"], lim["qmax"], vmax, eff, sid.default_settings(**kw), control_dt=DT)
    com, vcom, pose, fvel, cs, ft = centroidal_targets(m, 1, lift=lift)
    fvel[0, lift, :3] = [0.0, 0.0, 0.05]
    ib.set_target_centroidal(com, vcom, pose, fvel, cs, ft)
    ib.set_warm_start(False)
    ib.set_warm_start(True)
    na = m.nv - 6
    ws = np.full(em.emul_id_ws_doubles(C.byref(m), C.byref(d)), np.nan)
    worst = {"tau": 0.0, "a": 0.0}

    def on_step(i, q, v, qn, vn, tau, a):
        te, ae, fe, info = np.zeros(na), np.zeros(m.nv), np.zeros((m.nfeet, 12)), sid.IdInfo()
        rc = em.emul_id_solve(C.byref(m), C.byref(d), None, None, None, _p(com), _p(vcom), _p(pose), _p(fvel),
                              cs.ctypes.data_as(C.POINTER(C.c_int32)), _p(ft), 0, _p(q[0].copy()), _p(v[0].copy()), _p(te), _p(ae), _p(fe),
                              C.byref(info), _p(ws))
        assert rc == 0 and info.status == 0 and (fe[lift] == 0).all()
        worst["tau"] = max(worst["tau"], relerr(te, tau[0], max(np.abs(tau[0]).max(), 1.0)))
        worst["a"] = max(worst["a"], relerr(ae, a[0], max(np.abs(a[0]).max(), 1.0)))

    closed_loop(ib, m, lim, vmax, eff, np.array(m.q_ref[: m.nq]), 25 if name == "go2" else 10, on_step=on_step)
    assert worst["tau"] <= TOL and worst["a"] <= TOL_A, worst


@pytest.mark.gpu
def test_control_stack_mpc_tick_plus_ten_id_periods():
    """the loop of the reference's examples/go2_kinodynamics.py:196-305 on the CUDA path, batched (smpc_stack.run): MPC tick at
    100 Hz, plan interpolation + KinodynamicsID at 1 kHz, ideal plant. Half a second of walking: every QP solved, robots upright,
    torques within the limits, and the first control periods equal to the same loop on the oracle"""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import smpc_stack
    cuda, orc = sc.load_cuda(), load_oracle()
    B = 4
    _, lim = sc.go2_like_model()
    out = {}
    for name, lib, pre, idpre, ticks in (("cuda", cuda, "smpc", "smpc_id_", 50), ("orc", orc, "orc", "orc_id_", 3)):
        scen = bench.make_scenario("c5", B, seed=1000)
        b = bench.setup_backend(lib, pre, scen, B, 0)
        hist = []

        def on_tick(tick, q, v, tau, kid):
            assert all(x.status == 0 for x in kid.info())
            assert np.isfinite(q).all() and (q[:, 2] > 0.25).all() and (np.abs(tau) <= 23.7 + 1e-6).all()
            hist.append((q.copy(), tau.copy()))

        smpc_stack.run(lib, pre, lib, idpre, b, scen, lim, ticks, on_tick=on_tick)
        out[name] = hist
    for (qc, tc), (qo, to) in zip(out["cuda"][:3], out["orc"]):
        assert relerr(qc, qo) <= 1e-7 and relerr(tc, to, max(np.abs(to).max(), 1.0)) <= 1e-6


def test_infeasible_qp_is_reported_not_hidden():
    """the biped in single support with the stance contact as a HARD constraint: no acceleration / force pair satisfies the
    floating-base dynamics, the contact rows and the stance foot's cone together. The oracle and the emulated CUDA routine both stop
    with status 1 and a primal residual far above the tolerance (the reference asserts HQP_STATUS_OPTIMAL, kinodynamics-id.cpp:230;
    the host class throws)"""
    orc, em = load_oracle(), load_emul()
    kw = dict(centroidal=1, kp_feet_tracking=5.0, kp_posture=0.1, kp_contact=1.0, w_feet_tracking=1e3, w_posture=1.0, w_contact_force=0.1,
              contact_motion_equality=1, kp_com=1.0, w_com=1.0, kp_base=1.0, w_base=1.0)
    ib, m, lim, vmax, eff = make(orc, "orc_id_", "talos", 1, **kw)
    d = sid.make_desc(m, lim["qmin"], lim["qmax"], vmax, eff, sid.default_settings(**kw), control_dt=DT)
    com, vcom, pose, fvel, cs, ft = centroidal_targets(m, 1, lift=1)
    ib.set_target_centroidal(com, vcom, pose, fvel, cs, ft)
    q, v = np.array(m.q_ref[: m.nq]), np.zeros(m.nv)
    ib.solve(0.0, q[None], v[None])
    io = ib.info()[0]
    assert io.status == 1 and io.prim_res > 1e-7
    te, ae, fe, ie = np.zeros(m.nv - 6), np.zeros(m.nv), np.zeros((m.nfeet, 12)), sid.IdInfo()
    rc = em.emul_id_solve(C.byref(m), C.byref(d), None, None, None, _p(com), _p(vcom), _p(pose), _p(fvel),
                          cs.ctypes.data_as(C.POINTER(C.c_int32)), _p(ft), 0, _p(q), _p(v), _p(te), _p(ae), _p(fe), C.byref(ie), None)
    assert rc == 0 and ie.status == 1 and ie.prim_res > 1e-7
