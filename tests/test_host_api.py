"""Host-side mirror of the reference API (simple-mpc_b200/include/simple-mpc, Python module `simple_mpc`).

CPU tests. The MPC driver is exercised through tests/host_api/_smpc_hooks, which binds the SAME host code to
the CPU oracle's C ABI (the shipped module only ever binds the CUDA library). Known answers:
  * foot timing integers of reference tests/mpc.cpp:78-90 (170/110/219/160, after 10 ticks 160/100/209/150);
  * structural counts and setter/getter round trips of reference tests/problem.cpp.
"""
import numpy as np
import pytest
import host_api_utils as hu

sm = hu.import_simple_mpc()


def test_robot_handler_matches_reference_contract():
    # reference tests/robot_handler.cpp:60-61,79-82,122-152,236-252 on the synthetic quadruped
    mh = hu.go2_handler(sm)
    model = mh.getModel()
    assert (model.nq, model.nv) == (19, 18)
    assert mh.getFeetNb() == 4 and mh.getFeetFrameNames() == ["FL_foot", "FR_foot", "RL_foot", "RR_foot"]
    assert model.existFrame("FL_foot_ref") and mh.getFootNb("RL_foot") == 2
    assert abs(mh.getMass() - (6.921 + 4 * (0.678 + 1.152 + 0.194))) < 1e-12
    dh = sm.RobotDataHandler(mh)
    # "<foot>_ref" frames sit where the feet are at the reference configuration
    for f in range(4):
        assert np.allclose(dh.getFootRefPose(f).translation, dh.getFootPose(f).translation, atol=1e-12)
    x = mh.getReferenceState()
    x2 = x.copy()
    x2[7:19] += 0.1
    x2[19:] = np.arange(18) * 0.01
    d = mh.difference(x, x2)
    assert d.shape == (36,) and np.allclose(d[6:18], 0.1) and np.allclose(d[18:], x2[19:]) and np.allclose(d[:6], 0)
    # centroidal state = [com; hg]: momentum of a pure base translation is m v
    x3 = x.copy()
    x3[19:22] = [0.3, -0.1, 0.05]
    dh.updateInternalData(x3, False)
    xc = dh.getCentroidalState()
    assert np.allclose(xc[3:6], mh.getMass() * x3[19:22], atol=1e-12) and np.allclose(xc[6:9], 0, atol=1e-9)


def test_python_fixtures_equal_cpp_models():
    """simple-mpc_b200/python/smpc_ctypes.py (ctypes binding + synthetic robots) and Model::BuildGo2Like/BuildTalosLike are the
    same robots: identical FK and mass."""
    import smpc_ctypes as S
    for build, handler in [(S.go2_like_model, hu.go2_handler), (S.talos_like_model, hu.talos_handler)]:
        m, _ = build()
        mh = handler(sm)
        assert abs(S.total_mass(m) - mh.getMass()) < 1e-12
        q = np.array(m.q_ref[: m.nq])
        assert np.allclose(q, mh.getReferenceState()[: m.nq], atol=1e-14)
        fk = S.fk_numpy(m, q)
        dh = sm.RobotDataHandler(mh)
        for f in range(m.nfeet):
            assert np.allclose(fk["feet"][f][1], dh.getFootPose(f).translation, atol=1e-13)
        assert np.allclose(S.com_numpy(m, q), dh.getData().com[0], atol=1e-13)


def test_kinodynamics_problem_structure_and_round_trips():
    # reference tests/problem.cpp:106-195
    mh = hu.go2_handler(sm)
    ocp, conf = hu.go2_kino_problem(sm, mh, T=100)
    assert ocp.getSize() == 100
    assert ocp.getCostNumber() == 4 + 4  # state, control, centroidal, centroidal derivative + one pose cost per foot
    assert ocp.getProblem().stages[0].numConstraints() == 1 + 4  # joint box + one velocity constraint per stance foot
    assert ocp.getContactSupport(2) == 4 and ocp.getContactState(2) == [True] * 4
    assert ocp.getNu() == 24
    pose = sm.SE3.Identity()
    pose.translation = np.array([0.1, -0.2, 0.3])
    ocp.setReferencePose(3, "FR_foot", pose)
    assert np.allclose(ocp.getReferencePose(3, "FR_foot").translation, [0.1, -0.2, 0.3])
    f = np.array([1.0, 2.0, 30.0])
    ocp.setReferenceForce(3, "RL_foot", f)
    assert np.allclose(ocp.getReferenceForce(3, "RL_foot"), f)
    assert np.allclose(ocp.getReferenceControl(3)[6:9], f)
    v = np.arange(6) * 0.1
    ocp.setVelocityBase(2, v)
    assert np.allclose(ocp.getVelocityBase(2), v)
    pb = np.array([0, 0, 2, 0, 0, 0, 1.0])
    ocp.setPoseBase(2, pb)
    assert np.allclose(ocp.getPoseBase(2), pb)
    with pytest.raises(RuntimeError, match="Stage index exceeds stage vector size"):
        ocp.setReferencePose(100, "FR_foot", pose)
    with pytest.raises(RuntimeError, match="velocity_base size should be 6"):
        ocp.setVelocityBase(2, np.zeros(5))
    with pytest.raises(KeyError):
        bad = dict(conf)
        del bad["w_cent"]
        sm.KinodynamicsOCP(bad, mh)
    s = ocp.getSettings()
    assert np.array_equal(s["w_x"], conf["w_x"]) and s["force_size"] == 3
    # landing-edge detection of createStages (reference src/ocp-handler.cpp:21-56) with land_cstr on
    ocp2, _ = hu.go2_kino_problem(sm, mh, T=10, land_cstr=True, force_cone=True)
    names = mh.getFeetFrameNames()
    ident = {n: sm.SE3.Identity() for n in names}
    fr = {n: np.array([0, 0, 10.0]) for n in names}
    st = ocp2.createStage({"FL_foot": True, "FR_foot": False, "RL_foot": True, "RR_foot": True}, ident, fr,
                          {"FL_foot": True, "FR_foot": False, "RL_foot": False, "RR_foot": False})
    assert st.contact_mask == 0b1101 and st.land_mask == 0b0001
    assert st.numConstraints() == 1 + 3 * 2 + 1


def test_centroidal_problem_structure_and_round_trips():
    # reference tests/problem.cpp:198-285
    mh = hu.talos_handler(sm)
    ocp, conf = hu.talos_centroidal_problem(sm, mh)
    assert ocp.getSize() == 100 and ocp.getCostNumber() == 6 and ocp.getNu() == 12
    assert ocp.getProblem().stages[0].numConstraints() == 2  # one wrench cone per stance foot
    v = np.array([0.1, 0.2, 0, 0, 0, 0.3])
    ocp.setVelocityBase(5, v)
    assert np.allclose(ocp.getVelocityBase(5), v)
    ocp.setPoseBase(5, np.array([0, 0, 0.9]))
    assert np.allclose(ocp.getReferenceState(5), np.concatenate(([0, 0, 0.9], v)))
    f = np.array([0, 0, 400.0, 1, 2, 3])
    ocp.setReferenceForce(4, "right_sole_link", f)
    assert np.allclose(ocp.getReferenceForce(4, "right_sole_link"), f)


def test_foot_timings_match_reference_known_answers():
    """reference tests/mpc.cpp:46-90: Talos, T = 100, gait 10 DS / 50 (left stance) / 10 DS / 50 (right stance)."""
    hooks = hu.import_hooks()
    mh = hu.talos_handler(sm)
    ocp, _ = hu.talos_centroidal_problem(sm, mh, T=100)
    mpc = hooks.make_oracle_mpc(hu.talos_mpc_conf(mh, T_fly=50, T_contact=10), ocp, hu.LIBORC)
    assert len(mpc.xs) == 101 and len(mpc.us) == 100
    mpc.generateCycleHorizon(hu.talos_walk(10, 50))
    assert mpc.getFootTakeoffCycle("left_sole_link") == 170
    assert mpc.getFootTakeoffCycle("right_sole_link") == 110
    assert mpc.getFootLandCycle("left_sole_link") == 219
    assert mpc.getFootLandCycle("right_sole_link") == 160
    x = mh.getReferenceState()
    for _ in range(10):
        mpc.iterate(x)
    assert mpc.getFootTakeoffCycle("left_sole_link") == 160
    assert mpc.getFootTakeoffCycle("right_sole_link") == 100
    assert mpc.getFootLandCycle("left_sole_link") == 209
    assert mpc.getFootLandCycle("right_sole_link") == 150
    assert mpc.getStateDerivative(0).shape == (9,)
    assert len(mpc.Ks) == 100 and mpc.Ks[0].shape == (12, 9)


def test_go2_kinodynamics_mpc_walks_on_the_oracle():
    """reference examples/go2_kinodynamics.py as a closed loop without a simulator: x <- xs[1]. Checks the driver
    (gait timings 60/100/90/129 of SURVEY.md appendix D, swing references lift the feet by the apex, contact states enter the horizon)."""
    hooks = hu.import_hooks()
    mh = hu.go2_handler(sm)
    ocp, _ = hu.go2_kino_problem(sm, mh, T=50)
    mpc = hooks.make_oracle_mpc(hu.go2_mpc_conf(mh), ocp, hu.LIBORC)
    assert mpc.solver.results.prim_infeas < 1e-3  # the constructor solve (<= 100 iterations) converged
    mpc.generateCycleHorizon(hu.go2_trot())
    assert [mpc.getFootTakeoffCycle(n) for n in mh.getFeetFrameNames()] == [60, 100, 100, 60]
    assert [mpc.getFootLandCycle(n) for n in mh.getFeetFrameNames()] == [90, 129, 129, 90]
    mpc.switchToWalk(np.array([0.2, 0, 0, 0, 0, 0]))
    x = mh.getReferenceState()
    zmax = 0.0
    for k in range(25):
        mpc.iterate(x)
        x = mpc.xs[1]
        zmax = max(zmax, max(mpc.getReferencePose(t, "FL_foot").translation[2] for t in range(50)))
    assert len(mpc.xs) == 51 and len(mpc.us) == 50 and mpc.Ks[0].shape == (24, 36)
    assert mpc.getFootTakeoffCycle("FL_foot") == 35 and mpc.getFootLandCycle("FL_foot") == 65
    # after 25 ticks the first swing of FL/RR (cycle slots 10..39) occupies horizon stages 35..49
    assert ocp.getContactState(34) == [True] * 4 and ocp.getContactState(36) == [False, True, True, False]
    foot_z = sm.RobotDataHandler(mh).getFootPose(0).translation[2]
    assert zmax > foot_z + 0.03  # the swing reference lifts the foot (peak of the degree-8 Bezier: 0.273 x apex)
    assert mpc.getStateDerivative(0).shape == (36,)
    assert np.isfinite(np.array(mpc.us)).all()
    assert np.allclose(ocp.getVelocityBase(49), [0.2, 0, 0, 0, 0, 0]) and np.allclose(ocp.getVelocityBase(20), 0)
    assert mpc.xs[50][19] > 0.02  # the plan accelerates forward towards the commanded base velocity
    with pytest.raises(RuntimeError):
        mpc.getContactForces(0)  # only defined for the full-dynamics handler


def test_contact_state_order_follows_the_feet_registration_order():
    """Feet registered in a non-alphabetical order: getContactState, the device contact mask and the constraint names must
    all describe the same feet (handler order everywhere; the reference's kinodynamics handler mixes std::map order and handler
    order, src/kinodynamics.cpp:45-49 - equal only for alphabetical registration)."""
    model = sm.Model.BuildGo2Like()
    mh = sm.RobotModelHandler(model, "standing", "base")
    order = ["RR_foot", "FL_foot", "RL_foot", "FR_foot"]
    for n in order:
        mh.addPointFoot(n, "base")
    ocp, _ = hu.go2_kino_problem(sm, mh, T=4)
    dh = sm.RobotDataHandler(mh)
    phase = {"RR_foot": False, "FL_foot": True, "RL_foot": True, "FR_foot": False}
    pose = {n: dh.getFootPose(mh.getFootNb(n)) for n in order}
    force = {n: np.array([0, 0, 10.0]) for n in order}
    st = ocp.createStage(phase, pose, force, {n: False for n in order})
    assert list(st.contact_states) == [phase[n] for n in order] == [False, True, True, False]
    assert st.contact_mask == sum(1 << i for i, n in enumerate(order) if phase[n])
    assert [c for c in st.constraint_names if c.endswith("_frame_velocity")] == ["FL_foot_frame_velocity", "RL_foot_frame_velocity"]


def test_python_derived_handlers_are_driven_through_the_trampoline():
    """reference include/simple-mpc/python/py-ocp-handler.hpp:71-192 (PyOCPHandler): handlers derived in Python. The C++ driver
    (createProblem, MPC::generateCycleHorizon, MPC::iterate -> updateStepTrackerReferences) must reach the Python overrides
    through the virtual calls, and super() must reach the C++ implementation without recursion."""
    hooks = hu.import_hooks()
    mh = hu.go2_handler(sm)
    calls = dict(createStage=0, setReferencePose=0, setReferenceState=0)

    class MyKino(sm.KinodynamicsOCP):
        def createStage(self, contact_map, contact_pose, force_refs, land_constraint):
            calls["createStage"] += 1
            return super().createStage(contact_map, contact_pose, force_refs, land_constraint)

        def setReferencePose(self, t, ee_name, pose_ref):
            calls["setReferencePose"] += 1
            super().setReferencePose(t, ee_name, pose_ref)

        def setReferenceState(self, t, x_ref):
            calls["setReferenceState"] += 1
            super().setReferenceState(t, x_ref)

    _, conf = hu.go2_kino_problem(sm, mh, T=50)
    ocp = MyKino(conf, mh)
    ocp.createProblem(mh.getReferenceState(), 50, 3, -9.81, False)
    assert calls["createStage"] == 50  # one standing stage per knot (reference src/ocp-handler.cpp:96-137)
    mpc = hooks.make_oracle_mpc(hu.go2_mpc_conf(mh), ocp, hu.LIBORC)
    mpc.generateCycleHorizon(hu.go2_trot())
    assert calls["createStage"] >= 50 + 80  # the cycle horizon is built through the same virtual
    mpc.switchToWalk(np.array([0.2, 0, 0, 0, 0, 0]))
    n0 = calls["setReferencePose"]
    mpc.iterate(mh.getReferenceState())
    assert calls["setReferencePose"] > n0 and calls["setReferenceState"] > 0  # updateStepTrackerReferences, reference src/mpc.cpp:278-324
    assert len(mpc.xs) == 51 and np.isfinite(np.array(mpc.us)).all()

    # a handler derived from the abstract OCPHandler: implemented methods are used, missing ones are "pure virtual" errors
    class Bare(sm.OCPHandler):
        def getReferenceState(self, t):
            return np.arange(3.0) + t

    bare = Bare(mh)
    assert np.allclose(bare.getReferenceState(2), [2, 3, 4])
    with pytest.raises(RuntimeError, match="pure virtual"):
        bare.createTerminalCost()


def test_fulldynamics_handler_lowers_but_has_no_cpu_path():
    """no CPU fallback: on a box without a CUDA device the product `MPC` (always bound to libsmpc_cuda) refuses to
    construct - smpc_create fails with SMPC_ERR_CUDA -> RuntimeError - instead of solving on the CPU. (The driver
    runs the `not gpu` suite on the GPU-less build box; the CUDA path of this handler is covered by
    tests/test_gpu_parity.py.)"""
    import ctypes
    try:
        ctypes.CDLL("libcuda.so.1")
        pytest.skip("a CUDA driver is present: the no-fallback refusal can only be observed on a GPU-less box")
    except OSError:
        pass
    mh = hu.go2_handler(sm)
    nv = mh.getModel().nv
    conf = dict(timestep=0.01, w_x=np.eye(36), w_u=np.eye(12) * 1e-4, w_cent=np.eye(6), w_forces=np.eye(3) * 1e-3, w_frame=np.eye(3) * 1000,
                gravity=np.array([0, 0, -9.81]), force_size=3, mu=0.8, Lfoot=0.01, Wfoot=0.01, umin=-np.ones(12) * 23, umax=np.ones(12) * 23,
                qmin=np.array(mh.getModel().lowerPositionLimit), qmax=np.array(mh.getModel().upperPositionLimit),
                Kp_correction=np.zeros(3), Kd_correction=np.zeros(3), torque_limits=True, kinematics_limits=True, force_cone=True, land_cstr=True)
    ocp = sm.FullDynamicsOCP(conf, mh)
    ocp.createProblem(mh.getReferenceState(), 20, 3, -9.81, False)
    assert ocp.getNu() == nv - 6 and ocp.getCostNumber() == 3 + 4 + 4
    assert ocp.getProblem().stages[0].numConstraints() == 2 + 4  # torque box, joint box, one friction cone per stance foot
    with pytest.raises(RuntimeError):
        sm.MPC(hu.go2_mpc_conf(mh), ocp)


def test_interpolator_matches_reference_semantics_and_the_oracle_plan_interpolation():
    """Interpolator (reference src/interpolator.cpp:5-78): values at the knots, clamping beyond the last knot, linear
    parts, contact lookup; the state interpolation agrees with the oracle's restatement used for the batched kernel."""
    import ctypes as C
    import parity_utils as pu
    from smpc_ctypes import SolverOpts
    from oracle_libs import load_oracle
    mh = hu.go2_handler(sm)
    interp = sm.Interpolator(mh.getModel())
    orc = load_oracle()
    sc = pu.Go2Scenario(1, seed=21)
    opts = SolverOpts()
    orc.orc_default_solver_opts(C.byref(opts))
    b = pu.Backend(orc, "orc", sc.model, sc.ocpb, opts, 1)
    b.set_slots(sc.initial_slots())
    b.set_terminal(sc.terminals())
    b.set_warmstart(*sc.warmstart())
    b.solve(sc.x_start(), 5)  # a plan with a moving base
    xs, us, _, _ = b.results(gains=False)
    xl, ul = [xs[0, t] for t in range(51)], [us[0, t] for t in range(50)]
    dt = 0.01
    for delay in (0.0, 0.0042, 0.1234, 0.4899, 0.495, 0.5, 0.9):
        xo, uo = np.zeros((1, 37)), np.zeros((1, 24))
        assert orc.orc_interpolate_plan(b.ctx, 0, 1, C.c_double(delay), pu.P(xo), pu.P(uo), None) == 0
        assert np.allclose(interp.interpolateState(delay, dt, xl), xo[0], rtol=0, atol=1e-12)
        assert np.allclose(interp.interpolateLinear(delay, dt, ul), uo[0], rtol=0, atol=1e-12 * np.abs(us).max())
        assert np.allclose(interp.interpolateConfiguration(delay, dt, [x[:19] for x in xl]), xo[0, :19], rtol=0, atol=1e-12)
    assert np.array_equal(interp.interpolateState(3 * dt, dt, xl)[7:], xl[3][7:])
    cs = [[True, False, True, False]] * 3 + [[False, True, False, True]] * 2
    assert interp.interpolateContacts(0.025, dt, cs) == [True, False, True, False]
    assert interp.interpolateContacts(0.031, dt, cs) == [False, True, False, True]
    assert interp.interpolateContacts(7.0, dt, cs) == [False, True, False, True]
    with pytest.raises(RuntimeError):
        interp.interpolateState(0.0, dt, [np.zeros(5)])
    b.close()


def test_go2_fulldynamics_mpc_walks_on_the_oracle():
    """reference examples/go2_fulldynamics.py through the host layer: FullDynamicsOCP lowered to the C ABI descriptors and
    driven by the same `MPC` driver on the CPU oracle; contact forces carry the weight, torques stay inside the limits."""
    hooks = hu.import_hooks()
    mh = hu.go2_handler(sm)
    ocp, conf = hu.go2_full_problem(sm, mh, T=50)
    mpc = hooks.make_oracle_mpc(hu.go2_mpc_conf(mh), ocp, hu.LIBORC)
    mpc.generateCycleHorizon(hu.go2_trot())
    mpc.switchToWalk(np.array([0.2, 0, 0, 0, 0, 0]))
    x = mh.getReferenceState()
    for k in range(12):
        mpc.iterate(x)
        x = mpc.xs[1]
    assert len(mpc.us) == 50 and mpc.us[0].shape == (12,) and mpc.Ks[0].shape == (12, 36)
    us = np.array(mpc.us)
    assert np.isfinite(us).all() and np.abs(us).max() <= 23.7 * (1 + 1e-3)
    assert mpc.getStateDerivative(0).shape == (36,)
    # getContactForces (reference src/mpc.cpp:354-380): the stance feet carry the weight, swing rows are zero
    f0 = np.array(mpc.getContactForces(0))
    assert f0.shape == (4, 3) and abs(f0[:, 2].sum() - mh.getMass() * 9.81) < 0.25 * mh.getMass() * 9.81
    f49 = np.array(mpc.getContactForces(49))
    assert np.all(f49[0] == 0) and np.all(f49[3] == 0) and f49[1, 2] > 0 and f49[2, 2] > 0
    with pytest.raises(RuntimeError):
        mpc.getContactForces(50)
    # after 12 ticks the horizon tail holds cycle slots 10 and 11: the first FL + RR swing has entered the horizon
    assert ocp.getContactState(49) == [False, True, True, False] and ocp.getContactState(47) == [True] * 4


def test_friction_compensation_matches_the_reference_test():
    # reference tests/friction.cpp:10-43: dry 0.5 / viscous 0.05 on every actuated joint, random velocity and torque
    mh = hu.go2_handler(sm)
    model = mh.getModel()
    nu = model.nv - 6
    fr, dm = np.zeros(model.nv), np.zeros(model.nv)
    fr[-nu:], dm[-nu:] = 0.5, 0.05
    model.friction, model.damping = fr, dm
    fc = sm.FrictionCompensation(model, True)
    assert np.array_equal(fc.dry_friction, fr[-nu:]) and np.array_equal(fc.viscuous_friction, dm[-nu:])
    rng = np.random.default_rng(0)
    v, tau = rng.uniform(-1, 1, nu), rng.uniform(-1, 1, nu)
    v[3] = 0.0
    expect = tau + fr[6:] * np.sign(v) + dm[6:] * v
    assert np.allclose(fc.computeFriction(v, tau), expect, rtol=0, atol=1e-15)
    with pytest.raises(RuntimeError):
        fc.computeFriction(v[:-1], tau)
    # without a free flyer every dof is actuated; unset coefficients mean no friction
    assert sm.FrictionCompensation(sm.Model(), False).nu == 0
