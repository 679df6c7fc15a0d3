"""Device-side MPC driver (simple-mpc_b200/csrc/mpc_driver.cuh) in CPU emulation vs the host `MPC` class.

The checker is the host C++ mirror of the reference driver (simple-mpc_b200/src/mpc.cpp, following reference
src/mpc.cpp:101-324 and src/foot-trajectory.cpp:41-96) bound to the CPU oracle through the test-only hooks module; it
is the code that reproduces the reference's known answers 170/110/219/160 -> 160/100/209/150 (tests/test_host_api.py).
The thing checked is the per-instance device routine (gait queues, FK, Raibert foothold, degree-8 Bezier, references
written into the slot ring), compiled by g++ through par.cuh. Everything is compared per tick: take-off / landing
cycles (exact), contact states (exact), foot references, reference state, base velocity (<= 1e-12).
"""
import ctypes as C
import numpy as np
import pytest

import host_api_utils as hu
import parity_utils as pu
from smpc_ctypes import Slot, Terminal, MpcSettings, dbl_p
from oracle_libs import load_emul

sm = hu.import_simple_mpc()
FEET = ["FL_foot", "FR_foot", "RL_foot", "RR_foot"]


def P(a):
    return a.ctypes.data_as(dbl_p)


def rotated(gait, p):
    return gait[p:] + gait[:p]


class EmulDriver:
    def __init__(self, sc, phase, terminal_constraint=False):
        self.emu = load_emul()
        self.sc = sc
        n = len(sc.GAIT)
        table = (Slot * (n + 1))(*([sc.gait_slot(k) for k in range(n)] + [sc.gait_slot(0)]))
        st = MpcSettings(T_fly=30, T_contact=10, swing_apex=0.15, timestep=0.01, ncycle=n, standing_id=n)
        init = sc.initial_slots()
        term = sc.terminals()
        xref = np.array(sc.xref, float)
        self.h = C.c_void_p(self.emu.emul_mpc_create(C.byref(sc.model), C.byref(sc.ocpb.desc), table, n + 1, C.byref(st), int(phase),
                                                     P(xref), init, term))
        self.T, self.nx = sc.T, sc.nx

    def set_velocity_base(self, v):
        v = np.ascontiguousarray(v, float)
        self.emu.emul_mpc_set_velocity_base(self.h, P(v))

    def set_walking(self, w):
        self.emu.emul_mpc_set_walking(self.h, int(w))

    def tick(self, x):
        slots = (Slot * self.T)()
        term = Terminal()
        x0 = np.zeros(self.nx)
        take, land = (C.c_int * 4)(), (C.c_int * 4)()
        x = np.ascontiguousarray(x, float)
        ov = self.emu.emul_mpc_tick(self.h, P(x), slots, C.byref(term), P(x0), take, land)
        assert ov == 0
        return slots, term, x0, list(take), list(land)

    def close(self):
        self.emu.emul_mpc_destroy(self.h)


def host_mpc(phase, **over):
    hooks = hu.import_hooks()
    mh = hu.go2_handler(sm)
    ocp, _ = hu.go2_kino_problem(sm, mh, T=50, **over)
    # num_threads = 1: the batched device driver this object checks runs the SERIAL Riccati sweep (the leg-parallel sweep
    # is compared with the serial one in tests/test_parallel_riccati.py)
    mpc = hooks.make_oracle_mpc(dict(hu.go2_mpc_conf(mh), num_threads=1), ocp, hu.LIBORC)
    mpc.generateCycleHorizon(rotated(hu.go2_trot(), phase))
    return mh, ocp, mpc


def compare_tick(mpc, ocp, slots, take, land, T=50):
    first = lambda v: v if v is not None else -1  # noqa: E731
    for f, n in enumerate(FEET):
        assert first(mpc.getFootTakeoffCycle(n)) == take[f], (n, mpc.getFootTakeoffCycle(n), take[f])
        assert first(mpc.getFootLandCycle(n)) == land[f], (n, mpc.getFootLandCycle(n), land[f])
    err = 0.0
    for t in range(T):
        cs = ocp.getContactState(t)
        assert [bool((slots[t].contact_mask >> f) & 1) for f in range(4)] == list(cs), t
        for f, n in enumerate(FEET):
            ref = np.array(mpc.getReferencePose(t, n).translation)
            err = max(err, np.abs(ref - np.array(slots[t].foot_ref[f][9:12])).max())
        err = max(err, np.abs(np.array(ocp.getReferenceState(t)) - np.array(slots[t].x_ref[:37])).max())
        err = max(err, np.abs(np.array(ocp.getVelocityBase(t)) - np.array(slots[t].x_ref[19:25])).max())
    return err


@pytest.mark.parametrize("phase", [0, 37])
def test_device_driver_matches_host_mpc(phase):
    sc = pu.Go2Scenario(1, seed=3)
    mh, ocp, mpc = host_mpc(phase)
    drv = EmulDriver(sc, phase)
    vb = np.array([0.2, -0.05, 0, 0, 0, 0.3])
    mpc.switchToWalk(vb)
    drv.set_velocity_base(vb)
    x = np.array(mh.getReferenceState())
    assert np.allclose(x, sc.xref)
    rng = np.random.default_rng(phase)
    worst = 0.0
    for k in range(95):
        if k == 60:  # switchToStand: the horizon keeps cycling until its tail is a full-support stage (reference :222)
            mpc.switchToStand()
            drv.set_walking(0)
        mpc.iterate(x)
        slots, term, x0, take, land = drv.tick(x)
        worst = max(worst, compare_tick(mpc, ocp, slots, take, land))
        assert np.array_equal(x0, x)
        x = np.array(mpc.xs[1])
        x[19:] += rng.normal(size=18) * 0.01
        q = x[3:7]
        x[3:7] = q / np.linalg.norm(q)
    drv.close()
    assert worst < 1e-12, worst


def test_device_driver_timings_at_setup():
    """generateCycleHorizon's tables on the go2 example gait: 60/100/90/129 after 0 ticks;
    the device routine reports the same first entries after its first tick as the host class."""
    sc = pu.Go2Scenario(1, seed=0)
    mh, ocp, mpc = host_mpc(0)
    assert [mpc.getFootTakeoffCycle(n) for n in FEET] == [60, 100, 100, 60]
    drv = EmulDriver(sc, 0)
    x = np.array(mh.getReferenceState())
    mpc.iterate(x)
    _, _, _, take, land = drv.tick(x)
    assert take == [59, 99, 99, 59] and land == [89, 128, 128, 89]
    drv.close()


def test_device_driver_centroidal_talos_matches_host_mpc():
    """CentroidalOCP on the talos-like biped (config C2): the device routine reduces the measured MULTIBODY state to the
    9-D problem state [com; hg] (reference src/centroidal-dynamics.cpp:259-262 via RobotDataHandler::getCentroidalState,
    src/robot-handler.cpp:141-149), writes 6-D contact references (identity rotation + swing translation) and momentum
    references = velocity * mass on the last stage (src/centroidal-dynamics.cpp:237-238). Checked against the host MPC
    class on the oracle, tick by tick."""
    hooks = hu.import_hooks()
    mh = hu.talos_handler(sm)
    ocp, _ = hu.talos_centroidal_problem(sm, mh, T=100)
    mpc = hooks.make_oracle_mpc(hu.talos_mpc_conf(mh), ocp, hu.LIBORC)
    mpc.generateCycleHorizon(hu.talos_walk(20, 80))
    vb = np.array([0.1, 0.02, 0, 0, 0, 0.05])
    mpc.switchToWalk(vb)
    sc = pu.TalosCentroidalScenario(1, seed=4)
    emu = load_emul()
    n = sc.CYCLE
    table = (Slot * (n + 1))(*([sc.gait_slot(k) for k in range(n)] + [sc.gait_slot(0)]))
    st = MpcSettings(T_fly=80, T_contact=20, swing_apex=0.15, timestep=0.01, ncycle=n, standing_id=n)
    xref9 = np.ascontiguousarray(mpc.x_reference, float)  # MPC::x_reference_ = getReferenceState(0): [com ref; velocity ref]
    assert xref9.shape == (9,)
    h = C.c_void_p(emu.emul_mpc_create(C.byref(sc.model), C.byref(sc.ocpb.desc), table, n + 1, C.byref(st), 0, P(xref9),
                                       sc.initial_slots(), sc.terminals()))
    emu.emul_mpc_set_velocity_base(h, P(vb))
    names = ["left_sole_link", "right_sole_link"]
    x = np.array(mh.getReferenceState())
    rng = np.random.default_rng(1)
    worst = 0.0
    mass = mh.getMass()
    for k in range(30):
        x[29 + 6:] = rng.normal(size=22) * 0.05  # joint velocities: non-zero centroidal momentum
        x[29:29 + 6] = rng.normal(size=6) * 0.05
        x[7:29] = np.array(mh.getReferenceState())[7:29] + rng.normal(size=22) * 0.02
        mpc.iterate(x)
        slots = (Slot * 100)()
        term = Terminal()
        x0 = np.zeros(9)
        take, land = (C.c_int * 4)(), (C.c_int * 4)()
        xc = np.ascontiguousarray(x)
        assert emu.emul_mpc_tick(h, P(xc), slots, C.byref(term), P(x0), take, land) == 0
        first = lambda v: v if v is not None else -1  # noqa: E731
        for f, nme in enumerate(names):
            assert first(mpc.getFootTakeoffCycle(nme)) == take[f] and first(mpc.getFootLandCycle(nme)) == land[f]
        worst = max(worst, np.abs(np.array(mpc.xs[0]) - x0).max() / max(1.0, np.abs(x0).max()))  # x0 = [com; hg]
        for t in (0, 17, 50, 99):
            assert [bool((slots[t].contact_mask >> f) & 1) for f in range(2)] == list(ocp.getContactState(t))
            for f, nme in enumerate(names):
                pose = mpc.getReferencePose(t, nme)
                worst = max(worst, np.abs(np.array(pose.translation) - np.array(slots[t].foot_ref[f][9:12])).max())
        # last stage: com reference, momentum references = velocity * mass (getVelocityBase divides by the mass again)
        worst = max(worst, np.abs(np.array(ocp.getPoseBase(99)) - np.array(slots[99].x_ref[:3])).max())
        worst = max(worst, np.abs(np.array(ocp.getVelocityBase(99)) * mass - np.array(slots[99].x_ref[3:9])).max() / mass)
        assert np.allclose(np.array(slots[99].x_ref[3:9]), vb * mass, rtol=1e-14)
    emu.emul_mpc_destroy(h)
    assert worst < 1e-12, worst
