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
from smpc_ctypes import Slot, Terminal, MpcSettings, dbl_p, load_emul

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
    mpc = hooks.make_oracle_mpc(hu.go2_mpc_conf(mh), ocp, hu.LIBORC)
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
