"""The control stack of the reference's examples (examples/go2_kinodynamics.py:196-305), batched: every 10 ms one MPC tick
(smpc_mpc_iterate), then N_simu = 10 control periods of 1 ms in which the plan is interpolated (smpc_interpolate_plan), handed to
KinodynamicsID as its target (smpc_id_set_target) and turned into joint torques (smpc_id_solve). The simulator of the examples
(PyBullet) is not part of the path: the plant here is the ideal one - it realises the accelerations of the whole-body QP - which is
what the reference's own inverse-dynamics tests integrate (tests/inverse-dynamics/kinodynamics-id.cpp:46-60)."""
import ctypes as C
import time
import numpy as np

import smpc_id as sid
from smpc_scenarios import P

N_SIMU = 10


def example_id_settings():
    """examples/go2_kinodynamics.py:147-156"""
    return sid.default_settings(kp_base=7.0, kp_posture=10.0, kp_contact=10.0, w_base=100.0, w_posture=1.0, w_contact_force=1.0,
                                w_contact_motion=1.0)


def run(lib, prefix, id_lib, id_prefix, backend, sc, limits, ticks, vmax=30.1, effort=23.7, dt_mpc=0.01, on_tick=None, zero_base_acc_target=True):
    """backend: a set-up MPC context (bench.setup_backend); sc: its scenario. Returns the final states, the torques of the last
    control period and per-phase wall-clock times (ms per tick).
    zero_base_acc_target: KinodynamicsID::solve hands the base ACCELERATION target to the base task as its VELOCITY reference
    (reference src/inverse-dynamics/kinodynamics-id.cpp:224-225 calls setDerivative twice; kept, see oracle/orc_id.hpp). On the
    ideal plant used here that term (gain 2 sqrt(kp_base)) feeds the MPC's base acceleration back with the wrong sign of
    damping and the MPC -> ID -> plant loop diverges within 0.25 s (measured, tools/control_stack.py --raw-base-acc); the
    default therefore zeroes the six base entries of a_target. The example feeds them through to a physics simulator."""
    B, m = backend.B, sc.model
    nq, nv, nk = m.nq, m.nv, m.nfeet
    dt = dt_mpc / N_SIMU
    d = sid.make_desc(m, limits["qmin"], limits["qmax"], vmax, effort, example_id_settings(), control_dt=dt)
    kid = sid.IdBatch(id_lib, id_prefix, m, d, B)
    xm = np.ascontiguousarray(sc.xmeas_start())
    q, v = xm[:, :nq].copy(), xm[:, nq:].copy()
    f = lambda name: getattr(lib, prefix + "_" + name)
    x_i, u_i, cs = np.zeros((B, nq + nv)), np.zeros((B, backend.nu)), np.zeros((B, nk), np.int32)
    xd0, xd1 = np.zeros((B, 2 * nv)), np.zeros((B, 2 * nv))
    t_mpc = t_id = 0.0
    tau = None
    for tick in range(ticks):
        t0 = time.perf_counter()
        xm = np.ascontiguousarray(np.concatenate([q, v], axis=1))
        assert f("mpc_iterate")(backend.ctx, P(xm), 1) == 0, f("last_error")()
        assert f("get_xdot")(backend.ctx, 0, B, 0, P(xd0)) == 0 and f("get_xdot")(backend.ctx, 0, B, 1, P(xd1)) == 0
        t1 = time.perf_counter()
        for sub in range(N_SIMU):
            s = sub / float(N_SIMU)
            assert f("interpolate_plan")(backend.ctx, 0, B, C.c_double(s * dt_mpc), P(x_i), P(u_i), cs.ctypes.data_as(C.POINTER(C.c_int))) == 0
            acc = (1 - s) * xd0[:, nv:] + s * xd1[:, nv:]
            acc[:, 6:] = u_i[:, 3 * nk:]          # kinodynamics: the joint accelerations are controls (example :257-258)
            if zero_base_acc_target:
                acc[:, :6] = 0.0
            forces = np.zeros((B, nk, 6))
            forces[:, :, :3] = u_i[:, : 3 * nk].reshape(B, nk, 3)
            kid.set_target(x_i[:, :nq], x_i[:, nq:], acc, cs, forces)
            tau = kid.solve(tick * dt_mpc + sub * dt, q, v)
            a = kid.accelerations()
            q = sid.integrate_q_batch(q, (v + a / 2 * dt) * dt)
            v = v + a * dt
        t2 = time.perf_counter()
        t_mpc += t1 - t0
        t_id += t2 - t1
        if on_tick:
            on_tick(tick, q, v, tau, kid)
    return q, v, tau, {"mpc_ms_per_tick": 1e3 * t_mpc / ticks, "id_ms_per_tick": 1e3 * t_id / ticks, "kid": kid}
