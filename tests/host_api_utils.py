"""Builders shared by the host-API tests: the reference's example configurations
(examples/go2_kinodynamics.py:42-139, examples/talos_centroidal.py:49-117) on the synthetic robots, written
against the Python module `simple_mpc` exactly as a user of the reference would write them."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-mpc_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "host_api"))
LIBORC = os.path.join(ROOT, "oracle", "liborc.so")


def import_simple_mpc():
    import simple_mpc
    return simple_mpc


def import_hooks():
    import simple_mpc  # noqa: F401  (registers the base types)
    import _smpc_hooks
    return _smpc_hooks


def go2_handler(sm):
    model = sm.Model.BuildGo2Like()
    mh = sm.RobotModelHandler(model, "standing", "base")
    for n in ["FL_foot", "FR_foot", "RL_foot", "RR_foot"]:
        mh.addPointFoot(n, "base")
    return mh


def go2_kino_problem(sm, mh, T=50, **over):
    nv = mh.getModel().nv
    nu = nv - 6 + 12
    w_basepos = [0, 0, 100, 10, 10, 0]
    w_legpos = [1, 1, 1]
    w_basevel = [10, 10, 10, 10, 10, 10]
    w_legvel = [0.1, 0.1, 0.1]
    w_x = np.diag(np.array(w_basepos + w_legpos * 4 + w_basevel + w_legvel * 4, float))
    w_u = np.diag(np.concatenate((np.ones(12) * 0.01, np.ones(nv - 6) * 1e-5)))
    conf = dict(
        timestep=0.01, w_x=w_x, w_u=w_u, w_cent=np.diag([0, 0, 1, 0.1, 0.1, 10.0]), w_centder=np.diag([0, 0, 0, 0.1, 0.1, 0.1]),
        w_frame=np.eye(3) * 2000.0, gravity=np.array([0, 0, -9.81]), force_size=3,
        qmin=np.array(mh.getModel().lowerPositionLimit), qmax=np.array(mh.getModel().upperPositionLimit),
        mu=0.8, Lfoot=0.01, Wfoot=0.01, kinematics_limits=True, force_cone=False, land_cstr=False)
    conf.update(over)
    ocp = sm.KinodynamicsOCP(conf, mh)
    ocp.createProblem(mh.getReferenceState(), T, 3, -9.81, False)
    assert nu == ocp.getNu()
    return ocp, conf


def go2_full_problem(sm, mh, T=50, **over):
    """reference examples/go2_fulldynamics.py:43-77 on the synthetic go2-like robot"""
    nv = mh.getModel().nv
    w_x = np.array([0, 0, 0, 0, 0, 0] + [1, 1, 1] * 4 + [10] * 6 + [0.1, 0.1, 0.1] * 4, float)
    conf = dict(
        timestep=0.01, w_x=np.diag(w_x), w_u=np.eye(nv - 6) * 1e-4, w_cent=np.diag([0.04, 0.04, 0, 0, 0, 0.0]), gravity=np.array([0, 0, -9.81]),
        force_size=3, w_forces=np.diag([1e-4, 1e-4, 1e-4]), w_frame=np.eye(3) * 1000.0, umin=-np.ones(nv - 6) * 23.7, umax=np.ones(nv - 6) * 23.7,
        qmin=np.array(mh.getModel().lowerPositionLimit), qmax=np.array(mh.getModel().upperPositionLimit),
        Kp_correction=np.zeros(3), Kd_correction=np.zeros(3), mu=0.8, Lfoot=0.01, Wfoot=0.01, torque_limits=True, kinematics_limits=True,
        force_cone=False, land_cstr=False)
    conf.update(over)
    ocp = sm.FullDynamicsOCP(conf, mh)
    ocp.createProblem(mh.getReferenceState(), T, 3, -9.81, False)
    return ocp, conf


def go2_mpc_conf(mh):
    return dict(support_force=mh.getMass() * 9.81, TOL=1e-4, mu_init=1e-8, max_iters=1, num_threads=8, swing_apex=0.15,
                T_fly=30, T_contact=10, timestep=0.01)


def go2_trot(names=("FL_foot", "FR_foot", "RL_foot", "RR_foot"), T_ds=10, T_ss=30):
    quad = {n: True for n in names}
    # FL + RR lift first, then FR + RL (reference examples/go2_kinodynamics.py:117-139)
    p1 = dict(zip(names, [False, True, True, False]))
    p2 = dict(zip(names, [True, False, False, True]))
    return [quad] * T_ds + [p1] * T_ss + [quad] * T_ds + [p2] * T_ss


def talos_handler(sm):
    model = sm.Model.BuildTalosLike()
    mh = sm.RobotModelHandler(model, "half_sitting", "base_link")
    pts = np.array([[0.1, 0.075, 0], [-0.1, 0.075, 0], [-0.1, -0.075, 0], [0.1, -0.075, 0]])
    mh.addQuadFoot("left_sole_link", "base_link", pts)
    mh.addQuadFoot("right_sole_link", "base_link", pts)
    return mh


def talos_centroidal_problem(sm, mh, T=100):
    w_lin, w_ang = np.ones(3) * 0.001, np.ones(3) * 0.1
    conf = dict(
        timestep=0.01, w_u=np.diag(np.concatenate((w_lin, w_ang, w_lin, w_ang))), w_com=np.diag([0.0, 0, 0]),
        w_linear_mom=np.diag([0.01, 0.01, 100]), w_angular_mom=np.diag([0.1, 0.1, 1000]), w_linear_acc=0.01 * np.eye(3),
        w_angular_acc=0.01 * np.eye(3), gravity=np.array([0, 0, -9.81]), mu=0.8, Lfoot=0.1, Wfoot=0.075, force_size=6)
    ocp = sm.CentroidalOCP(conf, mh)
    dh = sm.RobotDataHandler(mh)
    ocp.createProblem(dh.getCentroidalState(), T, 6, -9.81, False)
    return ocp, conf


def talos_mpc_conf(mh, T_fly=80, T_contact=20):
    return dict(support_force=mh.getMass() * 9.81, TOL=1e-4, mu_init=1e-8, max_iters=1, num_threads=1, swing_apex=0.15,
                T_fly=T_fly, T_contact=T_contact, timestep=0.01)


def talos_walk(T_ds, T_ss):
    dbl = {"left_sole_link": True, "right_sole_link": True}
    left = {"left_sole_link": True, "right_sole_link": False}
    right = {"left_sole_link": False, "right_sole_link": True}
    return [dbl] * T_ds + [left] * T_ss + [dbl] * T_ds + [right] * T_ss
