"""ctypes binding of the C ABI (include/smpc.h, libsmpc_cuda.so) plus the synthetic robot / problem builders
used by bench.py, __graft_entry__.smoke() and the tests.

The robots are SYNTHETIC: example-robot-data (Go2 / Talos URDF + SRDF) is not available in this
environment, so `go2_like_model()` is a 12-DoF quadruped with Go2-like link lengths,
masses and the FL,FR,RL,RR joint order implied by reference benchmark/go2.cpp:64-69 and
examples/go2_kinodynamics.py:24-27. Every number reported on it says "go2-like (synthetic)".
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

MAX_JOINTS, MAX_NV, MAX_NQ, MAX_NX, MAX_NDX, MAX_NU, MAX_FEET = 32, 32, 36, 64, 64, 40, 4
KINO, FULL, CENT = 0, 1, 2
dbl_p = C.POINTER(C.c_double)


class ModelDesc(C.Structure):
    _fields_ = [
        ("njoints", C.c_int32), ("nq", C.c_int32), ("nv", C.c_int32),
        ("parent", C.c_int32 * MAX_JOINTS), ("jtype", C.c_int32 * MAX_JOINTS),
        ("idx_q", C.c_int32 * MAX_JOINTS), ("idx_v", C.c_int32 * MAX_JOINTS),
        ("axis", (C.c_double * 3) * MAX_JOINTS),
        ("place_R", (C.c_double * 9) * MAX_JOINTS), ("place_p", (C.c_double * 3) * MAX_JOINTS),
        ("mass", C.c_double * MAX_JOINTS), ("lever", (C.c_double * 3) * MAX_JOINTS),
        ("inertia", (C.c_double * 9) * MAX_JOINTS),
        ("nfeet", C.c_int32), ("foot_joint", C.c_int32 * MAX_FEET),
        ("foot_R", (C.c_double * 9) * MAX_FEET), ("foot_p", (C.c_double * 3) * MAX_FEET),
        ("footref_joint", C.c_int32 * MAX_FEET),
        ("footref_R", (C.c_double * 9) * MAX_FEET), ("footref_p", (C.c_double * 3) * MAX_FEET),
        ("base_joint", C.c_int32), ("base_R", C.c_double * 9), ("base_p", C.c_double * 3),
        ("q_ref", C.c_double * MAX_NQ),
    ]


class OcpDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("horizon", C.c_int32), ("force_size", C.c_int32),
        ("kinematics_limits", C.c_int32), ("torque_limits", C.c_int32), ("force_cone", C.c_int32),
        ("land_cstr", C.c_int32), ("terminal_constraint", C.c_int32),
        ("timestep", C.c_double), ("gravity", C.c_double * 3),
        ("mu", C.c_double), ("Lfoot", C.c_double), ("Wfoot", C.c_double),
        ("w_x", dbl_p), ("w_u", dbl_p), ("w_cent", dbl_p), ("w_centder", dbl_p), ("w_frame", dbl_p),
        ("w_forces", dbl_p), ("w_com", dbl_p), ("w_linear_mom", dbl_p), ("w_angular_mom", dbl_p),
        ("w_linear_acc", dbl_p), ("w_angular_acc", dbl_p),
        ("qmin", dbl_p), ("qmax", dbl_p), ("umin", dbl_p), ("umax", dbl_p),
        ("Kp_correction", C.c_double * 6), ("Kd_correction", C.c_double * 6),
    ]


class SolverOpts(C.Structure):
    _fields_ = [
        ("tol", C.c_double), ("mu_init", C.c_double), ("max_iters", C.c_int32), ("max_al_iters", C.c_int32),
        ("prim_alpha", C.c_double), ("prim_beta", C.c_double), ("dual_alpha", C.c_double), ("dual_beta", C.c_double),
        ("mu_update_factor", C.c_double), ("mu_lower_bound", C.c_double),
        ("inner_tol0", C.c_double), ("prim_tol0", C.c_double),
        ("reg_init", C.c_double), ("reg_min", C.c_double), ("reg_max", C.c_double), ("reg_inc_k", C.c_double),
        ("reg_inc_critical", C.c_double), ("reg_dec_k", C.c_double),
        ("ls_armijo_c1", C.c_double), ("ls_alpha_min", C.c_double), ("ls_contraction", C.c_double),
        ("ls_avg_eta", C.c_double), ("ls_max_steps", C.c_int32), ("num_threads", C.c_int32),
    ]


class Slot(C.Structure):
    _fields_ = [
        ("contact_mask", C.c_uint32), ("land_mask", C.c_uint32),
        ("x_ref", C.c_double * MAX_NX), ("u_ref", C.c_double * MAX_NU),
        ("foot_ref", (C.c_double * 12) * MAX_FEET), ("f_ref", (C.c_double * 6) * MAX_FEET),
        ("land_z", C.c_double * MAX_FEET),
    ]


class Terminal(C.Structure):
    _fields_ = [("x_ref", C.c_double * MAX_NX), ("com_ref", C.c_double * 3), ("tau", C.c_double)]


class MpcSettings(C.Structure):
    """smpc_mpc_settings (include/smpc.h)"""
    _fields_ = [("T_fly", C.c_int32), ("T_contact", C.c_int32), ("swing_apex", C.c_double), ("timestep", C.c_double),
                ("ncycle", C.c_int32), ("standing_id", C.c_int32)]


class SolveInfo(C.Structure):
    _fields_ = [
        ("num_iters", C.c_int32), ("al_iters", C.c_int32), ("conv", C.c_int32), ("ls_trials", C.c_int32),
        ("flags", C.c_int32), ("pad_", C.c_int32),
        ("alpha", C.c_double), ("prim_infeas", C.c_double), ("dual_infeas", C.c_double),
        ("traj_cost", C.c_double), ("merit", C.c_double), ("mu", C.c_double), ("preg", C.c_double),
    ]


def _p(a):
    return a.ctypes.data_as(dbl_p)


def load_cuda():
    # SMPC_LIB: alternative build of the same library (kernel tuning experiments under gpurun)
    path = os.environ.get("SMPC_LIB") or os.path.join(ROOT, "simple-mpc_b200", "libsmpc_cuda.so")
    if not os.path.exists(path):
        raise RuntimeError("simple-mpc_b200/libsmpc_cuda.so not built: run __graft_entry__.build()")
    lib = C.CDLL(path)
    lib.smpc_last_error.restype = C.c_char_p
    lib.smpc_launch_count.restype = C.c_int64
    lib.smpc_riccati_variant.restype = C.c_char_p
    return lib


# ----------------------------------------------------------------------------------------------
# robots
# ----------------------------------------------------------------------------------------------
def _rot(axis, ang):
    axis = np.asarray(axis, float)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def go2_like_model(orc=None):
    """12-DoF quadruped: free-flyer + 4 x (hip-x, thigh-y, calf-y), point feet; order FL, FR, RL, RR."""
    m = ModelDesc()
    I3 = np.eye(3).ravel()
    legs = [("FL", 1, 1), ("FR", 1, -1), ("RL", -1, 1), ("RR", -1, -1)]
    nj = 13
    m.njoints, m.nq, m.nv = nj, 19, 18
    # root
    m.parent[0], m.jtype[0], m.idx_q[0], m.idx_v[0] = -1, 0, 0, 0
    m.place_R[0][:] = I3
    m.mass[0] = 6.921
    m.lever[0][:] = [0.021112, 0.0, -0.005366]
    m.inertia[0][:] = np.diag([0.02448, 0.098077, 0.107]).ravel()
    j = 1
    m.nfeet = 4
    for li, (name, sx, sy) in enumerate(legs):
        hip, thigh, calf = j, j + 1, j + 2
        for jj, (par, ax, pl, mass, lev, inert) in enumerate([
            (0, [1, 0, 0], [sx * 0.1934, sy * 0.0465, 0.0], 0.678, [-sx * 0.0054, sy * 0.00194, -0.000105],
             [0.00048, 0.000884, 0.000596]),
            (hip, [0, 1, 0], [0.0, sy * 0.0955, 0.0], 1.152, [-0.00374, -sy * 0.0223, -0.0327],
             [0.00584, 0.0058, 0.00103]),
            (thigh, [0, 1, 0], [0.0, 0.0, -0.213], 0.194, [0.00548, -0.000975, -0.135],
             [0.00108, 0.0011, 3.29e-05]),
        ]):
            i = j + jj
            m.parent[i], m.jtype[i] = par, 1
            m.idx_q[i], m.idx_v[i] = 7 + (i - 1), 6 + (i - 1)
            m.axis[i][:] = ax
            m.place_R[i][:] = I3
            m.place_p[i][:] = pl
            m.mass[i] = mass
            m.lever[i][:] = lev
            m.inertia[i][:] = np.diag(inert).ravel()
        m.foot_joint[li] = calf
        m.foot_R[li][:] = I3
        m.foot_p[li][:] = [0.0, 0.0, -0.213]
        j += 3
    m.base_joint = 0
    m.base_R[:] = I3
    qref = np.array([0, 0, 0.335, 0, 0, 0, 1] + [0.068, 0.785, -1.44, -0.068, 0.785, -1.44,
                                                  0.068, 0.785, -1.44, -0.068, 0.785, -1.44], float)
    m.q_ref[:19] = qref
    # "<foot>_ref" frames: children of the base frame, placed where the feet are at the reference
    # configuration (reference src/robot-handler.cpp:28-53)
    for f in range(4):
        m.footref_joint[f] = 0
        m.footref_R[f][:] = I3
    fp = fk_numpy(m, qref)
    Mb = fp["base"]
    for f in range(4):
        Mf = fp["feet"][f]
        R = Mb[0].T @ Mf[0]
        p = Mb[0].T @ (Mf[1] - Mb[1])
        m.footref_R[f][:] = R.ravel()
        m.footref_p[f][:] = p
    limits = dict(
        qmin=np.array([-1.0472, -1.5708, -2.7227] * 4),
        qmax=np.array([1.0472, 3.4907, -0.83776] * 4),
    )
    return m, limits


def talos_like_model():
    """22-DoF biped (reduced-Talos sizes: nq 29, nv 28): free-flyer + left leg 6 + right leg 6 + torso 2 +
    left arm 4 + right arm 4, joint order as reference benchmark/talos.cpp:96-107; two QUAD feet (6-D contacts)
    at the soles. SYNTHETIC: Talos-like link lengths and masses, not the real URDF (absent here)."""
    m = ModelDesc()
    I3 = np.eye(3).ravel()
    X, Y, Z = [1, 0, 0], [0, 1, 0], [0, 0, 1]
    joints = [(-1, None, [0, 0, 0], 15.36, [-0.05, 0.0, -0.05], [0.20, 0.12, 0.16])]  # root: pelvis

    def add(par, ax, pl, mass, lev, inert):
        joints.append((par, ax, pl, mass, lev, inert))
        return len(joints) - 1

    soles = []
    for sy in (1, -1):  # left, right
        j = add(0, Z, [-0.02, sy * 0.085, -0.27], 1.85, [0.02, sy * 0.01, 0.03], [0.004, 0.006, 0.005])      # hip yaw
        j = add(j, X, [0, 0, 0], 1.98, [-0.01, sy * 0.005, 0.0], [0.004, 0.005, 0.004])                      # hip roll
        j = add(j, Y, [0, 0, 0], 6.24, [0.003, sy * 0.033, -0.21], [0.15, 0.15, 0.02])                        # hip pitch (thigh)
        j = add(j, Y, [0, 0, -0.38], 3.76, [0.011, sy * 0.003, -0.14], [0.06, 0.06, 0.008])                   # knee (shank)
        j = add(j, Y, [0, 0, -0.325], 1.29, [-0.01, sy * 0.01, 0.0], [0.002, 0.003, 0.002])                   # ankle pitch
        j = add(j, X, [0, 0, 0], 1.60, [-0.004, 0.0, -0.08], [0.004, 0.008, 0.008])                           # ankle roll (foot)
        soles.append(j)
    t1 = add(0, Z, [0, 0, 0.0722], 2.29, [0.0, 0.0, 0.03], [0.006, 0.006, 0.003])                              # torso yaw
    t2 = add(t1, Y, [0, 0, 0], 17.55, [-0.04, 0.0, 0.19], [0.45, 0.35, 0.30])                                  # torso pitch
    for sy in (1, -1):
        j = add(t2, Z, [0, sy * 0.1575, 0.232], 1.43, [-0.007, sy * 0.12, -0.02], [0.005, 0.002, 0.005])      # shoulder 1
        j = add(j, X, [0.006, sy * 0.1829, -0.03], 1.81, [0.02, sy * 0.003, -0.06], [0.008, 0.008, 0.003])    # shoulder 2
        j = add(j, Z, [0, 0, 0], 1.54, [0.007, sy * 0.0003, -0.15], [0.012, 0.012, 0.002])                    # shoulder 3
        j = add(j, Y, [0.02, 0, -0.273], 2.67, [-0.01, sy * 0.002, -0.12], [0.02, 0.02, 0.003])               # elbow (+ forearm, hand)
    nj = len(joints)
    m.njoints, m.nq, m.nv = nj, 7 + nj - 1, 6 + nj - 1
    for i, (par, ax, pl, mass, lev, inert) in enumerate(joints):
        m.parent[i] = par
        m.jtype[i] = 0 if i == 0 else 1
        m.idx_q[i] = 0 if i == 0 else 7 + (i - 1)
        m.idx_v[i] = 0 if i == 0 else 6 + (i - 1)
        if ax is not None:
            m.axis[i][:] = ax
        m.place_R[i][:] = I3
        m.place_p[i][:] = pl
        m.mass[i] = mass
        m.lever[i][:] = lev
        m.inertia[i][:] = np.diag(inert).ravel()
    m.nfeet = 2
    for f in range(2):
        m.foot_joint[f] = soles[f]
        m.foot_R[f][:] = I3
        m.foot_p[f][:] = [0.0, 0.0, -0.107]
    m.base_joint = 0
    m.base_R[:] = I3
    leg = [0.0, 0.0, -0.4, 0.8, -0.4, 0.0]
    arm_l, arm_r = [0.25, 0.17, 0.0, -0.52], [-0.25, -0.17, 0.0, -0.52]
    qj = leg + leg + [0.0, 0.006] + arm_l + arm_r
    qref = np.array([0, 0, 1.0, 0, 0, 0, 1] + qj, float)
    # put the soles on the ground at the reference ("half_sitting") configuration
    m.q_ref[: m.nq] = qref
    qref[2] -= fk_numpy(m, qref)["feet"][0][1][2]
    m.q_ref[: m.nq] = qref
    for f in range(2):
        m.footref_joint[f] = 0
    fp = fk_numpy(m, qref)
    Mb = fp["base"]
    for f in range(2):
        Mf = fp["feet"][f]
        m.footref_R[f][:] = (Mb[0].T @ Mf[0]).ravel()
        m.footref_p[f][:] = Mb[0].T @ (Mf[1] - Mb[1])
    nact = nj - 1
    limits = dict(qmin=-np.ones(nact) * 2.5, qmax=np.ones(nact) * 2.5, umin=-np.ones(nact) * 300.0, umax=np.ones(nact) * 300.0)
    return m, limits


def com_numpy(m, q):
    """centre of mass by plain numpy FK (test-side helper, independent of the C oracle)"""
    fk = fk_numpy(m, q)
    tot, acc = 0.0, np.zeros(3)
    for i in range(m.njoints):
        R, p = fk["joints"][i]
        acc += m.mass[i] * (R @ np.array(m.lever[i][:]) + p)
        tot += m.mass[i]
    return acc / tot


def talos_centroidal_ocp(m, T=100):
    """Problem recipe of reference examples/talos_centroidal.py:49-77."""
    w_u = np.diag(np.array(([0.001] * 3 + [0.1] * 3) * 2))
    arrays = dict(w_u=np.ascontiguousarray(w_u), w_com=np.zeros((3, 3)), w_linear_mom=np.diag([0.01, 0.01, 100.0]),
                  w_angular_mom=np.diag([0.1, 0.1, 1000.0]), w_linear_acc=0.01 * np.eye(3), w_angular_acc=0.01 * np.eye(3))
    d = OcpDesc()
    d.kind, d.horizon, d.force_size = CENT, T, 6
    d.timestep = 0.01
    d.gravity[:] = [0, 0, -9.81]
    d.mu, d.Lfoot, d.Wfoot = 0.8, 0.1, 0.075
    for k, v in arrays.items():
        setattr(d, k, _p(v))
    return OcpBundle(d, arrays)


def centroidal_slot(m, ocp_bundle, foot_poses, x_ref9, contact_mask=None, force_z=None):
    """Slot of a stage created by CentroidalOCP::createStage (reference src/centroidal-dynamics.cpp:39-106)."""
    d = ocp_bundle.desc
    nk, fs = m.nfeet, d.force_size
    s = Slot()
    if contact_mask is None:
        contact_mask = (1 << nk) - 1
    s.contact_mask, s.land_mask = contact_mask, 0
    s.x_ref[:9] = x_ref9
    nact = bin(contact_mask).count("1")
    if force_z is None:
        force_z = -total_mass(m) * d.gravity[2] / max(nact, 1)
    for f in range(nk):
        if contact_mask & (1 << f):
            s.u_ref[fs * f + 2] = force_z
        R, p = foot_poses[f]
        s.foot_ref[f][:9] = np.asarray(R).ravel()
        s.foot_ref[f][9:] = p
        s.land_z[f] = p[2]
    return s


def fk_numpy(m, q):
    """Plain numpy forward kinematics (independent of the C oracle) used only to place the foot reference
    frames and as a known-answer check of the oracle's FK."""
    q = np.asarray(q, float)
    oM = []
    for i in range(m.njoints):
        P = (np.array(m.place_R[i][:]).reshape(3, 3), np.array(m.place_p[i][:]))
        if m.jtype[i] == 0:
            x, y, z, w = q[3:7] / np.linalg.norm(q[3:7])
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            J = (R, q[0:3].copy())
        else:
            J = (_rot(m.axis[i][:], q[m.idx_q[i]]), np.zeros(3))
        li = (P[0] @ J[0], P[0] @ J[1] + P[1])
        if m.parent[i] < 0:
            oM.append(li)
        else:
            pr = oM[m.parent[i]]
            oM.append((pr[0] @ li[0], pr[0] @ li[1] + pr[1]))
    feet = []
    for f in range(m.nfeet):
        pj = oM[m.foot_joint[f]]
        R = np.array(m.foot_R[f][:]).reshape(3, 3)
        p = np.array(m.foot_p[f][:])
        feet.append((pj[0] @ R, pj[0] @ p + pj[1]))
    bj = oM[m.base_joint]
    base = (bj[0] @ np.array(m.base_R[:]).reshape(3, 3), bj[0] @ np.array(m.base_p[:]) + bj[1])
    return dict(joints=oM, feet=feet, base=base)


def total_mass(m):
    return float(sum(m.mass[i] for i in range(m.njoints)))


class OcpBundle:
    """OcpDesc + the numpy arrays its pointers refer to (kept alive together)."""

    def __init__(self, desc, arrays):
        self.desc = desc
        self.arrays = arrays


def go2_kino_ocp(m, limits, T=50, force_cone=False, land_cstr=False, kinematics_limits=True, terminal_constraint=False):
    """Problem recipe of reference examples/go2_kinodynamics.py:42-85."""
    nv = m.nv
    w_x = np.diag(np.array([0, 0, 100, 10, 10, 0] + [1, 1, 1] * 4 + [10] * 6 + [0.1, 0.1, 0.1] * 4, float))
    w_u = np.diag(np.concatenate([np.ones(12) * 0.01, np.ones(nv - 6) * 1e-5]))
    w_cent = np.diag(np.array([0, 0, 1, 0.1, 0.1, 10], float))
    w_centder = np.diag(np.array([0, 0, 0, 0.1, 0.1, 0.1], float))
    w_frame = np.eye(3) * 2000.0
    arrays = dict(w_x=np.ascontiguousarray(w_x), w_u=np.ascontiguousarray(w_u), w_cent=w_cent, w_centder=w_centder,
                  w_frame=w_frame, qmin=limits["qmin"].copy(), qmax=limits["qmax"].copy())
    d = OcpDesc()
    d.kind, d.horizon, d.force_size = KINO, T, 3
    d.kinematics_limits, d.force_cone, d.land_cstr = int(kinematics_limits), int(force_cone), int(land_cstr)
    d.terminal_constraint = int(terminal_constraint)
    d.timestep = 0.01
    d.gravity[:] = [0, 0, -9.81]
    d.mu, d.Lfoot, d.Wfoot = 0.8, 0.01, 0.01
    for k, v in arrays.items():
        setattr(d, k, _p(v))
    return OcpBundle(d, arrays)


def go2_full_ocp(m, limits, T=50, force_cone=False, land_cstr=False, kinematics_limits=True, torque_limits=True,
                 terminal_constraint=False, Kp=(0.0, 0.0, 0.0), Kd=(0.0, 0.0, 0.0), effort=23.7):
    """Problem recipe of reference examples/go2_fulldynamics.py:43-77 (C3 of SURVEY.md §8d flips force_cone / land_cstr on).
    effort: symmetric torque limit (Go2 hip/thigh motors, 23.7 N m; the URDF is absent)."""
    nv = m.nv
    w_x = np.diag(np.array([0, 0, 0, 0, 0, 0] + [1, 1, 1] * 4 + [10] * 6 + [0.1, 0.1, 0.1] * 4, float))
    w_u = np.eye(nv - 6) * 1e-4
    w_cent = np.diag(np.array([0.04, 0.04, 0, 0, 0, 0], float))
    w_forces = np.eye(3) * 1e-4
    w_frame = np.eye(3) * 1000.0
    arrays = dict(w_x=np.ascontiguousarray(w_x), w_u=np.ascontiguousarray(w_u), w_cent=w_cent, w_forces=w_forces,
                  w_frame=w_frame, qmin=limits["qmin"].copy(), qmax=limits["qmax"].copy(),
                  umin=-np.ones(nv - 6) * effort, umax=np.ones(nv - 6) * effort)
    d = OcpDesc()
    d.kind, d.horizon, d.force_size = FULL, T, 3
    d.kinematics_limits, d.torque_limits = int(kinematics_limits), int(torque_limits)
    d.force_cone, d.land_cstr = int(force_cone), int(land_cstr)
    d.terminal_constraint = int(terminal_constraint)
    d.timestep = 0.01
    d.gravity[:] = [0, 0, -9.81]
    d.mu, d.Lfoot, d.Wfoot = 0.8, 0.01, 0.01
    d.Kp_correction[:3] = Kp
    d.Kd_correction[:3] = Kd
    for k, v in arrays.items():
        setattr(d, k, _p(v))
    return OcpBundle(d, arrays)


def talos_full_ocp(m, limits, T=100, force_cone=True, land_cstr=False, kinematics_limits=True, torque_limits=True,
                   terminal_constraint=False, Kp=(0.0, 0.0, 50.0, 0.0, 0.0, 0.0), Kd=(100.0,) * 6):
    """Problem recipe of reference examples/talos_fulldynamics.py:48-95 (C4 of SURVEY.md §8d): 6-D contacts, wrench cones,
    torque + joint limits, Baumgarte Kp = [0, 0, 50, 0, 0, 0], Kd = 100."""
    nv = m.nv
    w_basepos, w_legpos, w_torsopos, w_armpos = [0, 0, 0, 10, 10, 10], [0.1] * 6, [1, 100], [1, 1, 10, 10]
    w_basevel, w_legvel, w_torsovel, w_armvel = [10] * 6, [1] * 6, [1, 100], [10] * 4
    w_x = np.diag(np.array(w_basepos + w_legpos * 2 + w_torsopos + w_armpos * 2 + w_basevel + w_legvel * 2 + w_torsovel + w_armvel * 2, float))
    arrays = dict(w_x=np.ascontiguousarray(w_x), w_u=np.eye(nv - 6) * 1e-4, w_cent=np.diag([0.1, 0.1, 10, 0.1, 0.1, 10.0]),
                  w_forces=np.eye(6) * 0.001, w_frame=np.eye(6) * 2000.0, qmin=limits["qmin"].copy(), qmax=limits["qmax"].copy(),
                  umin=limits["umin"].copy(), umax=limits["umax"].copy())
    d = OcpDesc()
    d.kind, d.horizon, d.force_size = FULL, T, 6
    d.kinematics_limits, d.torque_limits = int(kinematics_limits), int(torque_limits)
    d.force_cone, d.land_cstr = int(force_cone), int(land_cstr)
    d.terminal_constraint = int(terminal_constraint)
    d.timestep = 0.01
    d.gravity[:] = [0, 0, -9.81]
    d.mu, d.Lfoot, d.Wfoot = 0.8, 0.1, 0.075
    d.Kp_correction[:] = Kp
    d.Kd_correction[:] = Kd
    for k, v in arrays.items():
        setattr(d, k, _p(v))
    return OcpBundle(d, arrays)


def talos_kino_ocp(m, limits, T=100, force_cone=True, kinematics_limits=True, terminal_constraint=False):
    """Problem recipe of reference examples/talos_kinodynamics.py:44-112: 6-D contact wrenches as controls,
    CentroidalWrenchCone + 6-D LOCAL frame-velocity equality per stance foot (reference src/kinodynamics.cpp:104-121)."""
    nv = m.nv
    w_basepos, w_legpos, w_torsopos, w_armpos = [0, 0, 1000, 1000, 1000, 1000], [0.1] * 6, [1, 1000], [1, 1, 10, 10]
    w_basevel, w_legvel, w_torsovel, w_armvel = [10] * 6, [1] * 6, [0.1, 100], [10] * 4
    w_x = 10.0 * np.diag(np.array(w_basepos + w_legpos * 2 + w_torsopos + w_armpos * 2 + w_basevel + w_legvel * 2 + w_torsovel + w_armvel * 2, float))
    w_force = [0.001, 0.001, 0.01, 0.1, 0.1, 0.1]
    w_u = np.diag(np.array(w_force * 2 + [1e-4] * (nv - 6), float))
    arrays = dict(w_x=np.ascontiguousarray(w_x), w_u=np.ascontiguousarray(w_u), w_cent=np.diag([0.0, 0.0, 1, 0.1, 0.1, 10.0]),
                  w_centder=np.diag([0.0, 0, 0, 0.1, 0.1, 0.1]), w_frame=np.eye(6) * 100000.0,
                  qmin=limits["qmin"].copy(), qmax=limits["qmax"].copy())
    d = OcpDesc()
    d.kind, d.horizon, d.force_size = KINO, T, 6
    d.kinematics_limits, d.force_cone, d.land_cstr = int(kinematics_limits), int(force_cone), 0
    d.terminal_constraint = int(terminal_constraint)
    d.timestep = 0.01
    d.gravity[:] = [0, 0, -9.81]
    d.mu, d.Lfoot, d.Wfoot = 0.8, 0.1, 0.075
    for k, v in arrays.items():
        setattr(d, k, _p(v))
    return OcpBundle(d, arrays)


def full_slot(m, ocp_bundle, foot_poses, contact_mask=None, land_mask=0):
    """Slot of a stage created by FullDynamicsOCP::createStage (reference src/fulldynamics.cpp:78-214):
    control target 0 (:89), force reference mg/nk on every stance foot (examples/go2_fulldynamics.py:40)."""
    d = ocp_bundle.desc
    nk = m.nfeet
    s = Slot()
    if contact_mask is None:
        contact_mask = (1 << nk) - 1
    s.contact_mask, s.land_mask = contact_mask, land_mask
    xr = x_ref_of(m)
    s.x_ref[: len(xr)] = xr
    for f in range(nk):
        if contact_mask & (1 << f):
            s.f_ref[f][2] = -total_mass(m) * d.gravity[2] / nk
        R, p = foot_poses[f]
        s.foot_ref[f][:9] = np.asarray(R).ravel()
        s.foot_ref[f][9:] = p
        s.land_z[f] = p[2]
    return s


def default_opts(lib, prefix="orc"):
    o = SolverOpts()
    getattr(lib, prefix + "_default_solver_opts")(C.byref(o))
    return o


def x_ref_of(m):
    return np.concatenate([np.array(m.q_ref[: m.nq]), np.zeros(m.nv)])


def standing_slot(m, ocp_bundle, foot_poses, contact_mask=None, land_mask=0, force_z=None):
    """Slot of a stage created by createStage(contact_phase, contact_pose, force_map, land) for the kino handler."""
    d = ocp_bundle.desc
    nk, fs = m.nfeet, d.force_size
    s = Slot()
    if contact_mask is None:
        contact_mask = (1 << nk) - 1
    s.contact_mask, s.land_mask = contact_mask, land_mask
    xr = x_ref_of(m)
    s.x_ref[: len(xr)] = xr
    nact = bin(contact_mask).count("1")
    if force_z is None:
        force_z = -total_mass(m) * d.gravity[2] / max(nact, 1)
    for f in range(nk):
        if contact_mask & (1 << f):
            s.u_ref[fs * f + 2] = force_z
        R, p = foot_poses[f]
        s.foot_ref[f][:9] = np.asarray(R).ravel()
        s.foot_ref[f][9:] = p
        s.land_z[f] = p[2]
    return s
