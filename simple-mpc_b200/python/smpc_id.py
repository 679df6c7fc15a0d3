"""ctypes binding of include/smpc_id.h: the batched whole-body inverse-dynamics QP (the reference's
KinodynamicsID / CentroidalID, src/inverse-dynamics/*.cpp). `IdBatch` drives `batch` independent controllers through
the C ABI of whichever library it is given: libsmpc_cuda.so (prefix smpc_id_, the product) or - in tests and the
CPU legs of bench.py only - the oracle (prefix orc_id_)."""
import ctypes as C
import numpy as np

from smpc_ctypes import MAX_FEET, MAX_NV, ModelDesc, _p

MAX_FVAR = 12
POINT, QUAD = 0, 1


class IdSettings(C.Structure):
    _fields_ = [
        ("friction_coefficient", C.c_double), ("contact_weight_ratio_max", C.c_double), ("contact_weight_ratio_min", C.c_double),
        ("kp_base", C.c_double), ("kp_posture", C.c_double), ("kp_contact", C.c_double),
        ("w_base", C.c_double), ("w_posture", C.c_double), ("w_contact_motion", C.c_double), ("w_contact_force", C.c_double),
        ("contact_motion_equality", C.c_int32), ("centroidal", C.c_int32),
        ("kp_com", C.c_double), ("kp_feet_tracking", C.c_double), ("w_com", C.c_double), ("w_feet_tracking", C.c_double),
    ]


class IdDesc(C.Structure):
    _fields_ = [
        ("control_dt", C.c_double),
        ("foot_type", C.c_int32 * MAX_FEET),
        ("quad_points", ((C.c_double * 3) * 4) * MAX_FEET),
        ("qmin", C.c_double * MAX_NV), ("qmax", C.c_double * MAX_NV), ("vmax", C.c_double * MAX_NV), ("effort", C.c_double * MAX_NV),
        ("settings", IdSettings),
    ]


class IdInfo(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("outer_iters", C.c_int32), ("newton_iters", C.c_int32),
        ("nvar", C.c_int32), ("neq", C.c_int32), ("nin", C.c_int32),
        ("prim_res", C.c_double), ("dual_res", C.c_double),
    ]


def default_settings(**kw):
    """KinodynamicsID::Settings / CentroidalID::Settings defaults (kinodynamics-id.hpp:22-47, centroidal-id.hpp:17-26)."""
    s = IdSettings()
    s.friction_coefficient, s.contact_weight_ratio_max, s.contact_weight_ratio_min = 0.6, 10.0, 0.01
    s.w_base = s.w_posture = s.w_contact_motion = s.w_contact_force = s.w_com = s.w_feet_tracking = -1.0
    for k, v in kw.items():
        if not hasattr(s, k):
            raise KeyError(k)
        setattr(s, k, v)
    return s


def make_desc(model, qmin, qmax, vmax, effort, settings, control_dt=1e-3, quad_half=(0.1, 0.05)):
    """quad feet (Talos-like): a 2 Lx x 2 Ly rectangle in the sole frame, corner order of the reference's
    addQuadFoot call in tests/test_utils.cpp (the four sign combinations)."""
    d = IdDesc()
    d.control_dt = control_dt
    na = model.nv - 6
    for name, arr in (("qmin", qmin), ("qmax", qmax), ("vmax", vmax), ("effort", effort)):
        a = np.broadcast_to(np.asarray(arr, float), (na,))
        getattr(d, name)[:na] = a
    quad = model.nfeet == 2
    lx, ly = quad_half
    corners = [(lx, ly, 0.0), (lx, -ly, 0.0), (-lx, -ly, 0.0), (-lx, ly, 0.0)]
    for f in range(model.nfeet):
        d.foot_type[f] = QUAD if quad else POINT
        for k in range(4):
            d.quad_points[f][k][:] = corners[k]
    d.settings = settings
    return d


class IdBatch:
    def __init__(self, lib, prefix, model, desc, batch):
        self.lib, self.pre, self.model, self.B = lib, prefix, model, batch
        self.nq, self.nv, self.nfeet = model.nq, model.nv, model.nfeet
        self.na = self.nv - 6
        err = getattr(lib, prefix + "last_error", None) or getattr(lib, "smpc_last_error")
        err.restype = C.c_char_p
        self._err = err
        self.ctx = C.c_void_p()
        self._check(self._f("create")(C.byref(self.ctx), C.byref(model), C.byref(desc), C.c_int(batch)))

    def _f(self, name):
        return getattr(self.lib, self.pre + name)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("%s: rc %d: %s" % (self.pre, rc, (self._err() or b"").decode()))

    def close(self):
        if self.ctx:
            fn = self._f("destroy")
            fn.restype = None
            fn(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_target(self, q, v, a, contact, f_target, first=0):
        q, v, a = (np.ascontiguousarray(x, float) for x in (q, v, a))
        cs = np.ascontiguousarray(contact, np.int32)
        ft = np.ascontiguousarray(f_target, float)
        n = q.shape[0]
        assert q.shape == (n, self.nq) and v.shape == (n, self.nv) and a.shape == (n, self.nv)
        assert cs.shape == (n, self.nfeet) and ft.shape == (n, self.nfeet, 6)
        self._check(self._f("set_target")(self.ctx, C.c_int(first), C.c_int(n), _p(q), _p(v), _p(a), cs.ctypes.data_as(C.POINTER(C.c_int32)), _p(ft)))

    def set_target_centroidal(self, com, vcom, feet_pose, feet_vel, contact, f_target, first=0):
        com, vcom, fp, fv = (np.ascontiguousarray(x, float) for x in (com, vcom, feet_pose, feet_vel))
        cs = np.ascontiguousarray(contact, np.int32)
        ft = np.ascontiguousarray(f_target, float)
        n = com.shape[0]
        assert fp.shape == (n, self.nfeet, 12) and fv.shape == (n, self.nfeet, 6) and ft.shape == (n, self.nfeet, 6)
        self._check(self._f("set_target_centroidal")(
            self.ctx, C.c_int(first), C.c_int(n), _p(com), _p(vcom), _p(fp), _p(fv), cs.ctypes.data_as(C.POINTER(C.c_int32)), _p(ft)))

    def solve(self, t, q, v):
        q, v = np.ascontiguousarray(q, float), np.ascontiguousarray(v, float)
        assert q.shape == (self.B, self.nq) and v.shape == (self.B, self.nv)
        tau = np.zeros((self.B, self.na))
        self._check(self._f("solve")(self.ctx, C.c_double(t), _p(q), _p(v), _p(tau)))
        return tau

    def set_warm_start(self, on):
        self._check(self._f("set_warm_start")(self.ctx, C.c_int(1 if on else 0)))

    def accelerations(self):
        a = np.zeros((self.B, self.nv))
        self._check(self._f("get_accelerations")(self.ctx, C.c_int(0), C.c_int(self.B), _p(a)))
        return a

    def forces(self):
        f = np.zeros((self.B, self.nfeet, MAX_FVAR))
        self._check(self._f("get_forces")(self.ctx, C.c_int(0), C.c_int(self.B), _p(f)))
        return f

    def info(self):
        arr = (IdInfo * self.B)()
        self._check(self._f("get_info")(self.ctx, C.c_int(0), C.c_int(self.B), arr))
        return arr


def integrate_q(q, dv):
    """pinocchio::integrate for [free-flyer; revolute...] configurations (test / bench helper): q (+) dv."""
    q = np.array(q, float)
    out = q.copy()
    v, w = dv[:3], dv[3:6]
    th = np.linalg.norm(w)
    x, y, z, s = q[3:7]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * s), 2 * (x * z + y * s)],
                  [2 * (x * y + z * s), 1 - 2 * (x * x + z * z), 2 * (y * z - x * s)],
                  [2 * (x * z - y * s), 2 * (y * z + x * s), 1 - 2 * (x * x + y * y)]])
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-8:
        V = np.eye(3) + 0.5 * K
        dq = np.array([*(0.5 * w), 1.0])
    else:
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * K + (th - np.sin(th)) / th**3 * K @ K
        dq = np.array([*(np.sin(th / 2) / th * w), np.cos(th / 2)])
    out[:3] = q[:3] + R @ (V @ v)
    x2, y2, z2, s2 = dq
    out[3:7] = [s * x2 + x * s2 + y * z2 - z * y2, s * y2 - x * z2 + y * s2 + z * x2, s * z2 + x * y2 - y * x2 + z * s2,
                s * s2 - x * x2 - y * y2 - z * z2]
    out[3:7] /= np.linalg.norm(out[3:7])
    out[7:] = q[7:] + dv[6:]
    return out


def integrate_q_batch(q, dv):
    """integrate_q for a batch (B x nq, B x nv), vectorised."""
    q = np.asarray(q, float)
    out = q.copy()
    v, w = dv[:, :3], dv[:, 3:6]
    th = np.linalg.norm(w, axis=1)
    small = th < 1e-8
    ths = np.where(small, 1.0, th)
    cb = np.where(small, 0.5, (1 - np.cos(ths)) / ths**2)
    cc = np.where(small, 1.0 / 6.0, (ths - np.sin(ths)) / ths**3)
    wxv = np.cross(w, v)
    Vv = v + cb[:, None] * wxv + cc[:, None] * np.cross(w, wxv)
    x, y, z, s = q[:, 3], q[:, 4], q[:, 5], q[:, 6]
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * s), 2 * (x * z + y * s),
                  2 * (x * y + z * s), 1 - 2 * (x * x + z * z), 2 * (y * z - x * s),
                  2 * (x * z - y * s), 2 * (y * z + x * s), 1 - 2 * (x * x + y * y)], axis=1).reshape(-1, 3, 3)
    out[:, :3] = q[:, :3] + np.einsum("bij,bj->bi", R, Vv)
    sh = np.where(small, 0.5, np.sin(ths / 2) / ths)
    x2, y2, z2 = sh * w[:, 0], sh * w[:, 1], sh * w[:, 2]
    s2 = np.where(small, 1.0, np.cos(ths / 2))
    quat = np.stack([s * x2 + x * s2 + y * z2 - z * y2, s * y2 - x * z2 + y * s2 + z * x2, s * z2 + x * y2 - y * x2 + z * s2,
                     s * s2 - x * x2 - y * y2 - z * z2], axis=1)
    out[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    out[:, 7:] = q[:, 7:] + dv[:, 6:]
    return out
