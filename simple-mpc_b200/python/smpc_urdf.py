"""Minimal URDF / SRDF ingestion: what `pinocchio.buildModelFromUrdf(urdf, JointModelFreeFlyer())` +
`pinocchio.loadReferenceConfigurations(srdf)` + `buildReducedRobot(locked_joints, q)` do for the robots the reference
uses (example-robot-data Go2 / Talos; reference tests/test_utils.cpp:14-97, examples/utils.py:10-31), restated from the
URDF specification and Pinocchio's documented conventions [UPSTREAM, pinocchio 3.8]:

  * a free-flyer root joint ("root_joint") carries the root link; every revolute joint becomes a model joint, in
    depth-first order with the children of a link visited in alphabetical order of their JOINT names (urdfdom keeps the
    joints in a std::map);
  * fixed joints (and joints listed in `lock`, frozen at their reference value) are folded away: the child link's
    inertia is merged into the body of the supporting moving joint and its frame becomes a fixed frame of that joint;
  * every link is a frame (name = link name), placed in the frame of its supporting joint.

Only what the MPC path needs is supported: revolute joints with limits (continuous joints are nq = 2 in Pinocchio and
prismatic ones do not occur in these robots: both raise). Nothing here touches the GPU; it produces
  - `describe(...)`        a plain dict (joint tree, inertias, frames, limits, reference configurations)
  - `to_model(sm, d)`      a `simple_mpc.Model` built through the Python bindings
  - `to_model_desc(d,...)` the ctypes `ModelDesc` of the C ABI (include/smpc.h smpc_model_desc)
"""
import xml.etree.ElementTree as ET
import numpy as np


def _floats(s, n=None, default=None):
    if s is None:
        return np.array(default, float)
    v = np.array([float(t) for t in s.split()], float)
    if n is not None and v.size != n:
        raise ValueError("expected %d numbers, got %r" % (n, s))
    return v


def rpy_to_R(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def axis_angle_R(axis, ang):
    axis = np.asarray(axis, float)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def _origin(el):
    o = el.find("origin") if el is not None else None
    if o is None:
        return np.eye(3), np.zeros(3)
    return rpy_to_R(_floats(o.get("rpy"), 3, [0, 0, 0])), _floats(o.get("xyz"), 3, [0, 0, 0])


def _compose(a, b):
    return a[0] @ b[0], a[0] @ b[1] + a[1]


class _Body:
    """rigid-body inertia accumulated in the frame of its supporting joint: mass, first moment, inertia about the origin"""

    def __init__(self):
        self.m, self.mc, self.Io = 0.0, np.zeros(3), np.zeros((3, 3))

    def add(self, mass, com, Ic):
        self.m += mass
        self.mc += mass * com
        self.Io += Ic + mass * (com @ com * np.eye(3) - np.outer(com, com))

    def about_com(self):
        if self.m <= 0:
            return 0.0, np.zeros(3), np.zeros((3, 3))
        c = self.mc / self.m
        return self.m, c, self.Io - self.m * (c @ c * np.eye(3) - np.outer(c, c))


def parse_srdf(srdf_xml):
    """{state name: {joint name: values}} of every <group_state>"""
    out = {}
    if not srdf_xml:
        return out
    root = ET.fromstring(srdf_xml)
    for gs in root.iter("group_state"):
        out[gs.get("name")] = {j.get("name"): _floats(j.get("value")) for j in gs.findall("joint")}
    return out


def describe(urdf_xml, srdf_xml=None, root_joint="root_joint", lock=(), lock_state=None):
    """Kinematic tree of the URDF with a free-flyer root.
    lock: joint names frozen at their value in reference configuration `lock_state` (0 when absent), like
    buildReducedRobot (reference examples/utils.py:14-27)."""
    robot = ET.fromstring(urdf_xml)
    links = {ln.get("name"): ln for ln in robot.findall("link")}
    joints = robot.findall("joint")
    children, child_links = {}, set()
    for j in joints:
        children.setdefault(j.find("parent").get("link"), []).append(j)
        child_links.add(j.find("child").get("link"))
    roots = [n for n in links if n not in child_links]
    if len(roots) != 1:
        raise ValueError("URDF must have exactly one root link, found %r" % roots)
    states = parse_srdf(srdf_xml)
    locked_at = states.get(lock_state, {}) if lock_state else {}

    out_j = [dict(name=root_joint, parent=-1, type="freeflyer", axis=np.array([0, 0, 1.0]), R=np.eye(3), p=np.zeros(3), body=_Body(),
                  lower=None, upper=None, effort=None)]
    frames = {}

    def add_link_inertia(jidx, T, link):
        inert = link.find("inertial")
        if inert is None:
            return
        Ri, pi = _origin(inert)
        mass = float(inert.find("mass").get("value"))
        I = inert.find("inertia")
        Il = np.array([[float(I.get("ixx")), float(I.get("ixy")), float(I.get("ixz"))],
                       [float(I.get("ixy")), float(I.get("iyy")), float(I.get("iyz"))],
                       [float(I.get("ixz")), float(I.get("iyz")), float(I.get("izz"))]])
        Rc, pc = _compose(T, (Ri, pi))  # inertial frame in the supporting joint's frame
        out_j[jidx]["body"].add(mass, pc, Rc @ Il @ Rc.T)

    def visit(link_name, jidx, T):
        frames[link_name] = (jidx, T[0].copy(), T[1].copy())
        add_link_inertia(jidx, T, links[link_name])
        for j in sorted(children.get(link_name, []), key=lambda e: e.get("name")):
            name, jt = j.get("name"), j.get("type")
            Tj = _compose(T, _origin(j))
            child = j.find("child").get("link")
            if jt == "fixed" or name in lock:
                if name in lock and jt != "fixed":
                    ax = _floats(j.find("axis").get("xyz") if j.find("axis") is not None else None, 3, [1, 0, 0])
                    q = float(locked_at.get(name, [0.0])[0])
                    Tj = _compose(Tj, (axis_angle_R(ax, q), np.zeros(3)))
                frames[name] = (jidx, Tj[0].copy(), Tj[1].copy())
                visit(child, jidx, Tj)
            elif jt == "revolute":
                ax = _floats(j.find("axis").get("xyz") if j.find("axis") is not None else None, 3, [1, 0, 0])
                lim = j.find("limit")
                out_j.append(dict(name=name, parent=jidx, type="revolute", axis=ax / np.linalg.norm(ax), R=Tj[0], p=Tj[1], body=_Body(),
                                  lower=float(lim.get("lower", "-inf")) if lim is not None else -np.inf,
                                  upper=float(lim.get("upper", "inf")) if lim is not None else np.inf,
                                  effort=float(lim.get("effort", "inf")) if lim is not None else np.inf,
                                  velocity=float(lim.get("velocity", "inf")) if lim is not None else np.inf))
                k = len(out_j) - 1
                frames[name] = (k, np.eye(3), np.zeros(3))
                visit(child, k, (np.eye(3), np.zeros(3)))
            else:
                raise ValueError("joint %s: type %r is not supported (revolute and fixed only)" % (name, jt))

    visit(roots[0], 0, (np.eye(3), np.zeros(3)))
    for jd in out_j:
        jd["mass"], jd["lever"], jd["inertia"] = jd.pop("body").about_com()
    nj = len(out_j)
    refs = {}
    for sname, vals in states.items():
        q = np.zeros(7 + nj - 1)
        q[6] = 1.0
        if root_joint in vals and vals[root_joint].size == 7:
            q[:7] = vals[root_joint]
        for i, jd in enumerate(out_j[1:], start=1):
            if jd["name"] in vals:
                q[7 + i - 1] = vals[jd["name"]][0]
        refs[sname] = q
    return dict(joints=out_j, frames=frames, reference_configurations=refs, root_link=roots[0])


def to_model(sm, d):
    """simple_mpc.Model (the stand-in for pinocchio.Model) through the Python bindings"""
    m = sm.Model()
    j0 = d["joints"][0]
    m.addFreeFlyer(j0["name"], j0["mass"], j0["lever"], j0["inertia"])
    for jd in d["joints"][1:]:
        m.addRevoluteJoint(jd["parent"], jd["name"], jd["axis"], sm.SE3(jd["R"], jd["p"]), jd["mass"], jd["lever"], jd["inertia"])
    for name, (jidx, R, p) in d["frames"].items():
        if not m.existFrame(name):
            m.addFrame(sm.Frame(name, jidx, sm.SE3(R, p)))
    m.referenceConfigurations = {k: v.copy() for k, v in d["reference_configurations"].items()}
    rev = d["joints"][1:]
    m.lowerPositionLimit = np.array([jd["lower"] for jd in rev])
    m.upperPositionLimit = np.array([jd["upper"] for jd in rev])
    m.effortLimit = np.array([jd["effort"] for jd in rev])
    m.velocityLimit = np.array([jd.get("velocity", np.inf) for jd in rev])
    return m


def to_model_desc(d, feet, base_frame, reference_state, foot_ref_offsets=None):
    """ctypes ModelDesc (include/smpc.h smpc_model_desc) for the C ABI. feet: frame names in insertion order;
    "<foot>_ref" frames are placed under the base frame where the feet are at the reference configuration
    (reference src/robot-handler.cpp:28-53)."""
    from smpc_ctypes import ModelDesc, fk_numpy
    J = d["joints"]
    m = ModelDesc()
    nj = len(J)
    m.njoints, m.nq, m.nv = nj, 7 + nj - 1, 6 + nj - 1
    for i, jd in enumerate(J):
        m.parent[i] = jd["parent"]
        m.jtype[i] = 0 if i == 0 else 1
        m.idx_q[i] = 0 if i == 0 else 7 + (i - 1)
        m.idx_v[i] = 0 if i == 0 else 6 + (i - 1)
        m.axis[i][:] = jd["axis"]
        m.place_R[i][:] = np.asarray(jd["R"]).ravel()
        m.place_p[i][:] = jd["p"]
        m.mass[i] = jd["mass"]
        m.lever[i][:] = jd["lever"]
        m.inertia[i][:] = np.asarray(jd["inertia"]).ravel()
    m.nfeet = len(feet)
    for f, name in enumerate(feet):
        jidx, R, p = d["frames"][name]
        m.foot_joint[f] = jidx
        m.foot_R[f][:] = R.ravel()
        m.foot_p[f][:] = p
    jb, Rb, pb = d["frames"][base_frame]
    m.base_joint = jb
    m.base_R[:] = Rb.ravel()
    m.base_p[:] = pb
    q = np.asarray(d["reference_configurations"][reference_state], float)
    m.q_ref[: m.nq] = q
    for f in range(m.nfeet):
        m.footref_joint[f] = jb
        m.footref_R[f][:] = np.eye(3).ravel()
    fk = fk_numpy(m, q)
    Mb = fk["base"]
    for f in range(m.nfeet):
        Mf = fk["feet"][f]
        R = Mb[0].T @ Mf[0]
        p = Mb[0].T @ (Mf[1] - Mb[1])
        if foot_ref_offsets is not None:
            p = p + np.asarray(foot_ref_offsets[f], float)
        # frame placements are relative to the supporting JOINT: compose with the base frame's placement
        Rj, pj = _compose((Rb, pb), (R, p))
        m.footref_R[f][:] = Rj.ravel()
        m.footref_p[f][:] = pj
    limits = dict(qmin=np.array([jd["lower"] for jd in J[1:]]), qmax=np.array([jd["upper"] for jd in J[1:]]),
                  umin=-np.array([jd["effort"] for jd in J[1:]]), umax=np.array([jd["effort"] for jd in J[1:]]))
    return m, limits


def model_desc_to_urdf(m, foot_names, base_link="base", limits=None, name="robot"):
    """Write a ModelDesc back as URDF text (one link per joint, feet as fixed child links): lets the synthetic robots of
    smpc_ctypes.py be handed to the REAL simple-mpc / Pinocchio (scripts/dump_golden.py) and gives the reader above a
    round-trip test."""
    def rpy_of(R):
        R = np.asarray(R).reshape(3, 3)
        p = -np.arcsin(np.clip(R[2, 0], -1, 1))
        return np.array([np.arctan2(R[2, 1], R[2, 2]), p, np.arctan2(R[1, 0], R[0, 0])])

    def fmt(v):
        return " ".join("%.17g" % x for x in v)

    names = [base_link] + ["link%d" % i for i in range(1, m.njoints)]
    out = ['<?xml version="1.0"?>', '<robot name="%s">' % name]
    for i in range(m.njoints):
        I = np.array(m.inertia[i][:]).reshape(3, 3)
        out.append('  <link name="%s"><inertial><origin xyz="%s" rpy="0 0 0"/><mass value="%.17g"/>'
                   '<inertia ixx="%.17g" ixy="%.17g" ixz="%.17g" iyy="%.17g" iyz="%.17g" izz="%.17g"/></inertial></link>'
                   % (names[i], fmt(m.lever[i][:]), m.mass[i], I[0, 0], I[0, 1], I[0, 2], I[1, 1], I[1, 2], I[2, 2]))
    for i in range(1, m.njoints):
        lo = limits["qmin"][i - 1] if limits else -3.14
        hi = limits["qmax"][i - 1] if limits else 3.14
        ef = limits["umax"][i - 1] if limits and "umax" in limits else 100.0
        out.append('  <joint name="joint%02d" type="revolute"><parent link="%s"/><child link="%s"/><origin xyz="%s" rpy="%s"/>'
                   '<axis xyz="%s"/><limit lower="%.17g" upper="%.17g" effort="%.17g" velocity="100"/></joint>'
                   % (i, names[m.parent[i]], names[i], fmt(m.place_p[i][:]), fmt(rpy_of(m.place_R[i][:])), fmt(m.axis[i][:]), lo, hi, ef))
    for f, fn in enumerate(foot_names):
        out.append('  <link name="%s"/>' % fn)
        out.append('  <joint name="%s_fixed" type="fixed"><parent link="%s"/><child link="%s"/><origin xyz="%s" rpy="%s"/></joint>'
                   % (fn, names[m.foot_joint[f]], fn, fmt(m.foot_p[f][:]), fmt(rpy_of(m.foot_R[f][:]))))
    out.append("</robot>")
    q = np.array(m.q_ref[: m.nq])
    srdf = ['<?xml version="1.0"?>', '<robot name="%s">' % name, '  <group_state name="reference" group="all">',
            '    <joint name="root_joint" value="%s"/>' % fmt(q[:7])]
    for i in range(1, m.njoints):
        srdf.append('    <joint name="joint%02d" value="%.17g"/>' % (i, q[7 + i - 1]))
    srdf += ["  </group_state>", "</robot>"]
    return "\n".join(out), "\n".join(srdf)
