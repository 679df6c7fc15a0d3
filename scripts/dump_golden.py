#!/usr/bin/env python
"""dump_golden.py — golden vectors of the MPC::iterate path for configs C1-C4 of SURVEY.md §8d (C5 = C1 over a batch of
seeds), written by EITHER

  --backend reference   the REAL simple-mpc (Aligator 0.16 / Pinocchio 3.8) of the reference's pixi environment
                        (/root/reference/pixi.toml:15-24): `pixi run python scripts/dump_golden.py --backend reference`.
                        This is what pins the parity claim; it cannot run in the build container.
  --backend oracle      this repository's host API (module `simple_mpc` of simple-mpc_b200/) bound to the CPU oracle:
                        regression fixtures, marked "source": "oracle" - they pin the oracle against itself only.
  --backend cuda        the same host API on libsmpc_cuda.so (needs a GPU).

Both sides run the SAME script body: the reference's own example configurations (examples/go2_kinodynamics.py:42-139,
examples/talos_centroidal.py:49-117, examples/go2_fulldynamics.py:43-77, examples/talos_fulldynamics.py:48-135), a
constructor solve, then `--ticks` calls of MPC::iterate with the measured state = previous prediction xs[1] + seeded
noise. Per tick it records x_meas, xs, us, Ks[0:2], prim_infeas, dual_infeas, num_iters, and once the robot itself (joint
tree, inertias, frames) so that tests/test_golden.py can rebuild the identical problem without any URDF.

Robots: --robot real loads example-robot-data (reference backend) ; --robot synthetic uses the repo's go2-like /
talos-like models, handed to Pinocchio through the URDF text smpc_urdf.model_desc_to_urdf() writes, so that BOTH
backends can run the very same robot.

Output: tests/golden/<config>_<robot>_<source>.npz
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-mpc_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "simple-mpc_b200"))

GO2_FEET = ["FL_foot", "FR_foot", "RL_foot", "RR_foot"]
TALOS_FEET = ["left_sole_link", "right_sole_link"]
TALOS_LOCKED = ["arm_left_5_joint", "arm_left_6_joint", "arm_left_7_joint", "gripper_left_joint", "arm_right_5_joint",
                "arm_right_6_joint", "arm_right_7_joint", "gripper_right_joint", "head_1_joint", "head_2_joint"]


# ------------------------------------------------------------------------------------------------
# robots
# ------------------------------------------------------------------------------------------------
def load_robot(backend, which, synthetic):
    """returns (module simple_mpc, model, reference configuration name, base frame name, feet names, export dict)"""
    if backend == "reference":
        import pinocchio as pin
        import simple_mpc as sm
        if synthetic:
            import tempfile
            import smpc_urdf
            from smpc_ctypes import go2_like_model, talos_like_model
            m, lim = go2_like_model() if which == "go2" else talos_like_model()
            feet = GO2_FEET if which == "go2" else TALOS_FEET
            urdf, srdf = smpc_urdf.model_desc_to_urdf(m, feet, limits=lim)
            with tempfile.TemporaryDirectory() as td:
                pu, ps = os.path.join(td, "robot.urdf"), os.path.join(td, "robot.srdf")
                open(pu, "w").write(urdf)
                open(ps, "w").write(srdf)
                model = pin.buildModelFromUrdf(pu, pin.JointModelFreeFlyer())
                pin.loadReferenceConfigurations(model, ps, False)
            return sm, model, "reference", "base", feet, export_pinocchio(pin, model, feet, "base")
        import example_robot_data as erd
        if which == "go2":
            model = erd.load("go2").model
            return sm, model, "standing", "base", GO2_FEET, export_pinocchio(pin, model, GO2_FEET, "base")
        complete = erd.load("talos")
        q0 = complete.model.referenceConfigurations["half_sitting"]
        robot = complete.buildReducedRobot([complete.model.getJointId(n) for n in TALOS_LOCKED], q0)
        return sm, robot.model, "half_sitting", "base_link", TALOS_FEET, export_pinocchio(pin, robot.model, TALOS_FEET, "base_link")
    import simple_mpc as sm  # this repository's module (simple-mpc_b200/simple_mpc*.so)
    model = sm.Model.BuildGo2Like() if which == "go2" else sm.Model.BuildTalosLike()
    return (sm, model, "standing" if which == "go2" else "half_sitting", "base" if which == "go2" else "base_link",
            GO2_FEET if which == "go2" else TALOS_FEET, None)


def export_pinocchio(pin, model, feet, base):
    """the kinematic tree in the flat arrays of include/smpc.h smpc_model_desc (joint 0 = free-flyer root)"""
    nj = model.njoints - 1  # pinocchio joint 0 is the universe
    out = dict(njoints=nj, nq=model.nq, nv=model.nv, parent=np.zeros(nj, int), axis=np.zeros((nj, 3)), place_R=np.zeros((nj, 3, 3)),
               place_p=np.zeros((nj, 3)), mass=np.zeros(nj), lever=np.zeros((nj, 3)), inertia=np.zeros((nj, 3, 3)), names=[])
    axes = {"JointModelRX": [1, 0, 0], "JointModelRY": [0, 1, 0], "JointModelRZ": [0, 0, 1]}
    for i in range(1, model.njoints):
        j = model.joints[i]
        out["names"].append(model.names[i])
        out["parent"][i - 1] = model.parents[i] - 1
        out["place_R"][i - 1] = model.jointPlacements[i].rotation
        out["place_p"][i - 1] = model.jointPlacements[i].translation
        Y = model.inertias[i]
        out["mass"][i - 1], out["lever"][i - 1], out["inertia"][i - 1] = Y.mass, Y.lever, Y.inertia
        sn = j.shortname()
        if i > 1:
            out["axis"][i - 1] = axes[sn] if sn in axes else np.array(j.extract().axis)
    fr = {}
    for name in list(feet) + [base]:
        f = model.frames[model.getFrameId(name)]
        fr[name] = (int(f.parentJoint) - 1, np.array(f.placement.rotation), np.array(f.placement.translation))
    out["frames"] = fr
    out["lower"], out["upper"], out["effort"] = model.lowerPositionLimit[7:], model.upperPositionLimit[7:], model.effortLimit[6:]
    return out


def export_ours(sm, model, mh, feet, base):
    """same arrays from this repository's Model (so that reference- and oracle-written files have one format)"""
    import smpc_ctypes as S
    m, lim = S.go2_like_model() if model.nv == 18 else S.talos_like_model()
    nj = m.njoints
    out = dict(njoints=nj, nq=m.nq, nv=m.nv, parent=np.array(m.parent[:nj]), axis=np.array([m.axis[i][:] for i in range(nj)]),
               place_R=np.array([np.array(m.place_R[i][:]).reshape(3, 3) for i in range(nj)]),
               place_p=np.array([m.place_p[i][:] for i in range(nj)]), mass=np.array(m.mass[:nj]),
               lever=np.array([m.lever[i][:] for i in range(nj)]),
               inertia=np.array([np.array(m.inertia[i][:]).reshape(3, 3) for i in range(nj)]), names=["joint%02d" % i for i in range(nj)])
    fr = {}
    for f, name in enumerate(feet):
        fr[name] = (int(m.foot_joint[f]), np.array(m.foot_R[f][:]).reshape(3, 3), np.array(m.foot_p[f][:]))
    fr[base] = (int(m.base_joint), np.array(m.base_R[:]).reshape(3, 3), np.array(m.base_p[:]))
    out["frames"] = fr
    out["lower"], out["upper"] = lim["qmin"], lim["qmax"]
    out["effort"] = lim.get("umax", np.ones(nj - 1) * 23.7)
    return out


# ------------------------------------------------------------------------------------------------
# the reference's example configurations
# ------------------------------------------------------------------------------------------------
def build_problem(sm, cfg, mh, model):
    nv = model.nv
    g = np.array([0, 0, -9.81])
    lo, hi = np.array(model.lowerPositionLimit), np.array(model.upperPositionLimit)
    lo, hi = (lo[7:], hi[7:]) if lo.size == model.nq else (lo, hi)
    eff = np.array(model.effortLimit)
    eff = eff[6:] if eff.size == nv else eff
    if cfg == "c1":  # examples/go2_kinodynamics.py:42-85
        w_x = np.diag(np.array([0, 0, 100, 10, 10, 0] + [1, 1, 1] * 4 + [10] * 6 + [0.1] * 12, float))
        w_u = np.diag(np.concatenate((np.ones(12) * 0.01, np.ones(nv - 6) * 1e-5)))
        conf = dict(timestep=0.01, w_x=w_x, w_u=w_u, w_cent=np.diag([0, 0, 1, 0.1, 0.1, 10.0]), w_centder=np.diag([0, 0, 0, 0.1, 0.1, 0.1]),
                    gravity=g, force_size=3, w_frame=np.eye(3) * 2000.0, qmin=lo, qmax=hi, mu=0.8, Lfoot=0.01, Wfoot=0.01,
                    kinematics_limits=True, force_cone=False, land_cstr=False)
        ocp = sm.KinodynamicsOCP(conf, mh)
        ocp.createProblem(mh.getReferenceState(), 50, 3, g[2], False)
        mpc = dict(T_fly=30, T_contact=10)
        x0 = mh.getReferenceState()
    elif cfg == "c3":  # examples/go2_fulldynamics.py:43-77 with force_cone / land_cstr switched on
        w_x = np.diag(np.array([0] * 6 + [1, 1, 1] * 4 + [10] * 6 + [0.1] * 12, float))
        conf = dict(timestep=0.01, w_x=w_x, w_u=np.eye(nv - 6) * 1e-4, w_cent=np.diag([0.04, 0.04, 0, 0, 0, 0.0]), gravity=g, force_size=3,
                    w_forces=np.diag([1e-4] * 3), w_frame=np.eye(3) * 1000.0, umin=-eff, umax=eff, qmin=lo, qmax=hi,
                    Kp_correction=np.zeros(3), Kd_correction=np.zeros(3), mu=0.8, Lfoot=0.01, Wfoot=0.01, torque_limits=True,
                    kinematics_limits=True, force_cone=True, land_cstr=True)
        ocp = sm.FullDynamicsOCP(conf, mh)
        ocp.createProblem(mh.getReferenceState(), 50, 3, g[2], False)
        mpc = dict(T_fly=30, T_contact=10)
        x0 = mh.getReferenceState()
    elif cfg == "c2":  # examples/talos_centroidal.py:49-77
        w_lin, w_ang = np.ones(3) * 0.001, np.ones(3) * 0.1
        conf = dict(timestep=0.01, w_u=np.diag(np.concatenate((w_lin, w_ang, w_lin, w_ang))), w_com=np.diag([0.0, 0, 0]),
                    w_linear_mom=np.diag([0.01, 0.01, 100]), w_angular_mom=np.diag([0.1, 0.1, 1000]), w_linear_acc=0.01 * np.eye(3),
                    w_angular_acc=0.01 * np.eye(3), gravity=g, mu=0.8, Lfoot=0.1, Wfoot=0.075, force_size=6)
        ocp = sm.CentroidalOCP(conf, mh)
        dh = sm.RobotDataHandler(mh)
        x0 = dh.getCentroidalState()
        ocp.createProblem(x0, 100, 6, g[2], False)
        mpc = dict(T_fly=80, T_contact=20)
    elif cfg == "c4":  # examples/talos_fulldynamics.py:48-95
        w_x = np.diag(np.array([0, 0, 0, 10, 10, 10] + [0.1] * 12 + [1, 100] + [1, 1, 10, 10] * 2 + [10] * 6 + [1] * 12 + [1, 100] + [10] * 8, float))
        conf = dict(timestep=0.01, w_x=w_x, w_u=np.eye(nv - 6) * 1e-4, w_cent=np.diag([0.1, 0.1, 10, 0.1, 0.1, 10.0]), gravity=g, force_size=6,
                    w_forces=np.diag([0.001] * 6), w_frame=np.eye(6) * 2000.0, umin=-eff, umax=eff, qmin=lo, qmax=hi,
                    Kp_correction=np.array([0, 0, 50.0, 0, 0, 0]), Kd_correction=np.ones(6) * 100.0, mu=0.8, Lfoot=0.1, Wfoot=0.075,
                    torque_limits=True, kinematics_limits=True, force_cone=True, land_cstr=False)
        ocp = sm.FullDynamicsOCP(conf, mh)
        ocp.createProblem(mh.getReferenceState(), 100, 6, g[2], False)
        mpc = dict(T_fly=80, T_contact=20)
        x0 = mh.getReferenceState()
    else:
        raise ValueError(cfg)
    mpc_conf = dict(support_force=-mh.getMass() * g[2], TOL=1e-4, mu_init=1e-8, max_iters=1, num_threads=1, swing_apex=0.15,
                    timestep=0.01, **mpc)
    jsonable = {k: (np.asarray(v).tolist() if isinstance(v, np.ndarray) else v) for k, v in conf.items()}
    return ocp, mpc_conf, jsonable, x0


def gait(feet, T_ds, T_ss):
    if len(feet) == 4:  # trot: FL + RR lift first (examples/go2_kinodynamics.py:117-139)
        quad = {n: True for n in feet}
        p1 = dict(zip(feet, [False, True, True, False]))
        p2 = dict(zip(feet, [True, False, False, True]))
        return [quad] * T_ds + [p1] * T_ss + [quad] * T_ds + [p2] * T_ss
    dbl = {feet[0]: True, feet[1]: True}
    left = {feet[0]: True, feet[1]: False}
    right = {feet[0]: False, feet[1]: True}
    return [dbl] * T_ds + [left] * T_ss + [dbl] * T_ds + [right] * T_ss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["reference", "oracle", "cuda"], required=True)
    ap.add_argument("--configs", default="c1,c2,c3,c4")
    ap.add_argument("--robot", choices=["real", "synthetic"], default=None, help="default: real for the reference backend, synthetic otherwise")
    ap.add_argument("--ticks", type=int, default=12)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()
    synthetic = (args.robot or ("real" if args.backend == "reference" else "synthetic")) == "synthetic"
    if args.backend != "reference" and not synthetic:
        sys.exit("the URDFs of example-robot-data are not available to the b200 backends; use --robot synthetic")
    os.makedirs(args.out, exist_ok=True)
    for cfg in args.configs.split(","):
        which = "go2" if cfg in ("c1", "c3") else "talos"
        sm, model, refname, base, feet, export = load_robot(args.backend, which, synthetic)
        mh = sm.RobotModelHandler(model, refname, base)
        for n in feet:
            if which == "go2":
                mh.addPointFoot(n, base)
            else:
                pts = np.array([[0.1, 0.075, 0], [-0.1, 0.075, 0], [-0.1, -0.075, 0], [0.1, -0.075, 0]])
                mh.addQuadFoot(n, base, pts)
        ocp, mpc_conf, conf_json, x0 = build_problem(sm, cfg, mh, model)
        if args.backend == "reference" or args.backend == "cuda":
            mpc = sm.MPC(mpc_conf, ocp)
        else:
            sys.path.insert(0, os.path.join(ROOT, "tests", "host_api"))
            import _smpc_hooks
            mpc = _smpc_hooks.make_oracle_mpc(mpc_conf, ocp, os.path.join(ROOT, "oracle", "liborc.so"))
        if export is None:
            export = export_ours(sm, model, mh, feet, base)
        mpc.generateCycleHorizon(gait(feet, mpc_conf["T_contact"], mpc_conf["T_fly"]))
        vb = np.zeros(6)
        vb[0] = 0.2 if which == "go2" else 0.1
        mpc.switchToWalk(vb)
        rng = np.random.default_rng([args.seed, int(cfg[1])])
        rec = dict(x_meas=[], xs=[], us=[], Ks=[], prim_infeas=[], dual_infeas=[], num_iters=[])

        def record(xm):
            rec["x_meas"].append(np.array(xm))
            rec["xs"].append(np.array(mpc.xs))
            rec["us"].append(np.array(mpc.us))
            rec["Ks"].append(np.array(mpc.Ks)[:2])  # gains of the first two stages: K0 is what the robot consumes
            r = mpc.solver.results
            rec["prim_infeas"].append(r.prim_infeas)
            rec["dual_infeas"].append(r.dual_infeas)
            rec["num_iters"].append(r.num_iters)

        xref = np.array(mh.getReferenceState())
        record(xref)  # tick 0: the constructor solve
        x = xref.copy()
        nq = model.nq
        for k in range(args.ticks):
            if cfg == "c2":
                # the centroidal handler reduces the measured multibody state itself (src/centroidal-dynamics.cpp:259-262)
                x = xref.copy()
                x[nq:] = rng.normal(size=model.nv) * 0.02
            else:
                x = np.array(mpc.xs[1])
                x[nq:] += rng.normal(size=model.nv) * (0.05 if k == 0 else 0.01)
            mpc.iterate(x)
            record(x)
        source = "simple-mpc" if args.backend == "reference" else args.backend
        meta = dict(config=cfg, source=source, robot=("synthetic-" if synthetic else "") + which, feet=feet, base=base, reference_configuration=refname,
                    problem_conf=conf_json, mpc_conf=mpc_conf, velocity_base=vb.tolist(), ticks=args.ticks, seed=args.seed,
                    versions=versions(args.backend))
        flat = {"meta": json.dumps(meta)}
        for k, v in rec.items():
            flat[k] = np.array(v)
        for k, v in export.items():
            if k == "frames":
                for fn, (j, R, p) in v.items():
                    flat["frame__%s__joint" % fn], flat["frame__%s__R" % fn], flat["frame__%s__p" % fn] = np.array(j), R, p
            elif k == "names":
                flat["model_names"] = np.array(v)
            else:
                flat["model_" + k] = np.asarray(v)
        flat["q_ref"] = xref[:nq]
        path = os.path.join(args.out, "%s_%s_%s.npz" % (cfg, meta["robot"], source))
        np.savez_compressed(path, **flat)
        print("wrote", path, "iters", rec["num_iters"])


def versions(backend):
    if backend != "reference":
        return {"simple-mpc_b200": "this repository"}
    import aligator
    import pinocchio
    import simple_mpc
    return {"simple_mpc": getattr(simple_mpc, "__version__", "?"), "aligator": aligator.__version__, "pinocchio": pinocchio.__version__}


if __name__ == "__main__":
    main()
