#!/usr/bin/env python
"""dump_golden_id.py — golden vectors of the inverse-dynamics controllers (KinodynamicsID / CentroidalID, reference
src/inverse-dynamics/*.cpp), the companion of dump_golden.py:

  --backend reference   the REAL simple-mpc (TSID 1.9 / proxsuite 0.7.2) in the reference's pixi environment:
                        `pixi run python scripts/dump_golden_id.py --backend reference [--robot real|synthetic]`.
                        This is what would pin the parity claim of include/smpc_id.h; it cannot run in the build container.
  --backend cuda        this repository's Python module `simple_mpc` (host classes over libsmpc_cuda.so; needs a GPU):
                        regression fixtures marked "source": "cuda".

Both run the SAME body through the same Python API (KinodynamicsIDSettings / KinodynamicsID / CentroidalID, reference
bindings/expose-inverse-dynamics.cpp:62-94): the scenarios of the reference's tests (tests/inverse-dynamics/*.cpp: all tasks, hard
contacts, quad feet, a swing foot) run in closed loop with the tests' integration; per control period the file holds the measured
(q, v) and the returned torques and accelerations, plus the settings, the limits and the robot (dump_golden.py's format).
tests/test_golden.py replays tests/golden/id_*.npz through the CUDA path: 1e-9 against its own fixtures, 1e-4 relative against
the real reference (TSID's ProxQP wrapper stops at an absolute accuracy of 1e-5).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
sys.path.insert(0, os.path.join(ROOT, "simple-mpc_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "simple-mpc_b200"))
import dump_golden as dg  # noqa: E402
from smpc_id import integrate_q  # noqa: E402  (pure numpy: the same integration on both sides)

DT = 1e-3
QUAD_POINTS = np.array([[0.1, 0.075, 0], [-0.1, 0.075, 0], [-0.1, -0.075, 0], [0.1, -0.075, 0]])
SCENARIOS = {
    # reference tests/inverse-dynamics/kinodynamics-id.cpp:277-309 (allTasks), :198-213 (contactPoint_equality), :160-177 (contactQuad_cost)
    "kino_alltasks_go2": dict(robot="go2", kind="kino", s=dict(kp_base=10.0, kp_posture=1.0, kp_contact=10.0, w_base=10.0, w_posture=0.1,
                                                                w_contact_force=1.0, w_contact_motion=1.0)),
    "kino_contact_eq_go2": dict(robot="go2", kind="kino", s=dict(kp_base=1.0, kp_contact=10.0, w_base=1.0, w_contact_motion=10.0, w_contact_force=1.0,
                                                                  contact_motion_equality=True)),
    "kino_quad_talos": dict(robot="talos", kind="kino", s=dict(kp_base=1.0, kp_posture=1.0, kp_contact=10.0, w_base=1.0, w_posture=0.05,
                                                                w_contact_motion=10.0, w_contact_force=1.0)),
    # reference tests/inverse-dynamics/centroidal-id.cpp:344-409 (footTrackingTask)
    "centroidal_swing_go2": dict(robot="go2", kind="centroidal", lift=0, s=dict(kp_feet_tracking=5.0, kp_posture=0.1, kp_contact=1.0, w_feet_tracking=1e3,
                                                                                 w_posture=1.0, w_contact_force=0.1, contact_motion_equality=True)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["reference", "cuda"], required=True)
    ap.add_argument("--scenarios", default=",".join(SCENARIOS))
    ap.add_argument("--robot", choices=["real", "synthetic"], default=None)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()
    synthetic = (args.robot or ("real" if args.backend == "reference" else "synthetic")) == "synthetic"
    if args.backend != "reference" and not synthetic:
        sys.exit("the URDFs of example-robot-data are not available to the b200 backend; use --robot synthetic")
    os.makedirs(args.out, exist_ok=True)
    for name in args.scenarios.split(","):
        sc = SCENARIOS[name]
        which = sc["robot"]
        sm, model, refname, base, feet, export = dg.load_robot(args.backend, which, synthetic)
        mh = sm.RobotModelHandler(model, refname, base)
        for n in feet:
            if which == "go2":
                mh.addPointFoot(n, base)
            else:
                mh.addQuadFoot(n, base, QUAD_POINTS)
        if export is None:
            export = dg.export_ours(sm, model, mh, feet, base)
        settings = sm.CentroidalIDSettings() if sc["kind"] == "centroidal" else sm.KinodynamicsIDSettings()
        for k, v in sc["s"].items():
            setattr(settings, k, v)
        ctrl = (sm.CentroidalID if sc["kind"] == "centroidal" else sm.KinodynamicsID)(mh, DT, settings)
        nq, nv = model.nq, model.nv
        xref = np.array(mh.getReferenceState())
        q, v = xref[:nq].copy(), np.zeros(nv)
        if which == "go2" and sc["kind"] == "kino":  # the tests' crouched start
            for leg in range(4):
                q[7 + 3 * leg + 1] += 0.1
                q[7 + 3 * leg + 2] -= 0.2
            q[2] -= 0.02
        target = {}
        if sc["kind"] == "centroidal":
            dh = sm.RobotDataHandler(mh)
            dh.updateInternalData(xref, False)
            poses = [dh.getFootPose(i) for i in range(len(feet))]
            lift = sc["lift"]
            R, p = np.array(poses[lift].rotation), np.array(poses[lift].translation) + np.array([0, 0, 0.03])
            poses[lift] = sm.SE3(R, p)
            contact = [i != lift for i in range(len(feet))]
            w = mh.getMass() * 9.81 / (len(feet) - 1)
            forces = [np.array([0.0, 0.0, w if c else 0.0]) for c in contact]
            com = np.array(dh.getData().com[0]) if hasattr(dh.getData(), "com") else None
            ctrl.setTarget(com, np.zeros(3), poses, [np.zeros(6)] * len(feet), contact, forces)
            target = dict(com=com, lift=lift, lift_R=R, lift_p=p, contact=np.array(contact), forces=np.array(forces),
                          poses_R=np.array([np.array(P.rotation) for P in poses]), poses_p=np.array([np.array(P.translation) for P in poses]))
        rec = dict(q=[], v=[], tau=[], a=[])
        for i in range(args.steps):
            tau = np.array(ctrl.solve(i * DT, q, v))
            a = np.array(ctrl.getAccelerations())
            for k, x in (("q", q), ("v", v), ("tau", tau), ("a", a)):
                rec[k].append(np.array(x))
            q = integrate_q(q, (v + a / 2 * DT) * DT)
            v = v + a * DT
        source = "simple-mpc" if args.backend == "reference" else args.backend
        lim = dict(lower=np.array(model.lowerPositionLimit)[-(nv - 6):], upper=np.array(model.upperPositionLimit)[-(nv - 6):],
                   velocity=np.array(model.velocityLimit)[-(nv - 6):], effort=np.array(model.effortLimit)[-(nv - 6):])
        meta = dict(scenario=name, kind=sc["kind"], source=source, robot=("synthetic-" if synthetic else "") + which, feet=feet, base=base,
                    reference_configuration=refname, settings={k: (bool(v) if isinstance(v, bool) else float(v)) for k, v in sc["s"].items()},
                    control_dt=DT, steps=args.steps, versions=versions(args.backend))
        flat = {"meta": json.dumps(meta), "q_ref": xref[:nq]}
        for k, x in rec.items():
            flat[k] = np.array(x)
        for k, x in lim.items():
            flat["limit_" + k] = x
        for k, x in target.items():
            if x is not None:
                flat["target_" + k] = np.asarray(x)
        for k, x in export.items():
            if k == "frames":
                for fn, (j, R, p) in x.items():
                    flat["frame__%s__joint" % fn], flat["frame__%s__R" % fn], flat["frame__%s__p" % fn] = np.array(j), R, p
            elif k == "names":
                flat["model_names"] = np.array(x)
            else:
                flat["model_" + k] = np.asarray(x)
        path = os.path.join(args.out, "id_%s_%s_%s.npz" % (name, meta["robot"], source))
        np.savez_compressed(path, **flat)
        print("wrote", path, "max |tau| %.3f" % np.abs(flat["tau"]).max())


def versions(backend):
    if backend != "reference":
        return {"simple-mpc_b200": "this repository"}
    import simple_mpc
    import tsid
    return {"simple_mpc": getattr(simple_mpc, "__version__", "?"), "tsid": getattr(tsid, "__version__", "?")}


if __name__ == "__main__":
    main()
