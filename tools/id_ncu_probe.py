"""One batched inverse-dynamics solve under ncu (tools/round_end_profile.sh): go2-like robot, all tasks, B instances."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-mpc_b200", "python"))
import smpc_ctypes as sc, smpc_id as sid
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m, lim = sc.go2_like_model()
kw = dict(kp_base=10.0, kp_posture=1.0, kp_contact=10.0, w_base=10.0, w_posture=0.1, w_contact_motion=1.0, w_contact_force=1.0)
d = sid.make_desc(m, lim["qmin"], lim["qmax"], 30.0, 23.7, sid.default_settings(**kw))
ib = sid.IdBatch(sc.load_cuda(), "smpc_id_", m, d, B)
ib.set_warm_start(False)  # the profiled launch runs the whole algorithm (6-7 Newton steps), not a one-step warm solve
rng = np.random.default_rng(0)
q = np.repeat(np.array(m.q_ref[: m.nq])[None], B, 0); q[:, 7:] += 0.03 * rng.standard_normal((B, m.nv - 6))
for i in range(3):
    ib.solve(0.0, q, np.zeros((B, m.nv)))
print("newton avg", np.mean([x.newton_iters for x in ib.info()]))
