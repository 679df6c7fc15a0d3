"""Quick timing probe of the batched inverse-dynamics QP (used under gpurun while developing; bench.py --config id is the
reported number)."""
import sys, os, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-mpc_b200", "python"))
import smpc_ctypes as sc, smpc_id as sid

cuda = sc.load_cuda()
for name, B in (("go2", 148), ("go2", 4096), ("talos", 148), ("talos", 2048)):
    if name == "go2":
        m, lim = sc.go2_like_model(); vmax, eff = 30.0, 23.7
        kw = dict(kp_base=10.0, kp_posture=1.0, kp_contact=10.0, w_base=10.0, w_posture=0.1, w_contact_motion=1.0, w_contact_force=1.0)
    else:
        m, lim = sc.talos_like_model(); vmax, eff = 10.0, 300.0
        kw = dict(kp_base=1.0, kp_posture=1.0, kp_contact=10.0, w_base=1.0, w_posture=0.05, w_contact_motion=10.0, w_contact_force=1.0)
    d = sid.make_desc(m, lim["qmin"], lim["qmax"], vmax, eff, sid.default_settings(**kw))
    ib = sid.IdBatch(cuda, "smpc_id_", m, d, B)
    ib.set_warm_start(False)
    rng = np.random.default_rng(0)
    q = np.repeat(np.array(m.q_ref[: m.nq])[None], B, 0); q[:, 7:] += 0.03 * rng.standard_normal((B, m.nv - 6))
    v = np.zeros((B, m.nv))
    ms = C.c_float()
    ts = []
    for i in range(8):
        t0 = time.perf_counter(); tau = ib.solve(0.0, q, v); t1 = time.perf_counter()
        cuda.smpc_id_last_solve_ms(ib.ctx, C.byref(ms)); ts.append((ms.value, (t1 - t0) * 1e3))
        a = ib.accelerations()
        q = np.stack([sid.integrate_q(q[b], (v[b] + a[b] / 2 * 1e-3) * 1e-3) for b in range(B)]) if B <= 148 else q
        v = v + a * 1e-3 if B <= 148 else v
    inf = ib.info()
    print(name, "B", B, "kernel ms", ["%.3f" % t[0] for t in ts[2:]], "e2e ms", ["%.3f" % t[1] for t in ts[2:]],
          "newton avg %.1f max %d" % (np.mean([x.newton_iters for x in inf]), max(x.newton_iters for x in inf)),
          "unconverged", sum(x.status != 0 for x in inf), flush=True)
