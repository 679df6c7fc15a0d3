"""MPC at 100 Hz + KinodynamicsID at 1 kHz for a batch of go2-like robots on one GPU (simple-mpc_b200/python/smpc_stack.py):
wall-clock per 10 ms control window. Usage: python tools/control_stack.py [batch] [ticks]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "simple-mpc_b200", "python"))
import bench
import smpc_ctypes as sct
import smpc_stack

raw = "--raw-base-acc" in sys.argv
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(argv[0]) if len(argv) > 0 else 512
ticks = int(argv[1]) if len(argv) > 1 else 60
cu = sct.load_cuda()
sc = bench.make_scenario("c5", B, seed=1000)
b = bench.setup_backend(cu, "smpc", sc, B, 0)
_, lim = sct.go2_like_model()
x0 = sc.xmeas_start()[:, :3].copy()
stat = {"unconverged": 0, "zmin": 1e9}

def on_tick(tick, q, v, tau, kid):
    stat["unconverged"] += sum(x.status != 0 for x in kid.info())
    stat["zmin"] = min(stat["zmin"], float(q[:, 2].min()))

q, v, tau, t = smpc_stack.run(cu, "smpc", cu, "smpc_id_", b, sc, lim, ticks, on_tick=on_tick, zero_base_acc_target=not raw)
adv = q[:, 0] - x0[:, 0]
print(json.dumps({"robots": B, "ticks": ticks, "mpc_ms_per_tick": t["mpc_ms_per_tick"], "id_ms_per_tick_10_solves": t["id_ms_per_tick"],
                  "window_ms": t["mpc_ms_per_tick"] + t["id_ms_per_tick"], "real_time_window_ms": 10.0,
                  "x_advance_m": {"mean": float(adv.mean()), "min": float(adv.min()), "max": float(adv.max())},
                  "commanded_advance_m": {"mean": float((sc.vbase[:, 0] * ticks * 0.01).mean()), "min": float((sc.vbase[:, 0] * ticks * 0.01).min()),
                                          "max": float((sc.vbase[:, 0] * ticks * 0.01).max())},
                  "advance_tracking_rms_m": float(np.sqrt(np.mean((adv - sc.vbase[:, 0] * ticks * 0.01) ** 2))),
                  "base_height_min_m": stat["zmin"], "unconverged_id_solves": int(stat["unconverged"]),
                  "max_abs_tau": float(np.abs(tau).max()),
                  "note": "ideal plant (integrates the QP's accelerations); host-side glue in numpy inside the timed region"}))
