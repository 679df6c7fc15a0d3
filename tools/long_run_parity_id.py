"""Long closed-loop parity runs of the inverse-dynamics controllers: libsmpc_cuda (smpc_id_*) vs the oracle (orc_id_*) on identical
measurements, the loop driven by the CUDA solution (companion of tools/long_run_parity.py). Prints one JSON object.
Usage (GPU box): python tools/long_run_parity_id.py [steps]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-mpc_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import smpc_ctypes as sc
import smpc_id as sid
import test_inverse_dynamics as t
from oracle_libs import load_oracle

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cuda, orc = sc.load_cuda(), load_oracle()
out = {}
for name, kw in t.SCENARIOS:
    B = 8 if name == "go2" else 4
    n = steps if name == "go2" else steps // 4
    ig, m, lim, vmax, eff = t.make(cuda, "smpc_id_", name, B, **kw)
    io, *_ = t.make(orc, "orc_id_", name, B, **kw)
    q, v = t.perturbed_starts(m, B, 3, name), np.zeros((B, m.nv))
    w = dict(tau=0.0, a=0.0, newton_cuda=0, newton_orc=0, unconverged=0, iter_mismatch=0)
    for i in range(n):
        tg, to = ig.solve(i * t.DT, q, v), io.solve(i * t.DT, q, v)
        ag, ao = ig.accelerations(), io.accelerations()
        ic, ir = ig.info(), io.info()
        for b in range(B):
            w["tau"] = max(w["tau"], t.relerr(tg[b], to[b], max(np.abs(to[b]).max(), 1.0)))
            w["a"] = max(w["a"], t.relerr(ag[b], ao[b], max(np.abs(ao[b]).max(), 1.0)))
            w["newton_cuda"] += ic[b].newton_iters
            w["newton_orc"] += ir[b].newton_iters
            w["unconverged"] += int(ic[b].status != 0) + int(ir[b].status != 0)
            w["iter_mismatch"] += int(ic[b].newton_iters != ir[b].newton_iters)
        q, v = t.step_state(q, v, ag)
    w["solves"] = n * B
    key = "%s %s" % (name, "centroidal" if kw.get("centroidal") else ("hard contacts" if kw.get("contact_motion_equality") else "contacts as costs"))
    out[key] = w
print(json.dumps(out, indent=1))
