#!/usr/bin/env python
"""Long closed-loop parity runs, CUDA vs the CPU oracle (test infrastructure: the oracle is the checker). Worst normwise relative
difference over every instance and tick; iteration counts, accepted step sizes and linesearch trial counts must stay equal.
    python tools/long_run_parity.py > gpurun_out/long_run_parity.json      (needs a GPU; about two minutes)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "simple-mpc_b200", "python"), os.path.join(ROOT, "tests")]
import parity_utils as pu  # noqa: E402

runs = [
    ("kinodynamics go2, serial sweep", dict(batch=16, ticks=120, seed=11)),
    ("full dynamics go2, friction pyramid + landing rows", dict(batch=4, ticks=40, seed=12, scenario=pu.Go2FullScenario)),
    ("centroidal talos, serial sweep", dict(batch=4, ticks=60, seed=13, scenario=pu.TalosCentroidalScenario)),
    ("centroidal talos, leg-parallel sweep (num_threads = 8) vs the SERIAL oracle", dict(batch=4, ticks=60, seed=13, scenario=pu.TalosCentroidalScenario, cuda_threads=8)),
    ("kinodynamics talos, 6-D wrenches", dict(batch=3, ticks=30, seed=14, scenario=pu.TalosKinoScenario, force_cone=False)),
    ("full dynamics talos, 6-D contacts", dict(batch=2, ticks=20, seed=15, scenario=pu.TalosFullScenario)),
]
out = []
for name, kw in runs:
    r = pu.run_parity(**kw)
    out.append(dict(run=name, batch=kw["batch"], ticks=kw["ticks"], xs=r["xs"], us=r["us"], K=r["K"], u0_elem=r["u0_elem"], K0_elem=r["K0_elem"],
                    iters_equal=r["iters_equal"], alphas_and_trials_equal=r["alphas_equal"]))
    print(json.dumps(out[-1]), file=sys.stderr)
print(json.dumps(out, indent=1))
