"""The reference's C++ benchmark program (reference benchmark/go2.cpp) written against this repo's host classes
(simple-mpc_b200/benchmark/go2.cpp): builds the Go2 FullDynamicsOCP + MPC exactly as the reference program does, then times
MPC::iterate. GPU test: the program runs from C++ (no Python in the loop) and produces a finite plan."""
import json
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "simple-mpc_b200", "benchmark")


@pytest.mark.parametrize("prog", ["go2", "talos"])
def test_benchmark_program_is_built(prog):
    assert os.path.exists(os.path.join(BENCH, prog)), "run __graft_entry__.build()"


@pytest.mark.gpu
@pytest.mark.parametrize("prog", ["go2", "talos"])
def test_benchmark_program_runs_the_mpc_loop_from_cpp(prog):
    r = subprocess.run([os.path.join(BENCH, prog), "60"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["ticks"] == 60 and out["iterate_ms_p50"] > 0
    assert 0 < out["us0_norm"] < 1e5  # joint torques of a standing / walking robot: finite and not absurd
    print(out)
