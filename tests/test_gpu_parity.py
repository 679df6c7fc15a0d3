"""GPU parity: the CUDA library vs the CPU oracle through the C ABI, same seeded inputs.

Tolerance (stated by BASELINE.json north_star): <= 1e-8 relative on controls and gains, same iteration count.
The relative error is normwise per matrix: max|a-b| / max|b|.
"""
import pytest
import parity_utils as pu

pytestmark = pytest.mark.gpu
TOL = 1e-8


def test_kino_go2_single_instance():
    r = pu.run_parity(batch=1, ticks=6, seed=0, verbose=True)
    assert r["iters_equal"] and r["alphas_equal"], r
    assert r["us"] <= TOL and r["K"] <= TOL and r["xs"] <= TOL, r


def test_kino_go2_batch_random_phases():
    r = pu.run_parity(batch=8, ticks=4, seed=1, verbose=True)
    assert r["iters_equal"] and r["alphas_equal"], r
    assert r["us"] <= TOL and r["K"] <= TOL and r["xs"] <= TOL, r


def test_kino_go2_cone_land_terminal():
    r = pu.run_parity(batch=2, ticks=4, seed=2, verbose=True, force_cone=True, land_cstr=True, terminal_constraint=True)
    assert r["iters_equal"], r
    assert r["us"] <= 1e-6 and r["K"] <= 1e-6, r
