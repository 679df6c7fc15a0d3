"""The C-ABI shared library loads on a GPU-less box and exports every entry point include/smpc.h declares
(no compute call is made here). The oracle exports the mirrored orc_* names used by the parity tests."""
import ctypes as C
import os
import re
from smpc_ctypes import ROOT, load_cuda, ModelDesc, OcpDesc, SolverOpts, Slot, Terminal, SolveInfo
from oracle_libs import load_oracle


def declared_functions():
    names = set()
    for header in ("smpc.h", "smpc_id.h"):
        src = open(os.path.join(ROOT, "include", header)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(smpc_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = load_cuda()
    names = declared_functions()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_oracle_mirrors_the_solver_level_entry_points():
    orc = load_oracle()
    for n in ["create", "destroy", "set_slots", "set_slot", "set_foot_refs", "set_terminal", "cycle", "upload_stage_table",
              "cycle_ids", "set_warmstart", "shift_warmstart", "solve", "get_results", "get_head", "get_info", "get_duals",
              "get_ncmax", "get_xdot", "default_solver_opts", "problem_dims", "last_error"]:
        assert hasattr(orc, "orc_" + n), n


def test_oracle_mirrors_the_inverse_dynamics_entry_points():
    orc = load_oracle()
    for n in ["create", "destroy", "default_settings", "set_target", "set_target_centroidal", "solve", "get_accelerations",
              "get_forces", "get_info", "set_warm_start", "last_error"]:
        assert hasattr(orc, "orc_id_" + n), n
    import smpc_id
    a, b = smpc_id.IdSettings(), smpc_id.IdSettings()
    load_cuda().smpc_id_default_settings(C.byref(a))
    orc.orc_id_default_settings(C.byref(b))
    assert bytes(a) == bytes(b) == bytes(smpc_id.default_settings())
    assert C.sizeof(smpc_id.IdInfo) == 6 * 4 + 2 * 8 and C.sizeof(smpc_id.IdSettings) == 8 * 10 + 2 * 4 + 8 * 4


def test_struct_sizes_match_the_header():
    # sizes computed from the header's field list (8-byte alignment, no padding surprises)
    assert C.sizeof(Slot) == 8 + 8 * (64 + 40 + 4 * 12 + 4 * 6 + 4)
    assert C.sizeof(Terminal) == 8 * (64 + 3 + 1)
    assert C.sizeof(SolveInfo) == 6 * 4 + 7 * 8
    assert C.sizeof(SolverOpts) == 8 * 2 + 4 * 2 + 8 * 18 + 8
    assert C.sizeof(ModelDesc) % 8 == 0 and C.sizeof(OcpDesc) % 8 == 0


def test_default_options_are_identical_in_library_and_oracle():
    a, b = SolverOpts(), SolverOpts()
    load_cuda().smpc_default_solver_opts(C.byref(a))
    load_oracle().orc_default_solver_opts(C.byref(b))
    assert bytes(a) == bytes(b)
    assert a.tol == 1e-4 and a.mu_init == 1e-8 and a.max_iters == 1  # reference include/simple-mpc/mpc.hpp:39-42


def test_no_device_means_loud_failure_not_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    import parity_utils as pu
    lib = load_cuda()
    sc = pu.Go2Scenario(1, seed=0)
    o = SolverOpts()
    lib.smpc_default_solver_opts(C.byref(o))
    ctx = C.c_void_p()
    rc = lib.smpc_create(C.byref(ctx), 0, 1, C.byref(sc.model), C.byref(sc.ocpb.desc), C.byref(o))
    assert rc == -3, rc  # SMPC_ERR_CUDA
    assert b"no CPU fallback" in lib.smpc_last_error() or b"cuda" in lib.smpc_last_error().lower()


def test_inverse_dynamics_context_also_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    import smpc_id
    from smpc_ctypes import go2_like_model
    lib = load_cuda()
    m, lim = go2_like_model()
    d = smpc_id.make_desc(m, lim["qmin"], lim["qmax"], 30.0, 23.7, smpc_id.default_settings(w_posture=1.0))
    ctx = C.c_void_p()
    rc = lib.smpc_id_create(C.byref(ctx), C.byref(m), C.byref(d), C.c_int(1))
    assert rc == -3 and not ctx.value, rc  # SMPC_ERR_CUDA, no context handed out
    assert b"cuda" in lib.smpc_last_error().lower()
    # bad arguments are rejected before any CUDA call
    assert lib.smpc_id_create(C.byref(ctx), C.byref(m), C.byref(d), C.c_int(0)) == -1
    d.control_dt = 0.0
    assert lib.smpc_id_create(C.byref(ctx), C.byref(m), C.byref(d), C.c_int(1)) == -1
