"""Per-phase cycle split of the Riccati backward kernel (debug build -DSMPC_RIC_PROFILE, see riccati3.cuh RIC_T).
usage (on the GPU box): make -C simple-mpc_b200 BUILD=build_prof OUT=libsmpc_cuda_prof.so SMPC_DEFS=-DSMPC_RIC_PROFILE cuda
                        SMPC_LIB=simple-mpc_b200/libsmpc_cuda_prof.so python tools/ric_phases.py 148 444"""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "simple-mpc_b200", "python"))
import bench  # noqa: E402

NAMES = ["M=I+muP", "chol(I+muP)", "sweeps P~", "P~[A B], w", "[Q S R]+=[A B]^T PAB", "chol(R^)", "YW build + fwd sweep", "Sigma, G build",
         "chol(Sigma)", "sweeps G", "Y += W G", "bwd sweep Y", "gains out", "P update", "symmetrise"]

NAMES4 = ["TMA issue", "chol2(I+muP)", "sweeps P~", "P~[A B], w", "[Q S R]+=[A B]^T PAB", "chol2(R^)", "YW build + fwd sweep", "Sigma, G build",
          "chol2(Sigma)", "sweeps G", "Y += W G", "bwd sweep Y", "gains out", "P update", "symmetrise + M"]
cu = bench.load_cuda()
ric4 = os.environ.get("SMPC_RICCATI", "4") == "4"
prof = cu.smpc_debug_ric4_profile if ric4 else cu.smpc_debug_ric_profile
if ric4:
    NAMES = NAMES4
for B in [int(a) for a in sys.argv[1:]] or [148, 444]:
    sc = bench.make_scenario("c5", B, seed=0)
    b = bench.setup_backend(cu, "smpc", sc, B)
    xm = np.ascontiguousarray(sc.xmeas_start())
    cu.smpc_mpc_iterate(b.ctx, bench.P(xm), 1)
    out = (C.c_ulonglong * 16)()
    prof(out, 1)
    ticks = 5
    for _ in range(ticks):
        cu.smpc_mpc_xmeas_from_prediction(b.ctx)
        cu.smpc_mpc_iterate(b.ctx, None, 1)
    prof(out, 1)
    v = np.array(list(out)[:15], float) / (B * ticks * 50)  # cycles per stage per CTA
    print("B = %d (%s): %.0f cycles per stage per CTA" % (B, cu.smpc_riccati_variant(b.ctx).decode(), v.sum()))
    for n, c in zip(NAMES, v):
        print("  %-28s %8.0f  %5.1f %%" % (n, c, 100 * c / v.sum()))
    b.close()
