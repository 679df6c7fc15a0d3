#!/usr/bin/env python
"""bench.py — solved MPC steps/s (batch x Hz) of the MPC tick on B200, and the CPU restatement beside it.

Default workload (BASELINE.json configs[4]): B = 4096 independent Go2-kinodynamics MPC
instances PER GPU (weak scaling; ranks never exchange data on the data path), T = 50 knots, trot
10/30/10/30 with a random gait phase per instance, randomised states, one ProxDDP iteration per tick
(max_iters = 1, reference src/mpc.cpp:91,212). One "step" = one MPC tick of every instance.
`--config c1|c2|c3|c4|c5` selects the other BASELINE.json configurations (c1 = c5 at batch 1; c2 Talos-like
centroidal; c3 Go2-like full dynamics with cones + landing constraints; c4 Talos-like full dynamics, 6-D contacts).
`--total N` is the north-star case: N instances IN TOTAL sharded over the ranks (strong scaling), e.g.
`torchrun ... bench.py --gpus 8 --total 4096`.
Robots are SYNTHETIC (example-robot-data is not available here; simple-mpc_b200/python/smpc_ctypes.py).

One tick = smpc_mpc_iterate = MPC::iterate (reference src/mpc.cpp:189-218) for every instance: forward kinematics of the
measured state, recedeWithCycle, updateStepTrackerReferences (Raibert foothold + Bezier swing references for all T
stages), warm-start shift, one ProxDDP iteration.

  value : ticks with every input already in HBM; the measured state of tick k+1 is the predicted xs[1] of tick k,
          copied device-to-device.
  e2e   : the same tick through the C ABI with HOST buffers: measured states copied H2D from pinned memory,
          (x1, u0, u1, K0) copied D2H every step, next measured state formed on the host.
  roofline : Riccati backward kernel (the dominant kernel), algorithmic bytes of SURVEY.md §8d per launch.
  cpu_baseline : the oracle (CPU restatement, NOT Aligator) built -O3 with FMA (oracle/liborc_fast.so) running the
          SAME tick (orc_mpc_iterate: reference regeneration included) on a bounded sample; three rows as BASELINE.md §3:
          one instance / 1 thread, one instance / 8 threads over knots, one instance per thread on every host core.

`--impl reference` times that CPU tick on all host cores for exactly --steps ticks (the reference's own
Aligator/Pinocchio stack cannot be built or installed here: SURVEY.md §8c).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "simple-mpc_b200", "python"))
import smpc_scenarios as scn  # noqa: E402
import smpc_dist  # noqa: E402
from smpc_ctypes import MpcSettings, SolverOpts, load_cuda  # noqa: E402

P = scn.P
METRIC = "solved_mpc_steps_per_s"
UNIT = "steps/s"

# per configuration: scenario class, default instances per GPU, mean active constraint rows per stage (for the
# algorithmic byte / flop count of the Riccati sweep), sample size of the CPU arm (instances per host thread)
CONFIGS = {
    "c1": dict(scenario="Go2Scenario", batch=1, what="go2-like (synthetic) kinodynamics MPC tick, single instance"),
    "c2": dict(scenario="TalosCentroidalScenario", batch=4096, what="talos-like (synthetic) centroidal MPC tick, 6-D contacts + wrench cones"),
    "c3": dict(scenario="Go2FullScenario", batch=2048, what="go2-like (synthetic) full-dynamics MPC tick, friction pyramid + landing constraints"),
    "c4": dict(scenario="TalosFullScenario", batch=256, what="talos-like (synthetic) full-dynamics walking MPC tick, 6-D contacts, wrench cones, torque + joint limits"),
    "c5": dict(scenario="Go2Scenario", batch=4096, what="go2-like (synthetic) kinodynamics MPC tick"),
    # the low-level controller that runs after MPC::iterate (reference src/inverse-dynamics/*.cpp):
    # its own metric (QP solves per second), handled by run_id_bench
    "id": dict(scenario=None, batch=4096, what="go2-like (synthetic) KinodynamicsID whole-body QP (base + posture + contact tasks, 30 variables, 6 equalities, 44 inequalities)"),
}


def algorithmic_bytes_per_stage(nx, nu, nc):
    k = nu + nc
    return 8 * (2 * nx * nx + 2 * nx * nu + nu * nu + nx + nu + nc * (nx + nu + 1) + nx + (k + 2 * nx) * (nx + 1) + nx * nx + nx)


def algorithmic_flops_per_stage(nx, nu, nc, variant=""):
    """FP64 operations of one proximal Riccati stage as the kernel performs them (a fused multiply-add = 2), phase by
    phase of riccati3.cuh. variant: the name smpc_riccati_variant() returns - "kinostruct" products skip the structural
    zeros of the kinodynamics [A B] (12 dense rows of 2 nv + 60 isolated entries), "D0" drops the W / Sigma phases."""
    nab, ld = nx + nu, nx + 1
    struct, d0 = "kinostruct" in variant, "D0" in variant
    dense_rows = 12 if struct else nx
    extra = 60 if struct else 0                      # non-zeros of the sparse rows
    f = nx ** 3 / 3.0 + 2 * nx * nx * ld            # chol(I + mu P), two sweeps of [p | P]
    f += 2 * nx * (dense_rows * nab + extra) + 2 * nx * nx                        # P~ [A B], w
    f += 2 * (dense_rows * (nx * nab + nu * nu) + extra * nab) + 2 * dense_rows * nab   # [Q S R] += [A B]^T PAB, q, r
    f += nu ** 3 / 3.0 + nu * nu * (ld + (0 if d0 else nc))                        # chol(R^), forward sweep of [Y | W]
    if d0:
        f += 2 * nc * ld                                                          # G = [d | C] / mu
    else:
        f += nu * nc * nc + 2 * nc * nu * ld + nc ** 3 / 3.0 + 2 * nc * nc * ld + 2 * nu * nc * ld   # Sigma, G, chol, sweeps, Y += W G
    f += nu * nu * ld + 2 * nx * (nu + nc) * ld                                    # backward sweep of Y, P update
    return f


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                f = [s.strip() for s in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def load_oracle_fast():
    """The CPU restatement built -O3 with FMA. bench.py executes oracle/ ONLY here, as the timed CPU baseline."""
    path = os.path.join(ROOT, "oracle", "liborc_fast.so")
    if not os.path.exists(path):
        raise RuntimeError("oracle/liborc_fast.so not built: run __graft_entry__.build()")
    lib = C.CDLL(path)
    lib.orc_last_error.restype = C.c_char_p
    return lib


def make_scenario(cfg, batch, seed, first=0):
    return getattr(scn, CONFIGS[cfg]["scenario"])(batch, seed=seed, first=first)


def setup_backend(lib, prefix, sc, batch, device=0, num_threads=1):
    """constructor of MPC for every instance: problem, stage table, constructor solve, driver state"""
    opts = SolverOpts()
    getattr(lib, prefix + "_default_solver_opts")(C.byref(opts))
    opts.num_threads = num_threads  # > 1: PARALLEL (leg-condensed) Riccati sweep with that many legs (reference src/mpc.cpp:46-50)
    b = scn.Backend(lib, prefix, sc.model, sc.ocpb, opts, batch, device)
    table, n = sc.stage_table()  # cycle stages + the standing stage
    rc = getattr(lib, prefix + "_upload_stage_table")(b.ctx, table, n + 1)
    assert rc == 0
    b.set_slots(sc.initial_slots())
    b.set_terminal(sc.terminals())
    b.set_warmstart(*sc.warmstart())
    # constructor solve at the reference state, <= 100 iterations (reference src/mpc.cpp:17,85)
    b.solve(np.tile(sc.xref, (batch, 1)), 100)
    # batched driver: per-instance gait phase, commanded base velocity
    st = MpcSettings(T_fly=sc.T_FLY, T_contact=sc.T_CONTACT, swing_apex=0.15, timestep=0.01, ncycle=n, standing_id=n)
    ph = (C.c_int * batch)(*[int(p) for p in sc.phase])
    xref = np.array(sc.xref, float)
    assert getattr(lib, prefix + "_mpc_setup")(b.ctx, C.byref(st), ph, P(xref)) == 0, getattr(lib, prefix + "_last_error")()
    vb = np.ascontiguousarray(sc.vbase)
    assert getattr(lib, prefix + "_mpc_set_velocity_base")(b.ctx, 0, batch, P(vb)) == 0, getattr(lib, prefix + "_last_error")()
    return b


def run_cpu_ticks(orc, cfg, n_inst, ticks, warm, threads, knot_threads=1):
    """`ticks` whole MPC ticks (orc_mpc_iterate) of n_inst instances of the CPU restatement; OpenMP over instances
    (n_inst > 1) or over the knots of the one instance (knot_threads > 1)."""
    orc.orc_set_num_threads(max(threads, knot_threads))  # torchrun sets OMP_NUM_THREADS=1; the baseline names its own team
    sc = make_scenario(cfg, n_inst, seed=0)
    t0 = time.time()
    b = setup_backend(orc, "orc", sc, n_inst)
    ctor_s = time.time() - t0
    orc.orc_set_knot_threads(b.ctx, knot_threads)
    xm = np.ascontiguousarray(sc.xmeas_start())
    lat = []
    for k in range(warm + ticks):
        t1 = time.perf_counter()
        rc = orc.orc_mpc_iterate(b.ctx, P(xm), 1)
        assert rc == 0, orc.orc_last_error()
        x1 = b.head()[0]
        if k >= warm:
            lat.append(time.perf_counter() - t1)
        xm = np.ascontiguousarray(sc.next_measurement(xm, x1, None))
    wall = float(np.sum(lat))
    b.close()
    return dict(value=n_inst * ticks / wall, wall=wall, ctor_s=ctor_s, n_inst=n_inst, ticks=ticks, ms_per_step=1e3 * wall / ticks,
                p50_ms=float(np.percentile(np.array(lat) * 1e3, 50)))


def cpu_rows(cfg, ticks, warm, per_thread=2):
    """CPU-1, CPU-8 (threads over the knots of one instance), CPU-batch (one instance per thread)."""
    orc = load_oracle_fast()
    cores = os.cpu_count() or 1
    r1 = run_cpu_ticks(orc, cfg, 1, ticks, warm, 1, 1)
    k8 = min(8, cores)
    r8 = run_cpu_ticks(orc, cfg, 1, ticks, warm, k8, k8)
    rb = run_cpu_ticks(orc, cfg, per_thread * cores, ticks, warm, cores, 1)
    return {
        "value": rb["value"], "unit": UNIT, "cores": cores, "kind": "port",
        "sample": "%d instances x %d ticks (after %d warm-up), whole tick incl. reference regeneration (orc_mpc_iterate), OpenMP over "
                  "instances on %d threads; CPU restatement oracle/liborc_fast.so (-O3, AVX2+FMA), not Aligator" % (rb["n_inst"], ticks, warm, cores),
        "rows": {
            "cpu1": {"what": "one instance, 1 thread", "ms_per_tick_p50": r1["p50_ms"], "steps_per_s": r1["value"], "cores": 1},
            "cpu8": {"what": "one instance, OpenMP over knots (model evaluation only; the Riccati sweep is serial)", "ms_per_tick_p50": r8["p50_ms"],
                     "steps_per_s": r8["value"], "cores": k8},
            "batch": {"what": "one instance per thread", "steps_per_s": rb["value"], "cores": cores, "instances": rb["n_inst"],
                      "ms_per_instance_tick": 1e3 * rb["wall"] * cores / (rb["n_inst"] * ticks)},
        },
    }, rb


def id_problem(B, seed):
    """KinodynamicsID on the go2-like robot with the reference test's allTasks settings (tests/inverse-dynamics/
    kinodynamics-id.cpp:277-290); one perturbed measured state per instance."""
    import smpc_ctypes as sct
    import smpc_id as sid
    m, lim = sct.go2_like_model()
    kw = dict(kp_base=10.0, kp_posture=1.0, kp_contact=10.0, w_base=10.0, w_posture=0.1, w_contact_motion=1.0, w_contact_force=1.0)
    d = sid.make_desc(m, lim["qmin"], lim["qmax"], 30.1, 23.7, sid.default_settings(**kw), control_dt=1e-3)
    rng = np.random.default_rng(seed)
    q = np.repeat(np.array(m.q_ref[: m.nq])[None], B, 0)
    q[:, 7:] += 0.03 * rng.standard_normal((B, m.nv - 6))
    q[:, :2] += 0.01 * rng.standard_normal((B, 2))
    v = np.zeros((B, m.nv))
    return sid, m, d, np.ascontiguousarray(q), np.ascontiguousarray(v)


def id_trajectory(ib, sid, q, v, steps, dt=1e-3):
    """closed loop of every instance under its own controller (the integration of the reference's tests,
    tests/inverse-dynamics/kinodynamics-id.cpp:46-60): the measured states one control period apart that the timed region replays"""
    qs, vs = [q.copy()], [v.copy()]
    for t in range(steps - 1):
        ib.solve(t * dt, q, v)
        a = ib.accelerations()
        q = sid.integrate_q_batch(q, (v + a / 2 * dt) * dt)
        v = v + a * dt
        qs.append(q.copy())
        vs.append(v.copy())
    return np.ascontiguousarray(np.stack(qs)), np.ascontiguousarray(np.stack(vs))


def run_id_bench(args, rank, world, local_rank, W, K):
    """bench line of the batched inverse-dynamics QP: one step = one KinodynamicsID::solve of every instance at the next
    measured state of its 1 kHz closed loop (warm-started from the previous solve, as the controller object is in the
    reference's loop)."""
    B = args.batch or CONFIGS["id"]["batch"]
    metric, unit = "id_qp_solves_per_s", "solves/s"
    config = {"workload": "%s, %d instances/GPU, closed loop at 1 kHz, each solve warm-started from the previous one" % (CONFIGS["id"]["what"], B),
              "config": "id", "instances_total": B * world, "instances_this_rank": B,
              "l2": "everything lives in shared memory; HBM carries q, v in and tau, a, f out plus the 0.7 KB warm-start record (3 KB per instance)",
              "parallelism": "instances sharded over %d GPU(s), no data-path collective" % world}
    cores = os.cpu_count() or 1
    T = W + K

    def cpu_leg(steps, warm):
        sid, m, d, q, v = id_problem(8 * cores, 1000)
        orc = load_oracle_fast()
        orc.orc_id_last_error.restype = C.c_char_p
        ib = sid.IdBatch(orc, "orc_id_", m, d, q.shape[0])
        qs, vs = id_trajectory(ib, sid, q, v, warm + steps)
        ib.set_warm_start(False)
        ib.set_warm_start(True)
        for t in range(warm):
            ib.solve(0.0, qs[t], vs[t])
        t0 = time.perf_counter()
        for t in range(warm, warm + steps):
            ib.solve(0.0, qs[t], vs[t])
        dt = time.perf_counter() - t0
        return {"value": q.shape[0] * steps / dt, "unit": unit, "cores": cores, "kind": "port",
                "sample": "%d instances x %d consecutive solves of their closed loops (warm-started), OpenMP over instances; CPU restatement "
                          "oracle/liborc_fast.so (-O3, AVX2+FMA) of the TSID formulation + the same QP method, not TSID/ProxQP" % (q.shape[0], steps)}, dt / steps * 1e3

    if args.impl == "reference":
        if rank != 0:
            return
        cb, ms = cpu_leg(K, W)
        config["instances_total"] = config["instances_this_rank"] = 8 * cores
        print(json.dumps({"impl": "reference", "metric": metric, "value": cb["value"], "unit": unit, "n_gpus": args.gpus, "steps": K, "warmup": W,
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": config, "cpu_baseline": cb,
                          "e2e": {"value": cb["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    # stdout carries exactly one JSON line (see main(): NCCL's version banner is kept off it)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ.pop("NCCL_DEBUG")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cu = load_cuda()
    cu.smpc_id_launch_count.restype = C.c_int64
    sid, m, d, q, v = id_problem(B, 1000 + rank)
    ib = sid.IdBatch(cu, "smpc_id_", m, d, B)
    qs, vs = id_trajectory(ib, sid, q, v, T)
    qd, vd = torch.from_numpy(qs).cuda(), torch.from_numpy(vs).cuda()
    taud = torch.empty((B, m.nv - 6), dtype=torch.float64, device="cuda")
    ms = C.c_float()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def solve_dev(t):
        rc = cu.smpc_id_solve_dev(ib.ctx, C.c_double(0.0), C.c_void_p(qd[t].data_ptr()), C.c_void_p(vd[t].data_ptr()), C.c_void_p(taud.data_ptr()))
        assert rc == 0, cu.smpc_last_error()

    def forget():
        ib.set_warm_start(False)
        ib.set_warm_start(True)

    sampler = ClockSampler(local_rank)
    sampler.start()
    # device-resident arm: the measured states already in HBM; the context's own CUDA events time every launch on its stream
    forget()
    for t in range(W):
        solve_dev(t)
    ib.info()
    barrier()
    l0 = cu.smpc_id_launch_count(ib.ctx)
    kernel_ms = 0.0
    for t in range(W, T):
        solve_dev(t)
        cu.smpc_id_last_solve_ms(ib.ctx, C.byref(ms))  # waits for this launch's end event
        kernel_ms += ms.value
    barrier()
    launches = cu.smpc_id_launch_count(ib.ctx) - l0
    dev_ms = max_over_ranks(kernel_ms / K)
    infos = ib.info()
    # end-to-end arm: host buffers in, torques out, through the C ABI call a user makes
    forget()
    for t in range(W):
        ib.solve(0.0, qs[t], vs[t])
    barrier()
    t0 = time.perf_counter()
    for t in range(W, T):
        tau = ib.solve(0.0, qs[t], vs[t])
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) / K * 1e3)
    # cold-started solves of the last state, for the record
    ib.set_warm_start(False)
    cold_ms = 0.0
    for _ in range(3):
        solve_dev(T - 1)
        cu.smpc_id_last_solve_ms(ib.ctx, C.byref(ms))
        cold_ms += ms.value / 3
    cold_newton = float(np.mean([x.newton_iters for x in ib.info()]))
    sampler.stop_flag = True
    assert all(x.status == 0 for x in infos), "unconverged QP in the bench batch"
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbps_burst", peaks.get("hbm_gbps", 6539.9)))
    ws_bytes = 8.0 * 2 * (2 + 30 + 6 + 44)  # warm-start record read + written
    bytes_per_inst = 8.0 * (m.nq + m.nv + (m.nv - 6) + m.nv + m.nfeet * 12) + 1256 + 48 + ws_bytes
    achieved = bytes_per_inst * B / (dev_ms * 1e-3) / 1e9
    newton = [x.newton_iters for x in infos]
    line = {"metric": metric, "value": B * world / (dev_ms * 1e-3), "unit": unit, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": B * world / (e2e_ms * 1e-3), "unit": unit, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(8 * B * (m.nq + m.nv)), "d2h_bytes_per_step": int(8 * B * (m.nv - 6))},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                         "note": "k_id_solve keeps the whole QP in shared memory (44.5 KB per CTA, 4 CTAs per SM): it is bound by dependent-issue "
                                 "latency and CTA barriers (ncu, profiles/r02an_*), not by HBM or the FP64 pipe"},
            "newton_iters": {"mean": float(np.mean(newton)), "max": int(np.max(newton))},
            "outer_iters_mean": float(np.mean([x.outer_iters for x in infos])),
            "cold_start": {"ms_per_step": cold_ms, "solves_per_s": B * world / (cold_ms * 1e-3), "newton_iters_mean": cold_newton},
            "control_rate_hz_at_this_batch": 1e3 / e2e_ms,
            "clocks": sampler.summary()}
    if not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_leg(10, 3)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="instances per GPU (weak scaling)")
    ap.add_argument("--total", type=int, default=None, help="instances in total, sharded over the ranks (strong scaling)")
    ap.add_argument("--disturb", type=float, default=0.01, help="std of the velocity noise added to the measured state in the e2e arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)
    K = args.steps
    cfg = args.config
    if cfg == "id":
        return run_id_bench(args, rank, world, local_rank, W, K)
    strong = args.total is not None
    if strong:
        first, B = smpc_dist.shard_range(args.total, world, rank)
        total = args.total
    else:
        B = args.batch or CONFIGS[cfg]["batch"]
        first, total = rank * B, B * world
    probe = make_scenario(cfg, 1, seed=0)
    config = {"workload": "%s, %s, T=%d, max_iters=1" % (CONFIGS[cfg]["what"], ("%d instances in total" % total) if strong else ("%d instances/GPU" % B), probe.T),
              "config": cfg, "instances_total": total, "instances_this_rank": B, "horizon": probe.T, "nx": probe.nx, "ndx": probe.ndx, "nu": probe.nu,
              "l2": "per-tick working set (LQ knots + gains, MBs per instance) exceeds the 126 MB L2 for batch >= 64",
              "foot_refs": "regenerated on the device every tick for every instance (Raibert foothold + degree-8 Bezier, all T stages), inside the timed region",
              "parallelism": "instances sharded over %d GPU(s), no data-path collective" % world}

    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        orc = load_oracle_fast()
        r = run_cpu_ticks(orc, cfg, 2 * cores, K, W, cores, 1)
        config["instances_total"] = config["instances_this_rank"] = r["n_inst"]
        config["workload"] = "%s, %d instances (bounded CPU sample of the same workload), T=%d, max_iters=1" % (CONFIGS[cfg]["what"], r["n_inst"], probe.T)
        config["parallelism"] = "OpenMP, one instance per thread, %d threads" % cores
        config["foot_refs"] = "regenerated on the host every tick for every instance, inside the timed region (orc_mpc_iterate)"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d instances x %d ticks (after %d warm-up), whole tick incl. reference regeneration, OpenMP over instances; "
                                       "CPU restatement oracle/liborc_fast.so (-O3, AVX2+FMA), not Aligator" % (r["n_inst"], K, W)},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # stdout carries exactly one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION prints to stdout) is silenced
    # (levels VERSION and WARN both print it; anything more verbose is sent to stderr instead)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ.pop("NCCL_DEBUG")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cu = load_cuda()
    sc = make_scenario(cfg, B, seed=1000, first=first)  # RNG stream = global instance index: the same instance on any rank count
    if world > 1:
        # NCCL carries only shared read-only data: rank 0's robot table is broadcast, every rank checks it matches
        _, same = smpc_dist.broadcast_bytes(dist, bytes(sc.model))
        assert same, "model tables differ across ranks"
    b = setup_backend(cu, "smpc", sc, B, local_rank)
    nxm = sc.model.nq + sc.model.nv
    xm = np.ascontiguousarray(sc.xmeas_start())
    x_pinned = torch.empty((B, nxm), dtype=torch.float64).pin_memory()
    head = [torch.empty(s, dtype=torch.float64).pin_memory() for s in [(B, sc.nx), (B, sc.nu), (B, sc.nu), (B, sc.nu, sc.ndx)]]
    rng = np.random.default_rng(7 + rank)

    def barrier():
        cu.smpc_sync(b.ctx)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def tick_e2e(xm_in):
        x_pinned.numpy()[:] = xm_in
        rc = cu.smpc_mpc_iterate(b.ctx, C.c_void_p(x_pinned.data_ptr()), 1)
        assert rc == 0, cu.smpc_last_error()
        rc = cu.smpc_get_head(b.ctx, 0, B, *[C.c_void_p(h.data_ptr()) for h in head])
        assert rc == 0, cu.smpc_last_error()
        return head[0].numpy()

    def tick_dev():
        if sc.multibody_state:
            cu.smpc_mpc_xmeas_from_prediction(b.ctx)
        rc = cu.smpc_mpc_iterate(b.ctx, None, 1)
        assert rc == 0, cu.smpc_last_error()

    # ---- end-to-end arm ----
    # the synthetic plant disturbance (measured state = prediction + velocity noise) is drawn before the timed region:
    # generating it is input synthesis, not part of the path
    noise = rng.normal(size=(W + K, B, sc.model.nv)) * args.disturb
    lat_e2e = []
    for i in range(W):
        xm = sc.next_measurement(xm, tick_e2e(xm), noise[i])
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        t1 = time.perf_counter()
        xm = sc.next_measurement(xm, tick_e2e(xm), noise[W + i])
        lat_e2e.append(time.perf_counter() - t1)
    barrier()
    e2e_s = time.perf_counter() - t0
    infos = b.info()
    e2e_alphas = np.array([infos[i].alpha for i in range(B)])
    e2e_trials = np.array([infos[i].ls_trials for i in range(B)])
    # ---- device-resident arm (CUDA events on the context stream) ----
    for _ in range(W):
        tick_dev()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    cu.smpc_timer_stats(b.ctx, 1, None, None, None, None)
    l0 = cu.smpc_launch_count(b.ctx)
    cu.smpc_timer_record(b.ctx, 0)
    for _ in range(K):
        tick_dev()
    cu.smpc_timer_record(b.ctx, 1)
    ms = C.c_float()
    cu.smpc_timer_elapsed_ms(b.ctx, C.byref(ms))
    barrier()
    sampler.stop_flag = True
    l1 = cu.smpc_launch_count(b.ctx)
    ric_ms, ric_n, sol_ms, sol_n = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
    cu.smpc_timer_stats(b.ctx, 0, C.byref(ric_ms), C.byref(ric_n), C.byref(sol_ms), C.byref(sol_n))
    kms = (C.c_double * 4)()
    cu.smpc_kernel_times(b.ctx, kms)
    infos = b.info()
    iters = [infos[i].num_iters for i in range(B)]
    alphas = np.array([infos[i].alpha for i in range(B)])
    dev_s = ms.value * 1e-3
    p99_e2e = float(np.percentile(np.array(lat_e2e) * 1e3, 99))
    times = torch.tensor([dev_s, e2e_s, p99_e2e], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_s, e2e_s, p99_e2e = [float(v) for v in times.cpu()]
    sampler.join(timeout=2)

    value = total * K / dev_s
    e2e_value = total * K / e2e_s
    # roofline of the dominant kernel: Riccati backward, one launch per tick
    nc_mean = sc.mean_active_rows()
    bytes_launch = B * sc.T * algorithmic_bytes_per_stage(sc.ndx, sc.nu, nc_mean)
    variant = cu.smpc_riccati_variant(b.ctx).decode()
    flops_launch = B * sc.T * algorithmic_flops_per_stage(sc.ndx, sc.nu, nc_mean, variant)
    flops_dense = B * sc.T * algorithmic_flops_per_stage(sc.ndx, sc.nu, nc_mean)
    ric_avg_s = (ric_ms.value / max(ric_n.value, 1)) * 1e-3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic, traffic_src = None, None
    if cfg in ("c1", "c5"):
        # dram__bytes_read.sum + dram__bytes_write.sum of k_riccati3<36,24,24,128,18,1> from the ncu --set full capture at B = 1024
        # (profiles/r02w_ncu_summary.txt: 2.663 + 1.265 GB per launch), per instance, scaled to this launch
        traffic = (2.663376e9 + 1.264518e9) / 1024 * B
        traffic_src = "ncu --set full at B=1024, per instance x batch (profiles/r02w_ncu_summary.txt)"
    achieved = bytes_launch / ric_avg_s / 1e9 if ric_avg_s > 0 else 0.0
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * dev_s / K, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * nxm * 8,
                "d2h_bytes_per_step": B * 8 * (sc.nx + 2 * sc.nu + sc.nu * sc.ndx), "ms_per_step": 1e3 * e2e_s / K,
                "tick_ms_p50": float(np.percentile(np.array(lat_e2e) * 1e3, 50)), "tick_ms_p99_max_over_ranks": p99_e2e,
                "hz_per_instance": K / e2e_s, "disturb_sigma_v": args.disturb,
                "alpha_lt_1_frac": float((e2e_alphas < 1).mean()), "ls_trials_mean": float(e2e_trials.mean())},
        "gpu_launches": int(l1 - l0),
        "clocks": sampler.summary(),
        "roofline": {"bound": "hbm", "kernel": "proximal Riccati backward sweep, CTA per instance, variant " + cu.smpc_riccati_variant(b.ctx).decode(),
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes": bytes_launch, "launch_ms": 1e3 * ric_avg_s,
                     "share_of_step": ric_ms.value / max(sol_ms.value, 1e-12),
                     "fp64": {"achieved_tflops": flops_launch / ric_avg_s / 1e12 if ric_avg_s > 0 else 0.0,
                              "what": "operations the launched variant performs (structural zeros skipped); dense_equivalent_tflops = the dense "
                                      "stage of the reference algorithm / the same time",
                              "dense_equivalent_tflops": flops_dense / ric_avg_s / 1e12 if ric_avg_s > 0 else 0.0,
                              "nominal_peak_tflops": 37.0,
                              "peak_source": "tools/microbench_fp64.cu: 37.0 TFLOP/s DFMA measured on this pool's B200 (profiles/r01g_microbench_fp64.txt)"}},
        # every kernel of the solve, CUDA events on the context's stream around the first round of each solve, ms per launch
        "kernels": [{"name": n, "launch_ms": kms[i] / max(sol_n.value, 1)} for i, n in enumerate(
            ["linearisation (k_knots: dynamics, costs, constraints, LQ knot)", "Riccati backward sweep", "forward sweep",
             "trial evaluation (k_knots value-only) + linesearch decision"])],
        "solver": {"iters_per_tick_mean": float(np.mean(iters)), "alpha_mean": float(alphas.mean()),
                   "alpha_lt_1_frac": float((alphas < 1).mean())},
    }
    if strong:
        out["target_check"] = {"what": "%d instances in total over %d GPU(s), end to end through host buffers" % (total, world),
                               "hz_per_instance_e2e": K / e2e_s, "tick_ms_e2e_mean": 1e3 * e2e_s / K, "tick_ms_e2e_p99_max_over_ranks": p99_e2e,
                               "hz_per_instance_device": K / dev_s, "target_hz": 100}
    solo = rank == 0 and world == 1 and not strong
    if solo and not args.no_latency:
        # single-instance latency through the C ABI with host buffers (p50 / p99 over 200 ticks). Every row starts from a
        # fresh context on the same scenario, so all rows see the same sequence of ticks
        def latency_row(graph, legs):
            sc1 = make_scenario(cfg, 1, seed=0)
            if legs > 1:
                os.environ["SMPC_LEGS"] = str(legs)  # measure every supported shape, also those for which the serial sweep stays the default
            b1 = setup_backend(cu, "smpc", sc1, 1, local_rank, num_threads=max(legs, 1))
            os.environ.pop("SMPC_LEGS", None)
            row = None
            if legs <= 1 or cu.smpc_riccati_legs(b1.ctx) == legs:
                assert cu.smpc_mpc_use_graph(b1.ctx, graph) == 0
                cu.smpc_timer_stats(b1.ctx, 1, None, None, None, None)
                x1 = np.ascontiguousarray(sc1.xmeas_start())
                lat = []
                for k in range(220):
                    t1 = time.perf_counter()
                    assert cu.smpc_mpc_iterate(b1.ctx, P(x1), 1) == 0, cu.smpc_last_error()
                    h = b1.head()[0]
                    lat.append(time.perf_counter() - t1)
                    x1 = np.ascontiguousarray(sc1.next_measurement(x1, h, None))
                lat = np.array(lat[20:]) * 1e3
                row = {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99))}
                if not graph:  # the per-kernel event timers only exist outside the graph
                    k1, n1 = (C.c_double * 4)(), C.c_int64()
                    cu.smpc_kernel_times(b1.ctx, k1)
                    cu.smpc_timer_stats(b1.ctx, 0, None, None, None, C.byref(n1))
                    row["device_ms"] = {"linearisation": k1[0] / max(n1.value, 1), "riccati": k1[1] / max(n1.value, 1),
                                        "forward": k1[2] / max(n1.value, 1), "trial+decision": k1[3] / max(n1.value, 1)}
                row["riccati_variant"] = cu.smpc_riccati_variant(b1.ctx).decode()
            b1.close()
            return row

        plain, graph = latency_row(0, 1), latency_row(1, 1)
        out["latency_ms"] = {"batch": 1, "p50": plain["p50"], "p99": plain["p99"],
                             "what": "smpc_mpc_iterate + head read-back through host buffers, plain launches, one host synchronisation per tick",
                             "device_ms": plain["device_ms"], "riccati_variant": plain["riccati_variant"],
                             "cuda_graph": {"p50": graph["p50"], "p99": graph["p99"],
                                            "what": "same ticks, smpc_mpc_use_graph(1) - the default for batches up to the SM count: one graph launch per tick, the head of the plan copied to pinned memory by the graph"}}
        # MPCSettings::num_threads > 1 = the PARALLEL (leg-condensed) Riccati sweep, one CTA per leg (csrc/riccati_par.cuh)
        par = []
        for legs in (5, 10, 20):
            rp, rg = latency_row(0, legs), latency_row(1, legs)
            if rp and rg:
                par.append({"legs": legs, "p50": rg["p50"], "p99": rg["p99"], "riccati_device_ms": rp["device_ms"]["riccati"]})
        out["latency_ms"]["parallel_riccati"] = {"what": "same ticks (one graph launch per tick) with smpc_solver_opts.num_threads = legs: four launches per backward sweep, one CTA per leg; riccati_device_ms from the event timers of a plain-launch run", "rows": par}
    if solo and not args.no_latency and cfg == "c5":
        # north_star target: 4096 instances at >= 100 Hz on 8 GPUs = 512 instances per GPU with a tick <= 10 ms.
        # Ticks of a 512-instance context on this one GPU, end to end through host buffers (the 8-GPU run itself is
        # `--gpus 8 --total 4096`).
        scT = make_scenario(cfg, 512, seed=2000)
        bT = setup_backend(cu, "smpc", scT, 512, local_rank)
        xT = np.ascontiguousarray(scT.xmeas_start())
        latT = []
        for k in range(35):
            t1 = time.perf_counter()
            assert cu.smpc_mpc_iterate(bT.ctx, P(xT), 1) == 0
            hT = bT.head()[0]
            latT.append(time.perf_counter() - t1)
            xT = np.ascontiguousarray(scT.next_measurement(xT, hT, None))
        latT = np.array(latT[5:]) * 1e3
        out["target_check"] = {"what": "512 instances per GPU (4096 over 8 GPUs), ticks end to end through host buffers on ONE GPU", "batch_per_gpu": 512,
                               "ms_per_tick_p50": float(np.percentile(latT, 50)), "ms_per_tick_p99": float(np.percentile(latT, 99)),
                               "hz_per_instance": 1e3 / float(latT.mean()), "target_hz": 100}
        bT.close()
    if solo and not args.no_cpu_baseline:
        out["cpu_baseline"], _ = cpu_rows(cfg, 6, 2)
    b.close()
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
