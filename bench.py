#!/usr/bin/env python
"""bench.py — solved MPC steps/s (batch x Hz) of the Go2-like kinodynamics MPC tick on B200.

Workload (BASELINE.json configs[4]): B = 4096 independent Go2-kinodynamics MPC
instances PER GPU (weak scaling; ranks never exchange data on the data path), T = 50 knots, trot
10/30/10/30 with a random gait phase per instance, randomised states, one ProxDDP iteration per tick
(max_iters = 1, reference src/mpc.cpp:91,212). One "step" = one MPC tick of every instance.
Robot: synthetic go2-like quadruped (example-robot-data is not available here; tests/smpc_ctypes.py).

One tick = smpc_mpc_iterate = MPC::iterate (reference src/mpc.cpp:189-218) for every instance: forward kinematics of the
measured state, recedeWithCycle, updateStepTrackerReferences (Raibert foothold + Bezier swing references for all T
stages), warm-start shift, one ProxDDP iteration.

  value : ticks with every input already in HBM; the measured state of tick k+1 is the predicted xs[1] of tick k,
          copied device-to-device.
  e2e   : the same tick through the C ABI with HOST buffers: measured states copied H2D from pinned memory,
          (x1, u0, u1, K0) copied D2H every step, next measured state formed on the host.
  roofline : Riccati backward kernel (the dominant kernel), algorithmic bytes of SURVEY.md §8d per launch.
  cpu_baseline : the oracle (CPU restatement, NOT Aligator) on a bounded sample, OpenMP over instances.

`--impl reference` times the CPU restatement on the host cores (the reference's own Aligator/Pinocchio
stack cannot be built or installed here: SURVEY.md §8c).
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
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_utils as pu  # noqa: E402
from smpc_ctypes import Slot, SolverOpts, load_cuda, load_oracle  # noqa: E402

METRIC = "solved_mpc_steps_per_s"
UNIT = "steps/s"


def algorithmic_bytes_per_stage(nx, nu, nc):
    k = nu + nc
    return 8 * (2 * nx * nx + 2 * nx * nu + nu * nu + nx + nu + nc * (nx + nu + 1) + nx + (k + 2 * nx) * (nx + 1) + nx * nx + nx)


def algorithmic_flops_per_stage(nx, nu, nc):
    k = nu + nc
    return (7.0 / 3 * nx ** 3 + 4 * nx ** 3 + 4 * nx * nx * nu + 2 * nx * nu * nu + k ** 3 / 3.0 + 2 * k * k * (nx + 1)
            + 2 * nx * k * (nx + 1))


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


def setup_backend(lib, prefix, sc, batch, device=0):
    opts = SolverOpts()
    getattr(lib, prefix + "_default_solver_opts")(C.byref(opts))
    b = pu.Backend(lib, prefix, sc.model, sc.ocpb, opts, batch, device)
    table, n = sc.stage_table()  # 80 cycle stages + the standing stage
    rc = getattr(lib, prefix + "_upload_stage_table")(b.ctx, table, n + 1)
    assert rc == 0
    b.set_slots(sc.initial_slots())
    b.set_terminal(sc.terminals())
    b.set_warmstart(*sc.warmstart())
    # constructor solve at the reference state, <= 100 iterations (reference src/mpc.cpp:17,85)
    b.solve(np.tile(sc.xref, (batch, 1)), 100)
    if prefix == "smpc":
        # batched device-side driver: per-instance gait phase, commanded base velocity
        from smpc_ctypes import MpcSettings
        st = MpcSettings(T_fly=30, T_contact=10, swing_apex=0.15, timestep=0.01, ncycle=n, standing_id=n)
        ph = (C.c_int * batch)(*[int(p) for p in sc.phase])
        xref = np.array(sc.xref, float)
        assert lib.smpc_mpc_setup(b.ctx, C.byref(st), ph, pu.P(xref)) == 0, lib.smpc_last_error()
        vb = np.ascontiguousarray(sc.vbase)
        assert lib.smpc_mpc_set_velocity_base(b.ctx, 0, batch, pu.P(vb)) == 0, lib.smpc_last_error()
    return b


def cycle_ids(lib, prefix, b, sc, tick):
    ids = ((sc.phase + tick) % 80).astype(np.int32)
    rc = getattr(lib, prefix + "_cycle_ids")(b.ctx, 0, b.B, ids.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0, getattr(lib, prefix + "_last_error")()


def run_cpu(args, n_inst, ticks, warm):
    """CPU restatement on the host cores: n_inst instances spread over OpenMP threads."""
    orc = load_oracle()
    orc.orc_set_num_threads(os.cpu_count() or 1)  # torchrun sets OMP_NUM_THREADS=1; the baseline uses every host core
    sc = pu.Go2Scenario(n_inst, seed=0)
    t0 = time.time()
    b = setup_backend(orc, "orc", sc, n_inst)
    ctor_s = time.time() - t0
    x = sc.x_start()
    tick = 0
    for _ in range(warm):
        cycle_ids(orc, "orc", b, sc, tick)
        b.shift()
        b.solve(x, 1)
        x = b.head()[0]
        tick += 1
    lat = []
    t0 = time.time()
    for _ in range(ticks):
        t1 = time.time()
        cycle_ids(orc, "orc", b, sc, tick)
        b.shift()
        b.solve(x, 1)
        x = b.head()[0]
        lat.append(time.time() - t1)
        tick += 1
    wall = time.time() - t0
    cores = orc.orc_num_threads()
    b.close()
    return dict(value=n_inst * ticks / wall, wall=wall, cores=cores, ctor_s=ctor_s, n_inst=n_inst, ticks=ticks,
                ms_per_step=1e3 * wall / ticks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)
    K = args.steps
    config = {"workload": "go2-like (synthetic) kinodynamics MPC tick, %d instances/GPU, T=50, trot 10/30/10/30, max_iters=1" % args.batch,
              "batch_per_gpu": args.batch, "horizon": 50, "nx": 37, "ndx": 36, "nu": 24,
              "l2": "per-tick working set (LQ knots + gains, ~4 MB/instance) exceeds the 126 MB L2 for batch >= 64",
              "foot_refs": "regenerated on the device every tick for every instance (Raibert foothold + degree-8 Bezier, all T stages), inside the timed region",
              "parallelism": "instances sharded over %d GPU(s), no data-path collective" % world}

    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        n_inst = 2 * cores
        r = run_cpu(args, n_inst, max(1, min(K, 8)), min(W, 2))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": "%d instances x %d ticks, OpenMP over instances, solve only (stage-table references, no per-tick reference regeneration); CPU restatement (oracle/), not Aligator" % (r["n_inst"], r["ticks"])},
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
    B = args.batch
    sc = pu.Go2Scenario(B, seed=1000 + rank)
    if world > 1:
        # NCCL carries only shared read-only data: rank 0's robot table is broadcast, every rank checks it matches
        mb = torch.frombuffer(bytearray(bytes(sc.model)), dtype=torch.uint8).cuda()
        ref = mb.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, mb), "model tables differ across ranks"
    b = setup_backend(cu, "smpc", sc, B, local_rank)
    x = sc.x_start()
    x_pinned = torch.empty((B, sc.nx), dtype=torch.float64).pin_memory()
    head = [torch.empty(s, dtype=torch.float64).pin_memory() for s in [(B, sc.nx), (B, sc.nu), (B, sc.nu), (B, sc.nu, sc.ndx)]]
    rng = np.random.default_rng(7 + rank)

    def barrier():
        cu.smpc_sync(b.ctx)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def tick_e2e(tick):
        x_pinned.numpy()[:] = x
        rc = cu.smpc_mpc_iterate(b.ctx, C.c_void_p(x_pinned.data_ptr()), 1)
        assert rc == 0, cu.smpc_last_error()
        rc = cu.smpc_get_head(b.ctx, 0, B, *[C.c_void_p(h.data_ptr()) for h in head])
        assert rc == 0, cu.smpc_last_error()
        return head[0].numpy() + 0.0

    def tick_dev(tick):
        cu.smpc_mpc_xmeas_from_prediction(b.ctx)
        rc = cu.smpc_mpc_iterate(b.ctx, None, 1)
        assert rc == 0, cu.smpc_last_error()

    tick = 0
    # ---- end-to-end arm ----
    # the synthetic plant disturbance (measured state = prediction + velocity noise) is drawn before the timed region:
    # generating it is input synthesis, not part of the path
    noise = rng.normal(size=(W + K, B, 18)) * 0.01
    for i in range(W):
        x = tick_e2e(tick)
        x[:, 19:] += noise[i]
        tick += 1
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        x = tick_e2e(tick)
        x[:, 19:] += noise[W + i]
        tick += 1
    barrier()
    e2e_s = time.perf_counter() - t0
    # ---- device-resident arm (CUDA events on the context stream) ----
    for _ in range(W):
        tick_dev(tick)
        tick += 1
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    cu.smpc_timer_stats(b.ctx, 1, None, None, None, None)
    l0 = cu.smpc_launch_count(b.ctx)
    cu.smpc_timer_record(b.ctx, 0)
    for _ in range(K):
        tick_dev(tick)
        tick += 1
    cu.smpc_timer_record(b.ctx, 1)
    ms = C.c_float()
    cu.smpc_timer_elapsed_ms(b.ctx, C.byref(ms))
    barrier()
    sampler.stop_flag = True
    l1 = cu.smpc_launch_count(b.ctx)
    ric_ms, ric_n, sol_ms, sol_n = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
    cu.smpc_timer_stats(b.ctx, 0, C.byref(ric_ms), C.byref(ric_n), C.byref(sol_ms), C.byref(sol_n))
    infos = b.info()
    iters = [infos[i].num_iters for i in range(B)]
    alphas = np.array([infos[i].alpha for i in range(B)])
    dev_s = ms.value * 1e-3
    times = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_s, e2e_s = [float(v) for v in times.cpu()]
    sampler.join(timeout=2)

    total = B * world
    value = total * K / dev_s
    e2e_value = total * K / e2e_s
    # roofline of the dominant kernel: Riccati backward, one launch per tick
    nc_mean = 18.6  # trot: 12 joint-limit rows + 3 per stance foot (4 feet 20/80, 2 feet 60/80 of the cycle)
    bytes_launch = B * 50 * algorithmic_bytes_per_stage(36, 24, nc_mean)
    flops_launch = B * 50 * algorithmic_flops_per_stage(36, 24, nc_mean)
    ric_avg_s = (ric_ms.value / max(ric_n.value, 1)) * 1e-3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # dram__bytes_read.sum + dram__bytes_write.sum of k_riccati3 from the ncu --set full capture of this tree at
    # B = 1024 (profiles/r01i_ncu_full_B1024.csv: 2.666 + 1.272 GB per launch), per instance, scaled to this launch
    ncu_dram_bytes_per_instance = (2.665893e9 + 1.271974e9) / 1024
    achieved = bytes_launch / ric_avg_s / 1e9 if ric_avg_s > 0 else 0.0
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * dev_s / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * sc.nx * 8,
                "d2h_bytes_per_step": B * 8 * (sc.nx + 2 * sc.nu + sc.nu * sc.ndx), "ms_per_step": 1e3 * e2e_s / K},
        "gpu_launches": int(l1 - l0),
        "clocks": sampler.summary(),
        "roofline": {"bound": "hbm", "kernel": "proximal Riccati backward sweep, CTA per instance, variant " + cu.smpc_riccati_variant(b.ctx).decode(),
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                     "traffic": ncu_dram_bytes_per_instance * B, "traffic_source": "ncu --set full at B=1024, per instance x batch (profiles/r01i_ncu_full_B1024.csv)",
                     "algorithmic_bytes": bytes_launch, "launch_ms": 1e3 * ric_avg_s,
                     "share_of_step": ric_ms.value / max(sol_ms.value, 1e-12),
                     "fp64": {"achieved_tflops": flops_launch / ric_avg_s / 1e12 if ric_avg_s > 0 else 0.0,
                              "nominal_peak_tflops": 37.0,
                              "peak_source": "tools/microbench_fp64.cu: 37.0 TFLOP/s DFMA measured on this pool's B200 (profiles/r01g_microbench_fp64.txt)"}},
        "solver": {"iters_per_tick_mean": float(np.mean(iters)), "alpha_mean": float(alphas.mean()),
                   "alpha_lt_1_frac": float((alphas < 1).mean())},
    }
    if rank == 0 and world == 1 and not args.no_latency:
        # single-instance latency through the C ABI with host buffers (p50 / p99 over 200 ticks)
        sc1 = pu.Go2Scenario(1, seed=0)
        b1 = setup_backend(cu, "smpc", sc1, 1, local_rank)
        x1 = sc1.x_start()
        lat = []
        for k in range(220):
            t1 = time.perf_counter()
            x1 = np.ascontiguousarray(x1)
            assert cu.smpc_mpc_iterate(b1.ctx, pu.P(x1), 1) == 0
            x1 = b1.head()[0]
            lat.append(time.perf_counter() - t1)
        lat = np.array(lat[20:]) * 1e3
        out["latency_ms"] = {"batch": 1, "p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99))}
        b1.close()
    if rank == 0 and world == 1 and not args.no_latency:
        # north_star target: 4096 instances at >= 100 Hz on 8 GPUs = 512 instances per GPU with a tick <= 10 ms.
        # Device-resident ticks of a 512-instance context on this one GPU (CUDA events on the context stream).
        scT = pu.Go2Scenario(512, seed=2000)
        bT = setup_backend(cu, "smpc", scT, 512, local_rank)
        xT = np.ascontiguousarray(scT.x_start())
        assert cu.smpc_mpc_iterate(bT.ctx, pu.P(xT), 1) == 0
        for _ in range(5):
            cu.smpc_mpc_xmeas_from_prediction(bT.ctx)
            assert cu.smpc_mpc_iterate(bT.ctx, None, 1) == 0
        cu.smpc_timer_record(bT.ctx, 0)
        for _ in range(30):
            cu.smpc_mpc_xmeas_from_prediction(bT.ctx)
            assert cu.smpc_mpc_iterate(bT.ctx, None, 1) == 0
        cu.smpc_timer_record(bT.ctx, 1)
        msT = C.c_float()
        cu.smpc_timer_elapsed_ms(bT.ctx, C.byref(msT))
        out["target_check"] = {"what": "512 instances per GPU (4096 over 8 GPUs), device-resident ticks", "batch_per_gpu": 512,
                               "ms_per_tick": msT.value / 30, "hz_per_instance": 30e3 / msT.value, "target_hz": 100}
        bT.close()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        r = run_cpu(args, 2 * cores, 6, 2)
        out["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                               "sample": "%d instances x %d ticks, OpenMP over instances, solve only (stage-table references, no per-tick reference regeneration); CPU restatement (oracle/), not Aligator" % (r["n_inst"], r["ticks"]),
                               "ms_per_instance_tick": 1e3 * r["wall"] * r["cores"] / (r["n_inst"] * r["ticks"])}
    b.close()
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
