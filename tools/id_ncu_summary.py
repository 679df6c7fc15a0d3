#!/usr/bin/env python
"""Text summary of an `ncu --set full` capture of k_id_solve (tools/id_ncu_probe.py): key metrics of the launch and the stall
samples / executed instructions by region of csrc/id_qp.cuh (via tools/ncu_lines.py). Usage:
    tools/id_ncu_summary.py <report.ncu-rep> > profiles/<round>_ncu_id_summary.txt"""
import collections, csv, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg']
print("ncu --set full --clock-control none, k_id_solve (KinodynamicsID QP, go2-like robot, all tasks), B = 1024 instances, COLD-started solves")
print("(tools/id_ncu_probe.py 1024; report %s, not committed). One launch:\n" % rep)
for i, h in enumerate(hdr):
    if h in want:
        print("%-90s %-14s %s" % (h, units[i], vals[i]))
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "k_id_solve",
                      os.path.join(ROOT, "simple-mpc_b200", "libsmpc_cuda.so"), "400"], stdout=subprocess.PIPE, text=True).stdout
src = open(os.path.join(ROOT, "simple-mpc_b200", "csrc", "id_qp.cuh")).read().split("\n")
line_of = lambda pat: next(i + 1 for i, l in enumerate(src) if pat in l)
marks = sorted([
    (line_of("SMPC_DEV void id_dynamics"), 'dynamics (FK, CRBA, RNEA)'), (line_of("SMPC_DEV void id_frame"), 'base frame / task accumulation'),
    (line_of("SMPC_DEV void id_build_qp"), 'QP assembly'), (line_of("SMPC_DEV void id_block_reduce"), 'CTA reductions'),
    (line_of("SMPC_DEV void id_al_value"), 'subproblem value (H x, A x, C x)'), (line_of("SMPC_DEV void id_solve_qp"), 'solver set-up, pair table'),
    (line_of("// gradient (hx = H x is current)"), 'gradient + active-row list'), (line_of("// Newton matrix (lower triangle)"), 'Newton matrix build'),
    (line_of("// right-looking LDL^T, one barrier per column"), 'LDL^T factorisation (+ forward substitution)'),
    (line_of("// dx = U^-T D^-1 z"), 'back substitution (one warp)'), (line_of("id_block_reduce<false>(L, sm, 1, SC_SLOPE)") - 8, 'line search / accept'),
    (line_of("// multiplier update; relative KKT residuals"), 'multiplier update, KKT residuals'),
    (line_of("// One instance: q, v (global)"), 'epilogue (tau, a, f)')])
reg, inst = collections.Counter(), collections.Counter()
for l in out.split("\n"):
    m = re.match(r'\s*([\d.]+)%\s+(\d+) samples\s+(\d+) inst\s+(\S+):(\d+)', l)
    if not m:
        continue
    key = m.group(4)
    if key == 'id_qp.cuh':
        key = '(helpers)'
        for a, n in marks:
            if int(m.group(5)) >= a:
                key = n
    reg[key] += int(m.group(2))
    inst[key] += int(m.group(3))
tot, ti = sum(reg.values()), sum(inst.values())
print("\nstall samples / executed warp-instructions by region of csrc/id_qp.cuh (tools/ncu_lines.py):")
for k, v in reg.most_common():
    print("  %-48s %5.1f %% samples  %5.1f %% instructions" % (k, 100 * v / tot, 100 * inst[k] / ti))
print("\nhottest source lines:")
print("\n".join(out.split("\n")[1:13]))
