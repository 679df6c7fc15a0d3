#!/bin/bash
# The command sequence behind profiles/r02w_* (run under gpurun from the repo root): GPU tests, bench lines of every configuration,
# reference arm, ncu launch list and one `ncu --set full` capture of the four top kernels (numbers printed under ncu are not bench values).
set -x
R=${1:-r02w}
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${R}_pytest_gpu.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_c5.json 2> gpurun_out/${R}_c5.err
python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_B1024.csv \
  python bench.py --batch 1024 --steps 2 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${R}_ncu_launch.log 2>&1
# 60 matching launches are skipped: the constructor solve and the first ticks
ncu --set full --clock-control none --import-source on -k regex:'k_riccati3|k_knots|k_forward' --launch-skip 60 --launch-count 4 -f \
  -o gpurun_out/${R}_prof python bench.py --batch 1024 --steps 2 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${R}_ncu_full.log 2>&1
for c in c2 c3 c4; do
  python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${R}_bench_$c.json 2> gpurun_out/${R}_$c.err
done
# 8 x B200 (gpurun --gpus 8): weak scaling and the north-star case
# python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline
# python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --total 4096 --steps 20 --warmup 5 --no-cpu-baseline
# the inverse-dynamics controllers (profiles/r02ar_*): bench lines, cold-solve probe, ncu capture + summary, the chained control stack
python bench.py --config id > gpurun_out/${R}_bench_id.json 2> gpurun_out/${R}_id.err
python bench.py --config id --impl reference > gpurun_out/${R}_bench_id_reference.json 2>> gpurun_out/${R}_id.err
python bench.py --config id --batch 512 --no-cpu-baseline > gpurun_out/${R}_bench_id_b512.json 2>> gpurun_out/${R}_id.err
python tools/id_bench_probe.py > gpurun_out/${R}_id_probe_cold.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_id_solve -s 2 -c 1 -f -o gpurun_out/${R}_ncu_id_b1024 python tools/id_ncu_probe.py 1024 > gpurun_out/${R}_ncu_id.log 2>&1
# (here, after the report came back: python tools/id_ncu_summary.py gpurun_out/${R}_ncu_id_b1024.ncu-rep > profiles/${R}_ncu_id_summary.txt)
for n in 1 256 512; do python tools/control_stack.py $n 100 > gpurun_out/${R}_control_stack_$n.json 2>> gpurun_out/${R}_id.err; done
# python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --config id --no-cpu-baseline

