# scratch experiment driver (run under gpurun)
mkdir -p gpurun_out
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('steps/s %.0f  ms/step %.2f  riccati_ms %.2f  share %.3f' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['share_of_step']), d['roofline']['kernel'][-12:])"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== riccati3"; run
echo "== riccati2"; SMPC_RICCATI=2 run
