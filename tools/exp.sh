# scratch experiment driver (run under gpurun): bench summary for the default lib and any alt_*.so
mkdir -p gpurun_out
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('steps/s %.0f  ms/step %.2f  riccati_ms %.2f  share %.3f' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['share_of_step']))"; }
echo "== default"; run
cp simple-mpc_b200/libsmpc_cuda.so /tmp/keep.so
for f in simple-mpc_b200/alt_*.so; do [ -f "$f" ] || continue; echo "== $f"; cp $f simple-mpc_b200/libsmpc_cuda.so; run; done
cp /tmp/keep.so simple-mpc_b200/libsmpc_cuda.so
