mkdir -p gpurun_out/c19
O=gpurun_out/c19
( time timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1 ) 2>&1 | grep real; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; head -5 $O/smoke.log
for i in 1 2; do python bench.py --workload simsmall --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_simsmall_$i.json; python - $O/bench_simsmall_$i.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print('simsmall value %.2f G  e2e %.2f G (%.3f ms/step)  launches %s' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d['gpu_launches']))
PY
done
