mkdir -p gpurun_out/c27
O=gpurun_out/c27
for i in 1 2 3; do
for w in 2 1 0; do
BS_GPU_TMA_WIDE=$w python bench.py --workload native_fp64 --steps 20 --warmup 5 --no-ncu 2>/dev/null | tail -1 > $O/bench_fp64_shape${w}_$i.json
python - $O/bench_fp64_shape${w}_$i.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(sys.argv[1].split('/')[-1], 'value %.1f G' % (d['value']/1e9), 'frac %.3f' % d['roofline']['frac'], 'probe %.3f' % d['roofline'].get('frac_of_probe'), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'e2e %.1f' % (d['e2e']['value']/1e9))
PY
done; done
