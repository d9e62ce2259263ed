mkdir -p gpurun_out/c31
O=gpurun_out/c31
( time timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1 ) 2>&1 | grep real; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
( time python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 ) 2>&1 | grep real; echo "smoke rc=$?"; tail -3 $O/smoke.log
( time python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err ) 2>&1 | grep real; echo "ref rc=$?"; cut -c1-200 $O/bench_reference.json
( time python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -2 $O/bench_default.err
python - $O/bench_default.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][0])
print('value %.1f G' % (d['value']/1e9), 'frac %.3f' % d['roofline']['frac'], 'probe', d['roofline'].get('frac_of_probe'), 'traffic', d['roofline'].get('traffic'), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'e2e %.1f G' % (d['e2e']['value']/1e9), 'run_order %.1f' % (d['e2e_run_order']['value']/1e9), 'launches', d['gpu_launches'])
oc=d['other_configs']
print('simsmall %.2f G e2e %.2f' % (oc['simsmall']['value']/1e9, oc['simsmall']['e2e']['value']/1e9), '| fp64 %.1f G frac %.3f' % (oc['native_fp64']['value']/1e9, oc['native_fp64']['frac_of_measured_peak']), '| sw %.2f G' % (oc['swaptions_native']['value']/1e9), '| strong_1b %.1f G' % (d['strong_1b']['value']/1e9), '| e2e_file wall', d['e2e_file']['ours']['wall_s'])
PY
