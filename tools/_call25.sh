mkdir -p gpurun_out/c25
O=gpurun_out/c25
( time timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1 ) 2>&1 | grep real; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
( time python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 ) 2>&1 | grep real; echo "smoke rc=$?"; tail -3 $O/smoke.log
( time python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err ) 2>&1 | grep real; echo "ref rc=$?"; cut -c1-250 $O/bench_reference.json
( time python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -2 $O/bench_default.err
python - $O/bench_default.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][0])
print('value %.1f G' % (d['value']/1e9), 'frac %.3f' % d['roofline']['frac'], 'probe', d['roofline'].get('frac_of_probe'), 'traffic', d['roofline'].get('traffic'), d['roofline'].get('traffic_source','')[:20], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'e2e %.1f G' % (d['e2e']['value']/1e9), 'run_order %.1f' % (d['e2e_run_order']['value']/1e9), 'launches', d['gpu_launches'])
oc=d['other_configs']
print('simsmall %.2f G e2e %.2f' % (oc['simsmall']['value']/1e9, oc['simsmall']['e2e']['value']/1e9), '| fp64 %.1f G frac %.3f' % (oc['native_fp64']['value']/1e9, oc['native_fp64']['roofline']['frac']), '| sw %.2f G' % (oc['swaptions_native']['value']/1e9), '| strong_1b %.1f G' % (d['strong_1b']['value']/1e9), '| e2e_file', (d.get('e2e_file') or {}).get('ours'))
print('cpu_baseline', d['cpu_baseline'])
PY
SEL="price_aos_equals_soa or tma_variant_sizes or tma_shapes or fp64_against_oracle_sizes or degenerate or launch_failure"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma_variant_sizes or tma_shapes" > $O/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $O/racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_sw_gpu_parity.py -m gpu -x -q -k "golden or oracle or start_index or high_rates or modulus" > $O/sw_memcheck.log 2>&1; echo "sw memcheck rc=$?"; tail -3 $O/sw_memcheck.log
