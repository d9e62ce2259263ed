mkdir -p gpurun_out/c17
O=gpurun_out/c17
( time timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1 ) 2>&1 | grep real; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
for i in 1 2 3; do
python bench.py --workload native_fp64 --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_fp64_default_$i.json
BS_GPU_TMA_WIDE=0 python bench.py --workload native_fp64 --steps 20 --warmup 5 --no-ncu 2>/dev/null | tail -1 > $O/bench_fp64_shape0_$i.json
done
for f in $O/bench_fp64_*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(sys.argv[1], 'value %.1f G' % (d['value']/1e9), 'frac %.3f' % d['roofline']['frac'], 'probe', d['roofline'].get('frac_of_probe'), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'e2e %.1f' % (d['e2e']['value']/1e9), d['config'].get('blocks'), d['config'].get('threads_per_block'))
PY
done
python tools/sustained.py --help > /dev/null 2>&1 && python tools/sustained.py --fp 8 2>&1 | tail -12 > $O/sustained_fp64.txt; cat $O/sustained_fp64.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bs_map_tma -s 2 -c 1 -o $O/prof_f64_tma_shape2 -f python tools/profile_target.py --n 10000000 --fp 8 --math fast --runs 6 > $O/ncu_f64.log 2>&1; tail -2 $O/ncu_f64.log
( time python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -2 $O/bench_default.err; cut -c1-300 $O/bench_default.json
