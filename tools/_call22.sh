mkdir -p gpurun_out/c22
O=gpurun_out/c22
( time timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1 ) 2>&1 | grep real; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; head -6 $O/smoke.log
HEAD=$PWD/p3arsec_b200/lib/libbs_gpu_head_eb7b6b2.so
for i in 1 2; do
python tools/tune_repeat.py --which fp64tma2 --rounds 5 2>&1 | grep -E "fast|PROBE/tma " >> $O/tma_new.txt
BS_GPU_LIB=$HEAD python tools/tune_repeat.py --which fp64tma2 --rounds 5 2>&1 | grep -E "fast|PROBE/tma " >> $O/tma_head.txt
done
echo new; cat $O/tma_new.txt; echo head; cat $O/tma_head.txt
for i in 1 2 3; do
python bench.py --workload native_fp64 --steps 20 --warmup 5 --no-ncu 2>/dev/null | tail -1 > $O/bench_fp64_new_$i.json
BS_GPU_LIB=$HEAD python bench.py --workload native_fp64 --steps 20 --warmup 5 --no-ncu 2>/dev/null | tail -1 > $O/bench_fp64_head_$i.json
done
for f in $O/bench_fp64_*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(sys.argv[1], 'value %.1f G' % (d['value']/1e9), 'frac %.3f' % d['roofline']['frac'], 'probe', d['roofline'].get('frac_of_probe'), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'e2e %.1f' % (d['e2e']['value']/1e9))
PY
done
python tools/sustained.py --fp 8 --rois 3 2>&1 | tail -18 > $O/sustained_fp64.txt; cat $O/sustained_fp64.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bs_map_tma -s 2 -c 1 -o $O/prof_f64_tma_r2final -f python tools/profile_target.py --n 10000000 --fp 8 --math fast --runs 6 > $O/ncu_f64.log 2>&1; tail -2 $O/ncu_f64.log
