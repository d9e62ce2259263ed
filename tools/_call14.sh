mkdir -p gpurun_out/c14
O=gpurun_out/c14
nvidia-smi -L | head -2
( time timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1 ) 2>&1 | grep real; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -7 $O/smoke.log
T64=$PWD/p3arsec_b200/lib/libbs_gpu_tab64.so
for i in 1 2 3; do
python tools/tune_repeat.py --which fp64tma2 --rounds 5 2>&1 | grep -E "fast/tma |PROBE/tma " >> $O/tma_tab256.txt
BS_GPU_LIB=$T64 python tools/tune_repeat.py --which fp64tma2 --rounds 5 2>&1 | grep -E "fast/tma |PROBE/tma " >> $O/tma_tab64.txt
done
echo tab256; cat $O/tma_tab256.txt; echo tab64; cat $O/tma_tab64.txt
for i in 1 2; do
python bench.py --workload native_fp64 --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_fp64_tab256_$i.json
BS_GPU_LIB=$T64 python bench.py --workload native_fp64 --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_fp64_tab64_$i.json
done
for f in $O/bench_fp64_*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(sys.argv[1], 'value %.1f G' % (d['value']/1e9), 'frac', d['roofline']['frac'], 'probe', d['roofline'].get('frac_of_probe'), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'e2e %.1f' % (d['e2e']['value']/1e9))
PY
done
( time python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -2 $O/bench_default.err; cut -c1-600 $O/bench_default.json
( time python bench.py --impl reference > $O/bench_ref_default.json 2> $O/bench_ref_default.err ) 2>&1 | grep real; cut -c1-400 $O/bench_ref_default.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bs_map_tma -s 2 -c 1 -o $O/prof_f64_tma_tab256 -f python tools/profile_target.py --n 10000000 --fp 8 --math fast --runs 6 > $O/ncu_f64.log 2>&1; tail -2 $O/ncu_f64.log
