mkdir -p gpurun_out/c15
O=gpurun_out/c15
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fp64 or tma or degenerate or golden or native" 2>&1 | tail -3
NOILP=$PWD/p3arsec_b200/lib/libbs_gpu_noilp.so
for i in 1 2 3; do
python tools/tune_repeat.py --which fp64tma2 --rounds 5 2>&1 | grep -E "fast/tma|PROBE/tma " >> $O/tma_ilp.txt
BS_GPU_LIB=$NOILP python tools/tune_repeat.py --which fp64tma2 --rounds 5 2>&1 | grep -E "fast/tma|PROBE/tma " >> $O/tma_noilp.txt
done
echo ilp; cat $O/tma_ilp.txt; echo noilp; cat $O/tma_noilp.txt
for i in 1 2; do
python bench.py --workload native_fp64 --steps 20 --warmup 5 --no-ncu 2>/dev/null | tail -1 > $O/bench_fp64_ilp_$i.json
BS_GPU_LIB=$NOILP python bench.py --workload native_fp64 --steps 20 --warmup 5 --no-ncu 2>/dev/null | tail -1 > $O/bench_fp64_noilp_$i.json
BS_GPU_TMA_WIDE=2 python bench.py --workload native_fp64 --steps 20 --warmup 5 --no-ncu 2>/dev/null | tail -1 > $O/bench_fp64_ilp_wide2_$i.json
done
for f in $O/bench_fp64_*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(sys.argv[1], 'value %.1f G' % (d['value']/1e9), 'frac %.3f' % d['roofline']['frac'], 'probe', d['roofline'].get('frac_of_probe'), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'e2e %.1f' % (d['e2e']['value']/1e9))
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bs_map_tma -s 2 -c 1 -o $O/prof_f64_tma_ilp -f python tools/profile_target.py --n 10000000 --fp 8 --math fast --runs 6 > $O/ncu_f64.log 2>&1; tail -2 $O/ncu_f64.log
