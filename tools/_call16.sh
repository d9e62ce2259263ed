mkdir -p gpurun_out/c16
O=gpurun_out/c16
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fp64 or tma or degenerate or golden" 2>&1 | tail -2
CT=$PWD/p3arsec_b200/lib/libbs_gpu_ctab.so
for i in 1 2 3; do
python tools/tune_repeat.py --which fp64tma2 --rounds 5 2>&1 | grep -E "fast|PROBE/tma " >> $O/tma_gtab.txt
BS_GPU_LIB=$CT python tools/tune_repeat.py --which fp64tma2 --rounds 5 2>&1 | grep -E "fast|PROBE/tma " >> $O/tma_ctab.txt
done
echo gtab; cat $O/tma_gtab.txt; echo ctab; cat $O/tma_ctab.txt
for i in 1 2; do
python bench.py --workload native_fp64 --steps 20 --warmup 5 --no-ncu 2>/dev/null | tail -1 > $O/bench_fp64_gtab_$i.json
BS_GPU_LIB=$CT python bench.py --workload native_fp64 --steps 20 --warmup 5 --no-ncu 2>/dev/null | tail -1 > $O/bench_fp64_ctab_$i.json
BS_GPU_TMA_WIDE=2 python bench.py --workload native_fp64 --steps 20 --warmup 5 --no-ncu 2>/dev/null | tail -1 > $O/bench_fp64_gtab_wide2_$i.json
done
for f in $O/bench_fp64_*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(sys.argv[1], 'value %.1f G' % (d['value']/1e9), 'frac %.3f' % d['roofline']['frac'], 'probe', d['roofline'].get('frac_of_probe'), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'e2e %.1f' % (d['e2e']['value']/1e9))
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bs_map_tma -s 2 -c 1 -o $O/prof_f64_tma_gtab -f python tools/profile_target.py --n 10000000 --fp 8 --math fast --runs 6 > $O/ncu_f64.log 2>&1; tail -2 $O/ncu_f64.log
