mkdir -p gpurun_out/c24
O=gpurun_out/c24
( time timeout 600 python -m pytest tests/test_sw_gpu_parity.py -m gpu -x -q > $O/pytest_sw_gpu.log 2>&1 ) 2>&1 | grep real; echo "pytest rc=$?"; tail -3 $O/pytest_sw_gpu.log
HEAD=$PWD/p3arsec_b200/lib/libsw_gpu_head_eb7b6b2.so
PREV=$PWD/p3arsec_b200/lib/libsw_gpu_prev_3734a99.so
show() { python -c "
import json,sys; d=json.load(open('$1')); print('$1', round(d['value']/1e9,3), 'G trials/s, e2e', round(d['e2e']['value']/1e9,2), 'kernels ms', d.get('ms_per_step'), d['clocks']['sm_mhz'], d['clocks']['reasons'], 'frac', d['roofline'].get('frac'), 'alg', (d['roofline'].get('algorithmic') or {}).get('frac'))"; }
for i in 1 2 3; do
python tools/sw_bench.py --workload native --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/sw_native_new_$i.json; show $O/sw_native_new_$i.json
SW_GPU_LIB=$PREV python tools/sw_bench.py --workload native --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/sw_native_prev_$i.json; show $O/sw_native_prev_$i.json
done
SW_GPU_LIB=$HEAD python tools/sw_bench.py --workload native --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/sw_native_head_1.json; show $O/sw_native_head_1.json
for w in simlarge simmedium simsmall; do
python tools/sw_bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/sw_${w}_new.json; show $O/sw_${w}_new.json
done
python tools/sw_bench.py --workload native --mode lean --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/sw_native_lean_new.json; show $O/sw_native_lean_new.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sw_sim_one -s 4 -c 1 -o $O/prof_sw_one_r2final2 -f python tools/sw_bench.py --workload native --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_sw.log 2>&1; tail -1 $O/ncu_sw.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sw_sim_fast -c 1 -o $O/prof_sw_lean_r2final -f python tools/sw_bench.py --workload native --mode lean --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_sw_lean.log 2>&1; tail -1 $O/ncu_sw_lean.log
