mkdir -p gpurun_out/c13
timeout 300 python -m pytest tests/test_sw_gpu_parity.py -m gpu -x -q -s 2>&1 | tail -8 > gpurun_out/c13/pytest_sw.log; cat gpurun_out/c13/pytest_sw.log
for i in 1 2 3; do python tools/sw_bench.py --workload native --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/c13/sw_native_$i.json; python -c "
import json; d=json.load(open('gpurun_out/c13/sw_native_$i.json')); print('native', round(d['value']/1e9,3), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"; done
python tools/sw_bench.py --workload simlarge --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('simlarge', round(d['value']/1e9,3))"
python tools/sw_bench.py --workload native --mode lean --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lean', round(d['value']/1e9,3))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sw_sim_one -s 4 -c 1 -o gpurun_out/c13/prof_sw_paired -f python tools/sw_bench.py --workload native --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/c13/ncu_sw.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "driver_binary or err_chk or tma or fp64" 2>&1 | tail -3
