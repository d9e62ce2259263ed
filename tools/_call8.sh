mkdir -p gpurun_out/c8
timeout 300 python -m pytest tests/test_sw_gpu_parity.py -m gpu -x -q > gpurun_out/c8/pytest_sw.log 2>&1; echo "pytest sw rc=$?"; tail -2 gpurun_out/c8/pytest_sw.log
for i in 1 2; do
python tools/sw_bench.py --workload native --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c8/sw_twologs_$i.json 2>/dev/null
SW_GPU_LIB=$PWD/p3arsec_b200/lib/libsw_gpu_composite.so python tools/sw_bench.py --workload native --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c8/sw_composite_$i.json 2>/dev/null
done
python tools/sw_bench.py --workload simlarge --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c8/sw_twologs_simlarge.json 2>/dev/null
SW_GPU_LIB=$PWD/p3arsec_b200/lib/libsw_gpu_composite.so python tools/sw_bench.py --workload simlarge --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c8/sw_composite_simlarge.json 2>/dev/null
for f in gpurun_out/c8/sw_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e9,3),'G trials/s', round(d['ms_per_step'],3),'ms')"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sw_sim_one -s 4 -c 1 -o gpurun_out/c8/prof_sw_twologs -f python tools/sw_bench.py --workload native --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/c8/ncu_sw.log 2>&1
timeout 400 python tools/sustained.py --fp 8 --n 10000000 --runs 4000 --rois 2 > gpurun_out/c8/sustained_fp64.txt 2>&1; cat gpurun_out/c8/sustained_fp64.txt
timeout 200 python tools/simsmall_pdl.py > gpurun_out/c8/simsmall_pdl.txt 2>&1; cat gpurun_out/c8/simsmall_pdl.txt
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/c8/bench_n1.json 2> gpurun_out/c8/bench_n1.err; tail -c 1500 gpurun_out/c8/bench_n1.json; tail -3 gpurun_out/c8/bench_n1.err
