mkdir -p gpurun_out/c9
timeout 600 python tools/tune_repeat.py --which fp64tma2 --rounds 9 > gpurun_out/c9/tune_fp64tma2.txt 2>&1; cat gpurun_out/c9/tune_fp64tma2.txt
BS_GPU_TMA_WIDE=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma or geometry or fp64" > gpurun_out/c9/pytest_shape2.log 2>&1; echo "pytest shape2 rc=$?"; tail -2 gpurun_out/c9/pytest_shape2.log
BS_GPU_TMA_WIDE=2 timeout 400 python tools/sustained.py --fp 8 --n 10000000 --runs 4000 --rois 2 2>&1 | grep "^tma" > gpurun_out/c9/sustained_fp64_shape2.txt; cat gpurun_out/c9/sustained_fp64_shape2.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c9/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c9/pytest.log
timeout 200 python bench.py --workload simsmall --steps 20 --warmup 5 --headline-only --no-cpu-baseline > gpurun_out/c9/simsmall.json 2>&1; cut -c1-400 gpurun_out/c9/simsmall.json
SW_GPU_LIB= timeout 300 ncu --set full --clock-control none --import-source on -k regex:sw_sim_one -s 4 -c 1 -o gpurun_out/c9/prof_sw_composite -f python tools/sw_bench.py --workload native --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/c9/ncu_sw.log 2>&1
