set -x
mkdir -p gpurun_out/c3
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "fp64 or reference or degenerate or geometry" > gpurun_out/c3/pytest.log 2>&1; echo "pytest rc=$?"
timeout 600 python tools/tune_repeat.py --which fp64r2 --rounds 7 > gpurun_out/c3/tune_fp64.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bs_map -s 2 -c 1 -o gpurun_out/c3/prof_f64_pipe -f python tools/profile_target.py --n 10000000 --fp 8 --runs 6 > gpurun_out/c3/ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bs_map -s 2 -c 1 -o gpurun_out/c3/prof_f64_nopipe -f python tools/profile_target.py --n 10000000 --fp 8 --runs 6 --threads 256 > gpurun_out/c3/ncu2.log 2>&1
tail -5 gpurun_out/c3/pytest.log; cat gpurun_out/c3/tune_fp64.txt
