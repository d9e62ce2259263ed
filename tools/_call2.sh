set -x
mkdir -p gpurun_out/c2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "fp64 or reference or aos or launch_failure or degenerate" > gpurun_out/c2/pytest.log 2>&1; echo "pytest rc=$?"
timeout 600 python tools/tune_repeat.py --which fp64r2 --rounds 5 > gpurun_out/c2/tune_fp64.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bs_map -s 2 -c 1 -o gpurun_out/c2/prof_f64 -f python tools/profile_target.py --n 10000000 --fp 8 --runs 6 > gpurun_out/c2/ncu.log 2>&1
timeout 600 python tools/fp32_adversarial.py --random-millions 50 --out gpurun_out/c2/r02_fp32_adversarial_quick.json > /dev/null 2> gpurun_out/c2/adv.err
tail -5 gpurun_out/c2/pytest.log; cat gpurun_out/c2/tune_fp64.txt; tail -3 gpurun_out/c2/adv.err
