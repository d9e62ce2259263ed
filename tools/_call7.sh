mkdir -p gpurun_out/c7
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/c7/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c7/pytest.log
STEPS=20 WARMUP=5 bash tools/run_all_configs.sh gpurun_out/c7/cfg 2> gpurun_out/c7/cfg.err
timeout 600 python tools/fp32_adversarial.py --random-millions 200 --out gpurun_out/c7/r02_fp32_adversarial.json > /dev/null 2> gpurun_out/c7/adv.err; tail -3 gpurun_out/c7/adv.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sw_sim_one -s 4 -c 1 -o gpurun_out/c7/prof_sw_one -f python tools/sw_bench.py --workload native --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/c7/ncu_sw.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bs_map -s 2 -c 1 -o gpurun_out/c7/prof_f64_tma -f python tools/profile_target.py --n 10000000 --fp 8 --runs 6 > gpurun_out/c7/ncu_f64.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bs_map -s 2 -c 1 -o gpurun_out/c7/prof_f32 -f python tools/profile_target.py --n 10000000 --fp 4 --runs 6 > gpurun_out/c7/ncu_f32.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c7/launches_bench_native.csv python bench.py --steps 2 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/c7/launches.log 2>&1
ls gpurun_out/c7/cfg
