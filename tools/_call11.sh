mkdir -p gpurun_out/c11
SEL="price_aos_equals_soa or reference_mode_err_chk or tma_variant_sizes or launch_failure or caf_message"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/c11/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/c11/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "price_aos_equals_soa or tma_variant_sizes" > gpurun_out/c11/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/c11/racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_sw_gpu_parity.py -m gpu -x -q -k "golden or oracle or start_index" > gpurun_out/c11/sw_memcheck.log 2>&1; echo "sw memcheck rc=$?"; tail -3 gpurun_out/c11/sw_memcheck.log
( time python bench.py --steps 20 --warmup 5 > gpurun_out/c11/bench_n1.json 2> gpurun_out/c11/bench_n1.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -2 gpurun_out/c11/bench_n1.err
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c11/bench_ref_n1.json 2> gpurun_out/c11/bench_ref_n1.err ) 2>&1 | grep real
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c11/smoke.log 2>&1; tail -8 gpurun_out/c11/smoke.log
