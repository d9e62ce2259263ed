mkdir -p gpurun_out/c12
BS_GPU_LIB=$PWD/p3arsec_b200/lib/libbs_gpu_arriveall.so timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma_variant_sizes" > gpurun_out/c12/racecheck_arriveall.log 2>&1; echo "racecheck arrive-all rc=$?"; tail -3 gpurun_out/c12/racecheck_arriveall.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "price_aos_equals_soa or reference_mode_err or fp64_against_oracle_sizes" > gpurun_out/c12/racecheck_rest.log 2>&1; echo "racecheck rest rc=$?"; tail -3 gpurun_out/c12/racecheck_rest.log
timeout 900 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "price_aos or reference_m or tma_variant_sizes or launch_failure or caf_message or fp64_against_oracle_sizes or degenerate or err_chk_reference" > gpurun_out/c12/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/c12/memcheck.log
for i in 1 2 3; do
python tools/tune_repeat.py --which fp64tma2 --rounds 5 2>&1 | grep "fast/tma " >> gpurun_out/c12/tma_default.txt
BS_GPU_LIB=$PWD/p3arsec_b200/lib/libbs_gpu_arriveall.so python tools/tune_repeat.py --which fp64tma2 --rounds 5 2>&1 | grep "fast/tma " >> gpurun_out/c12/tma_arriveall.txt
done
echo default; cat gpurun_out/c12/tma_default.txt; echo arriveall; cat gpurun_out/c12/tma_arriveall.txt
