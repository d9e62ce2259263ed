mkdir -p gpurun_out/c18
O=gpurun_out/c18
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "fp64_reference or fp64_reference_math or reference_mode" > $O/pytest_ref.log 2>&1; echo "rc=$?"; grep -E "fp64|passed|failed|Error|assert" $O/pytest_ref.log | tail -20
python tools/tune_repeat.py --which fp64ref --rounds 3 2>&1 | tail -6 | tee $O/fp64_ref_speed.txt
