mkdir -p gpurun_out/c4
timeout 600 python tools/tune_repeat.py --which fp64tma --rounds 9 > gpurun_out/c4/tune_fp64tma.txt 2>&1
BS_GPU_TMA_WIDE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma or geometry" > gpurun_out/c4/pytest_wide.log 2>&1; echo "pytest wide rc=$?"
cat gpurun_out/c4/tune_fp64tma.txt; tail -3 gpurun_out/c4/pytest_wide.log
