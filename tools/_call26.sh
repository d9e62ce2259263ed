mkdir -p gpurun_out/c26
O=gpurun_out/c26
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fp64 or err_chk or golden or tma or geometry or degenerate" > $O/pytest_fp64.log 2>&1 ) 2>&1 | grep real; echo "pytest rc=$?"; tail -2 $O/pytest_fp64.log
PREV=$PWD/p3arsec_b200/lib/libbs_gpu_prev.so
for i in 1 2; do
python tools/errchk_perf.py 2>&1 | tail -4 >> $O/errchk_new.txt
BS_GPU_LIB=$PREV python tools/errchk_perf.py 2>&1 | tail -4 >> $O/errchk_prev.txt
python tools/tune_repeat.py --which fp64tma2 --rounds 5 2>&1 | grep -E "fast|PROBE/tma " >> $O/tma_new.txt
BS_GPU_LIB=$PREV python tools/tune_repeat.py --which fp64tma2 --rounds 5 2>&1 | grep -E "fast|PROBE/tma " >> $O/tma_prev.txt
done
echo errchk new; cat $O/errchk_new.txt; echo errchk prev; cat $O/errchk_prev.txt
echo new; cat $O/tma_new.txt; echo prev; cat $O/tma_prev.txt
