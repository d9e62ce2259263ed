mkdir -p gpurun_out/c30
O=gpurun_out/c30
L=$PWD/p3arsec_b200/lib
for i in 1 2; do
for v in default g2 g5 g3t5 g3t3; do
if [ $v = default ]; then unset SW_GPU_LIB; else export SW_GPU_LIB=$L/libsw_gpu_$v.so; fi
python tools/sw_bench.py --workload native --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/sw_native_${v}_$i.json
python -c "
import json; d=json.load(open('$O/sw_native_${v}_$i.json')); print('$v $i', round(d['value']/1e9,3), 'G trials/s, kernels ms', round(d['ms_per_step'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'], 'parity', d.get('parity_spot_max_rel'))"
done; done
