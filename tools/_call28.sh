mkdir -p gpurun_out/c28
O=gpurun_out/c28
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 300 python -m pytest tests -m gpu -x -q -k "shard or devices or multi or aos_sharded" > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -2 $O/pytest_multi.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err ) 2>&1 | grep real; echo "bench $N rc=$?"
python - $O/bench_n$N.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][0])
print('value %.1f G' % (d['value']/1e9), d['scaling'], 'e2e %.1f G' % (d['e2e']['value']/1e9), 'frac_copy', d['e2e'].get('frac_of_copies_in_then_out'), 'native_weak %.1f G e2e %.1f G' % (d['native_weak']['value']/1e9, d['native_weak']['e2e']['value']/1e9), 'inproc %.1f G biteq %s' % (d['inproc']['value']/1e9, d['inproc']['bit_equal_to_single_device']), 'e2e_file wall', ((d.get('e2e_file') or {}).get('ours') or {}).get('wall_s'), d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
python tools/sw_bench.py --gpus $N --workload native --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/sw_native_inproc_n$N.json; python -c "
import json; d=json.load(open('$O/sw_native_inproc_n$N.json')); print('sw inproc n=$N', round(d['value']/1e9,2), 'G trials/s, e2e', round(d['e2e']['value']/1e9,2))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29549 tools/sw_bench.py --gpus $N --workload native --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/sw_native_torchrun_n$N.json; python -c "
import json; d=json.load(open('$O/sw_native_torchrun_n$N.json')); print('sw torchrun n=$N', round(d['value']/1e9,2), 'G trials/s, e2e', round(d['e2e']['value']/1e9,2))"
python tools/sw_bench.py --gpus 1 --workload native --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/sw_native_inproc_n1.json; python -c "
import json; d=json.load(open('$O/sw_native_inproc_n1.json')); print('sw inproc n=1', round(d['value']/1e9,2), 'G trials/s, e2e', round(d['e2e']['value']/1e9,2))"
