mkdir -p gpurun_out/c21
O=gpurun_out/c21
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 300 python -m pytest tests -m gpu -x -q -k "shard or devices or multi or aos_sharded" > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -2 $O/pytest_multi.log
for n in $N 4; do
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 > $O/bench_n$n.json 2> $O/bench_n$n.err ) 2>&1 | grep real; echo "bench $n rc=$?"
python - $O/bench_n$n.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][0])
print('value %.1f G' % (d['value']/1e9), d['scaling'], 'e2e %.1f G' % (d['e2e']['value']/1e9), 'frac_copy', d['e2e'].get('frac_of_copies_in_then_out'), 'native_weak %.1f G e2e %.1f G' % (d['native_weak']['value']/1e9, d['native_weak']['e2e']['value']/1e9), 'inproc %.1f G biteq %s' % (d['inproc']['value']/1e9, d['inproc']['bit_equal_to_single_device']), 'e2e_file wall', ((d.get('e2e_file') or {}).get('ours') or {}).get('wall_s'), d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
done
timeout 300 python bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/bench_ref_n$N.json 2> /dev/null; cut -c1-200 $O/bench_ref_n$N.json
for n in 4 $N; do python tools/sw_bench.py --gpus $n --workload native --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/sw_native_inproc_n$n.json; python -c "
import json; d=json.load(open('$O/sw_native_inproc_n$n.json')); print('sw inproc n=$n', round(d['value']/1e9,2), 'G trials/s, e2e', round(d['e2e']['value']/1e9,2))"; done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29549 tools/sw_bench.py --gpus $N --workload native --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/sw_native_torchrun_n$N.json; python -c "
import json; d=json.load(open('$O/sw_native_torchrun_n$N.json')); print('sw torchrun n=$N', round(d['value']/1e9,2), 'G trials/s, e2e', round(d['e2e']['value']/1e9,2))"
