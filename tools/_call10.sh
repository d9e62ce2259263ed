mkdir -p gpurun_out/c10
N=$(nvidia-smi -L | wc -l)
timeout 600 python tools/e2e_file_bench.py --gpus $N --reps 3 --ref-sample 300000 > gpurun_out/c10/e2e_file_auto.json 2> gpurun_out/c10/e2e_auto.err; cut -c1-700 gpurun_out/c10/e2e_file_auto.json
BS_GPU_DEVICES=2 timeout 600 python tools/e2e_file_bench.py --gpus $N --reps 2 --ref-sample 0 > gpurun_out/c10/e2e_file_2gpus.json 2>> gpurun_out/c10/e2e_auto.err; cut -c1-500 gpurun_out/c10/e2e_file_2gpus.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/c10/bench_n$N.json 2> gpurun_out/c10/bench_n$N.err; echo "bench $N rc=$?"; tail -c 300 gpurun_out/c10/bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/c10/bench_n4.json 2> gpurun_out/c10/bench_n4.err; echo "bench 4 rc=$?"
timeout 300 python bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/c10/bench_ref_n8.json 2>&1
for f in gpurun_out/c10/bench_n*.json; do python - $f <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][0])
print(sys.argv[1], 'value %.1f G' % (d['value']/1e9), 'e2e %.1f G' % (d['e2e']['value']/1e9), 'frac_copy', d['e2e'].get('frac_of_copies_in_then_out'), 'native_weak %.1f G e2e %.1f G' % (d['native_weak']['value']/1e9, d['native_weak']['e2e']['value']/1e9), 'inproc %.1f G biteq %s' % (d['inproc']['value']/1e9, d['inproc']['bit_equal_to_single_device']), 'e2e_file', (d.get('e2e_file') or {}).get('ours'))
PY
done
