mkdir -p gpurun_out/c6
N=$(nvidia-smi -L | wc -l)
nvidia-smi --query-gpu=index,name,persistence_mode,power.limit --format=csv > gpurun_out/c6/box.txt 2>&1
(free -g; nproc; lscpu | grep -i "numa\|model name\|socket"; nvidia-smi topo -m | head -14) >> gpurun_out/c6/box.txt 2>&1
( time nvidia-smi -L ) >> gpurun_out/c6/box.txt 2>&1
p3arsec_b200/bin/cuinit_probe >> gpurun_out/c6/cuinit.jsonl 2>&1
CUDA_VISIBLE_DEVICES=0 p3arsec_b200/bin/cuinit_probe >> gpurun_out/c6/cuinit.jsonl 2>&1
p3arsec_b200/bin/cuinit_probe $N 1 >> gpurun_out/c6/cuinit.jsonl 2>&1
p3arsec_b200/bin/cuinit_probe 1 >> gpurun_out/c6/cuinit.jsonl 2>&1
CUDA_VISIBLE_DEVICES=0 p3arsec_b200/bin/cuinit_probe >> gpurun_out/c6/cuinit.jsonl 2>&1
CUDA_VISIBLE_DEVICES=0,1 p3arsec_b200/bin/cuinit_probe >> gpurun_out/c6/cuinit.jsonl 2>&1
CUDA_MODULE_LOADING=EAGER CUDA_VISIBLE_DEVICES=0 p3arsec_b200/bin/cuinit_probe >> gpurun_out/c6/cuinit.jsonl 2>&1
p3arsec_b200/bin/cuinit_probe >> gpurun_out/c6/cuinit.jsonl 2>&1
timeout 300 p3arsec_b200/bin/h2d_ceiling --reps 6 --kinds hostalloc,register,register_thp > gpurun_out/c6/h2d.jsonl 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -s -k "sharded or shards_over_devices or two_gpus or aos_sharded" > gpurun_out/c6/pytest_multi.log 2>&1; echo "pytest multi rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/c6/bench_n$N.json 2> gpurun_out/c6/bench_n$N.err; echo "bench rc=$?"
tail -c 400 gpurun_out/c6/bench_n$N.err
timeout 600 python tools/e2e_file_bench.py --gpus $N --reps 3 --ref-full > gpurun_out/c6/e2e_file_auto.json 2> gpurun_out/c6/e2e_auto.err
timeout 600 python tools/e2e_file_bench.py --gpus $N --reps 2 --devices all --ref-sample 300000 > gpurun_out/c6/e2e_file_all.json 2>> gpurun_out/c6/e2e_all.err
timeout 600 python tools/e2e_file_bench.py --gpus $N --reps 2 --devices all --fast-exit 0 --ref-sample 300000 > gpurun_out/c6/e2e_file_all_slowexit.json 2> gpurun_out/c6/e2e_all.err
head -30 gpurun_out/c6/box.txt; tail -3 gpurun_out/c6/pytest_multi.log; cat gpurun_out/c6/cuinit.jsonl; cat gpurun_out/c6/e2e_file_*.json | cut -c1-900
