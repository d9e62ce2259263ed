set -x
mkdir -p gpurun_out/c1
nvidia-smi --query-gpu=name,persistence_mode,power.limit,clocks.max.sm --format=csv > gpurun_out/c1/box.txt 2>&1
free -g >> gpurun_out/c1/box.txt; nproc >> gpurun_out/c1/box.txt; cat /sys/kernel/mm/transparent_hugepage/enabled >> gpurun_out/c1/box.txt
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/c1/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1/box.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c1/bench_n1.json 2> gpurun_out/c1/bench_n1.err; echo "bench rc=$?" >> gpurun_out/c1/box.txt
timeout 600 python tools/fp32_adversarial.py --random-millions 200 --out gpurun_out/c1/r02_fp32_adversarial.json > /dev/null 2> gpurun_out/c1/adv.err; echo "adv rc=$?" >> gpurun_out/c1/box.txt
timeout 300 p3arsec_b200/bin/h2d_ceiling --gpus 1 > gpurun_out/c1/h2d_n1.jsonl 2>&1
for i in 1 2; do p3arsec_b200/bin/cuinit_probe >> gpurun_out/c1/cuinit.jsonl 2>&1; CUDA_MODULE_LOADING=EAGER p3arsec_b200/bin/cuinit_probe >> gpurun_out/c1/cuinit.jsonl 2>&1; done
tail -5 gpurun_out/c1/pytest.log; cat gpurun_out/c1/box.txt; tail -3 gpurun_out/c1/adv.err
