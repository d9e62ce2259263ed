#!/bin/bash
# run_all_configs.sh [outdir] -- one bench line per BASELINE.json config that fits ONE GPU, plus the sibling Map, each with
# its reference arm beside it.  Run on a GPU box:   gpurun --timeout 1500 -- 'bash tools/run_all_configs.sh gpurun_out/cfg'
# and copy the JSON lines into profiles/r02_bench_*.json.  (Config 4 at 2/4/8 GPUs and config 5 at 8 GPUs need a
# multi-GPU box: `torchrun ... bench.py --gpus N` and `tools/e2e_file_bench.py --gpus 8`.)
set -u
out=${1:-gpurun_out/cfg}
mkdir -p "$out"
S=${STEPS:-20}; W=${WARMUP:-5}
run() { name=$1; shift; echo "== $name: $*" >&2; timeout 900 "$@" > "$out/$name.json" 2> "$out/$name.err" || echo "$name rc=$?" >&2; tail -c 300 "$out/$name.json" >&2; echo >&2; }
# config 1: simsmall vs the FastFlow build on the host cores
run r02_bench_simsmall_n1            python bench.py --workload simsmall --steps $S --warmup $W
run r02_bench_reference_simsmall     python bench.py --workload simsmall --steps $S --warmup $W --impl reference
# config 2: native fp32 (the headline line: roofline with probe + live ncu traffic, e2e, e2e_run_order, strong_1b anchor)
run r02_bench_native_n1              python bench.py --steps $S --warmup $W
run r02_bench_reference_native       python bench.py --steps $S --warmup $W --impl reference
# config 3: native with fptype=double
run r02_bench_native_fp64_n1         python bench.py --workload native_fp64 --steps $S --warmup $W
run r02_bench_reference_native_fp64  python bench.py --workload native_fp64 --steps 3 --warmup 1 --impl reference --no-full-size
# config 4 at N = 1: the 1B-option set on one GPU (device-resident)
run r02_bench_synth1b_n1             python bench.py --workload synth1b --steps 5 --warmup 3 --headline-only
# config 5 at N = 1: end-to-end file -> prices file, NUM_RUNS = 1
run r02_e2e_file_native_n1           python tools/e2e_file_bench.py --gpus 1 --reps 3
# the sibling Map
run r02_sw_bench_native              python bench.py --workload swaptions_native --steps $S --warmup $W
run r02_sw_bench_ref                 python bench.py --workload swaptions_native --steps 2 --warmup 1 --impl reference
run r02_sw_bench_simsmall            python bench.py --workload swaptions_simsmall --steps $S --warmup $W
