mkdir -p gpurun_out/c32
O=gpurun_out/c32
python tools/e2e_file_bench.py --gpus 1 --reps 5 > $O/e2e_file_new.json 2> $O/e2e_file_new.err; tail -2 $O/e2e_file_new.err
python - $O/e2e_file_new.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(d['ours'], d.get('ours_all_reps'))
PY
