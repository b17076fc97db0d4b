mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --train-steps 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
print(d['roofline']['frac'], d['roofline']['ms_per_launch'])
print(d['train']['ms_per_step'])
print([ (l['layer'],l['ms']) for l in d['layers']])
PY
