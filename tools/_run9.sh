set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r2_pytest_gpu.log 2>&1
grep -n '^E  \|^FAILED\|passed\|failed' gpurun_out/r2_pytest_gpu.log | head -40
python bench.py --steps 5 --warmup 3 --cpu-sample 0 --train-steps 3 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_d.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches_per_step')})
print(d['roofline']['frac'], d['roofline']['ms_per_launch'])
print(d['train']['ms_per_step'])
print([ (l['layer'],l['ms']) for l in d['layers']])
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train.csv python tools/ncu_forward.py 2 256 train > gpurun_out/r2_ncu3.log 2>&1
