set -x
mkdir -p gpurun_out
P2M_TRACE=1 python -m pose2mesh_release_b200.build --force > /dev/null 2>&1
python tools/umma_trace.py 128 128 256 0 > gpurun_out/r2_trace_l17b.txt 2>&1
sed -n 14,60p gpurun_out/r2_trace_l17b.txt
python -m pose2mesh_release_b200.build --force > /dev/null 2>&1
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu.log 2>&1
grep -n '^E  \|^FAILED\|passed\|failed' gpurun_out/r2_pytest_gpu.log | head -20
python bench.py --steps 5 --warmup 3 --cpu-sample 0 --train-steps 3 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_e.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches_per_step')})
print(d['roofline']['frac'], d['roofline']['ms_per_launch'])
print(d['train']['ms_per_step'])
print([ (l['layer'],l['ms']) for l in d['layers']])
PY
