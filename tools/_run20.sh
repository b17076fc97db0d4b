set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -3 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','e2e')})
print(d['train'])
PY
