set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
python -m pytest tests/test_gpu_at_size.py -m gpu -q -k data_parallel 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -5 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','e2e')})
print(d['train'])
PY
