N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
head -c 300 gpurun_out/r2_bench_n$N.json; echo
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n$N.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('train',{}).get('ms_per_step'), d.get('train',{}).get('value'), d.get('train',{}).get('allreduce_ms'))
PY
