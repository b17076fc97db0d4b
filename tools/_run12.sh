set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "backward or gradient or train_step or variants" 2>&1 | tail -3
python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r2_bench_train_d.json 2> gpurun_out/r2_bench_train_d.err
head -c 330 gpurun_out/r2_bench_train_d.json; echo
ncu --set full --clock-control none --import-source on -k regex:'k_cheb_conv_umma' -s 32 -c 1 -o gpurun_out/r2_l11 -f python tools/ncu_forward.py 2 > gpurun_out/r2_ncu4.log 2>&1
tail -2 gpurun_out/r2_ncu4.log
