set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/r2_base_bench.json 2> gpurun_out/r2_base_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_base_launches_fwd.csv python tools/ncu_forward.py 2 > gpurun_out/r2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_cheb_(conv_umma|t1)' -s 69 -c 3 -o gpurun_out/r2_base_l17 -f python tools/ncu_forward.py 2 > gpurun_out/r2_ncu2.log 2>&1
tail -3 gpurun_out/r2_ncu2.log
cat gpurun_out/r2_base_bench.json | head -c 600
