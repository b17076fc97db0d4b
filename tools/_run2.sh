set -x
mkdir -p gpurun_out
free -g | head -2; nproc
( time python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -40 gpurun_out/r2_pytest_gpu.log
ncu --set full --clock-control none --import-source on -k regex:'k_cheb_(conv_umma|t1)' -s 77 -c 6 -o gpurun_out/r2_base_l17 -f python tools/ncu_forward.py 2 > gpurun_out/r2_ncu2.log 2>&1
tail -2 gpurun_out/r2_ncu2.log
