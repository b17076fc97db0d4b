set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:k_cheb_(conv_umma|t1)' -s 77 -c 6 -o gpurun_out/r2_l17b -f python tools/ncu_forward.py 2 > gpurun_out/r2_ncu2.log 2>&1
tail -2 gpurun_out/r2_ncu2.log
