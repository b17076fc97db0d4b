set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_cheb_conv_umma<256' -s 16 -c 1 -o gpurun_out/r2_l11 -f python tools/ncu_forward.py 2 > gpurun_out/r2_ncu4.log 2>&1
tail -2 gpurun_out/r2_ncu4.log
ncu --set full --clock-control none --import-source on -k regex:'k_cheb_dw_umma' -s 1 -c 1 -o gpurun_out/r2_dw -f python tools/ncu_forward.py 1 256 train > gpurun_out/r2_ncu5.log 2>&1
tail -2 gpurun_out/r2_ncu5.log
