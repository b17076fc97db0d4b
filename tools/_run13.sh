set -x
mkdir -p gpurun_out
P2M_TRACE=1 python -m pose2mesh_release_b200.build --force > /dev/null 2>&1
python tools/umma_trace.py 128 128 256 0 > gpurun_out/r2_trace_l17.txt 2>&1
head -70 gpurun_out/r2_trace_l17.txt
