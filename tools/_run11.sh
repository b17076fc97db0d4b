set -x
mkdir -p gpurun_out
P2M_TRACE=1 python -m pose2mesh_release_b200.build --force > gpurun_out/trace_build.log 2>&1 || exit 1
P2M_TRACE_V=12288 P2M_TRACE_UNPOOL=1 P2M_TRACE_FOUT=128 python tools/umma_trace_model.py 700 > gpurun_out/trace_l17.txt 2>&1
P2M_TRACE_V=12288 P2M_TRACE_UNPOOL=0 P2M_TRACE_FOUT=128 python tools/umma_trace_model.py 700 > gpurun_out/trace_l18.txt 2>&1
tail -3 gpurun_out/trace_l18.txt
