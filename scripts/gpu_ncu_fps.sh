set -x
mkdir -p gpurun_out
M=1024 timeout 300 python scripts/fps_once.py > gpurun_out/fps_once.log 2>&1 && \
M=1024 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fps_kernel -c 1 -o gpurun_out/fps_prof2 python scripts/fps_once.py > gpurun_out/ncu_fps2.log 2>&1
tail -3 gpurun_out/ncu_fps2.log
