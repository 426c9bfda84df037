set -x
mkdir -p gpurun_out
timeout 900 python scripts/fps_sweep.py > gpurun_out/fps_sweep.log 2>&1
timeout 300 python scripts/fps_once.py > gpurun_out/fps_once.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fps_kernel -c 1 -o gpurun_out/fps_prof python scripts/fps_once.py > gpurun_out/ncu_fps.log 2>&1
tail -5 gpurun_out/ncu_fps.log
tail -60 gpurun_out/fps_sweep.log
