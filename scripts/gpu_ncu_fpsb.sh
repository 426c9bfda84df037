set -x
mkdir -p gpurun_out
export TSMDET_FPS_ALGO=bucket
M=2048 timeout 300 python scripts/fps_once.py > gpurun_out/fpsb_once.log 2>&1 && \
M=2048 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fps_bucket -c 1 -o gpurun_out/fpsb_prof python scripts/fps_once.py > gpurun_out/ncu_fpsb.log 2>&1
tail -3 gpurun_out/ncu_fpsb.log
