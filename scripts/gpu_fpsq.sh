mkdir -p gpurun_out
QUICK=1 timeout 600 python scripts/fps_sweep.py > gpurun_out/fps_quick.log 2>&1; cat gpurun_out/fps_quick.log
