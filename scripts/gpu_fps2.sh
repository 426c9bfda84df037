set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fps_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/test_fps_gpu.log 2>&1; echo "exit $?"
tail -15 gpurun_out/test_fps_gpu.log
timeout 900 python scripts/fps_sweep.py > gpurun_out/fps_sweep2.log 2>&1
grep -v "^{" gpurun_out/fps_sweep2.log | tail -5
