mkdir -p gpurun_out
timeout 900 python scripts/fps_bucket_check.py > gpurun_out/fpsb_check.log 2>&1; echo "exit $?"
grep -v "^checked" gpurun_out/fpsb_check.log | tail -60
