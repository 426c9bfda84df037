mkdir -p gpurun_out
timeout 900 python scripts/bq_grid_check.py > gpurun_out/bqg_check.log 2>&1; echo "exit $?"
tail -30 gpurun_out/bqg_check.log
