mkdir -p gpurun_out
timeout 600 python scripts/concurrency_probe.py > gpurun_out/probe.log 2>&1; cat gpurun_out/probe.log | tail -12
