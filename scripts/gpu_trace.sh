mkdir -p gpurun_out
NO_NMS=1 timeout 300 python scripts/trace_step.py > gpurun_out/trace.log 2>&1; cat gpurun_out/trace.log | tail -28
