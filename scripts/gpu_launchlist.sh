set -x
mkdir -p gpurun_out
REPS=3 timeout 300 python scripts/one_step.py > gpurun_out/one_step.log 2>&1 && \
REPS=3 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/one_step.py > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
