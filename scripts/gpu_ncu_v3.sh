mkdir -p gpurun_out
export TSMDET_FPS_ALGO=bucket
REPS=3 timeout 300 python scripts/one_step.py > gpurun_out/one_step.log 2>&1 || exit 1
REPS=3 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v3.csv python scripts/one_step.py > gpurun_out/ncu_list.log 2>&1; echo "list $?"
export REPS=1
for spec in "nmslazy:nms_lazy_kernel:1" "mlp3:sa_mlp_tc_kernel:3" "mlp1:sa_mlp_tc_kernel:1"; do
  IFS=: read name pat skip <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s $((skip-1)) -c 1 -f -o gpurun_out/r01_$name python scripts/one_step.py > gpurun_out/ncu_$name.log 2>&1
  echo "$name exit $?"
done
