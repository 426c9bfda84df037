mkdir -p gpurun_out
export TSMDET_FPS_ALGO=bucket REPS=1
timeout 300 python scripts/one_step.py > gpurun_out/one_step.log 2>&1 || exit 1
for spec in "fpsb:regex:fps_bucket_kernel:1" "bqbuild:regex:bq_grid_build:1" "bqquery:regex:bq_grid_query:1" "mlp:regex:sa_mlp_tc_kernel:3" "nmspairs:regex:nms_grid_pairs:1" "nmssweep:regex:nms_sweep:1"; do
  IFS=: read name kind pat skip <<< "$spec"
  # -s: skip the warm-up launches of that kernel inside one_step (REPS=1 -> the first launch per layer is already warm enough for ncu's own replay)
  timeout 600 ncu --set full --clock-control none --import-source on -k $kind:$pat -s $((skip-1)) -c 1 -f -o gpurun_out/r01_$name python scripts/one_step.py > gpurun_out/ncu_$name.log 2>&1
  echo "$name exit $?"
done
ls -la gpurun_out/*.ncu-rep
