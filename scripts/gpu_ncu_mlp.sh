mkdir -p gpurun_out
timeout 300 python scripts/mlp_once.py > gpurun_out/mlp_once.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sa_mlp_tc -c 1 -o gpurun_out/mlp_prof python scripts/mlp_once.py > gpurun_out/ncu_mlp.log 2>&1
tail -2 gpurun_out/ncu_mlp.log
