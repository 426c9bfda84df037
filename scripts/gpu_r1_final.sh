# launch list of one step, ncu --set full of the multi-pick FPS kernel, pipeline depth sweep (each only after its own command ran clean)
mkdir -p gpurun_out
export TSMDET_FPS_ALGO=bucket
REPS=3 timeout 300 python scripts/one_step.py > gpurun_out/one_step.log 2>&1 && \
REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v4.csv python scripts/one_step.py > gpurun_out/ncu_list.log 2>&1
echo "launch list $?"
M=4096 timeout 300 python scripts/fps_once.py > gpurun_out/fpsb_once.log 2>&1 && \
M=4096 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_bucket -s 1 -c 1 -f -o gpurun_out/r01_fpsb_k8 python scripts/fps_once.py > gpurun_out/ncu_fpsb_k8.log 2>&1
echo "ncu fpsb $?"
unset TSMDET_FPS_ALGO
for dpt in ${DEPTHS:-6 8 10 12}; do
timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --depth $dpt > gpurun_out/bench_d$dpt.log 2> gpurun_out/bench_d$dpt.err; echo "bench $dpt $?"
done
python - <<'PY'
import json
for dpt in (6, 8, 10, 12):
  try:
    d=json.loads(open(f'gpurun_out/bench_d{dpt}.log').read().strip().splitlines()[-1])
    print(dpt, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['config']['ms_per_step_single_in_flight'], d['clocks'])
  except Exception as e: print('no bench', dpt, e)
PY
