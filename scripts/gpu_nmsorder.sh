mkdir -p gpurun_out
for v in 0 1; do for dpt in 1 2; do
TSMDET_NMS_AFTER_FPS=$v timeout 600 python bench.py --steps 30 --warmup 6 --precision bf16 --no-cpu-baseline --depth $dpt > gpurun_out/bench_n${v}_d$dpt.log 2> gpurun_out/bench_n${v}_d$dpt.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n${v}_d$dpt.log').read().strip().splitlines()[-1])
print('nms_after_fps=$v depth=$dpt', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['config']['ms_per_step_single_in_flight'])
PY
done; done
