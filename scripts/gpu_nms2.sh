mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_iou_nms_gpu.py tests/test_sa_module_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/test_iou_nms_gpu.log 2>&1; echo "exit $?"; tail -5 gpurun_out/test_iou_nms_gpu.log
for dpt in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 6 --precision bf16 --no-cpu-baseline --depth $dpt > gpurun_out/bench_d$dpt.log 2> gpurun_out/bench_d$dpt.err; echo "bench $dpt $?"; tail -3 gpurun_out/bench_d$dpt.err
done
python - <<'PY'
import json
for dpt in (1,2):
  try:
    d=json.loads(open(f'gpurun_out/bench_d{dpt}.log').read().strip().splitlines()[-1])
    print(dpt, round(d['value'],1), round(d['ms_per_step'],3), d['e2e']['value'], d['gpu_launches'], d['config']['ms_per_step_single_in_flight'])
    if dpt==1:
      for k in d['kernels']: print(k['name'], round(k['ms'],4), {a:b for a,b in k.items() if a in ('tflops','us_per_iter','tests_per_s')})
  except Exception as e: print('no bench', dpt, e)
PY
