set -x
mkdir -p gpurun_out
echo skip-tests
for dpt in ${DEPTHS:-1 2 3}; do
timeout 600 python bench.py --steps 30 --warmup 6 --precision bf16 --no-cpu-baseline --depth $dpt > gpurun_out/bench_d$dpt.log 2> gpurun_out/bench_d$dpt.err; echo "bench $dpt $?"
tail -3 gpurun_out/bench_d$dpt.err
done
python - <<'PY'
import json
import os
for dpt in [int(x) for x in os.environ.get("DEPTHS","1 2 3").split()]:
  try:
    d=json.loads(open(f'gpurun_out/bench_d{dpt}.log').read().strip().splitlines()[-1])
    print(dpt, round(d['value'],1), round(d['ms_per_step'],3), d['e2e']['value'], d['gpu_launches'], d['config']['ms_per_step_single_in_flight'], d['clocks'])
    if dpt==1:
      for k in d['kernels']: print(k['name'], round(k['ms'],4), {a:b for a,b in k.items() if a in ('tflops','us_per_iter','tests_per_s')})
  except Exception as e: print('no bench', dpt, e)
PY
