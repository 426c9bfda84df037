# usage: bash scripts/gpu_iter.sh "<test files>" : run the given gpu test files, then bench at depth 1 and 2
mkdir -p gpurun_out
for f in $1; do
timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/$f.log 2>&1; echo "exit $f $?"; tail -4 gpurun_out/$f.log; done
for dpt in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 6 --precision bf16 --no-cpu-baseline --depth $dpt > gpurun_out/bench_d$dpt.log 2> gpurun_out/bench_d$dpt.err; echo "bench $dpt $?"; tail -3 gpurun_out/bench_d$dpt.err
done
python - <<'PY'
import json
for dpt in (1,2):
  try:
    d=json.loads(open(f'gpurun_out/bench_d{dpt}.log').read().strip().splitlines()[-1])
    print(dpt, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['gpu_launches'], round(d['config']['ms_per_step_single_in_flight'],3))
    if dpt==1:
      for k in d['kernels']: print(k['name'], round(k['ms'],4), {a:b for a,b in k.items() if a in ('tflops','us_per_iter','tests_per_s')})
  except Exception as e: print('no bench', dpt, e)
PY
timeout 300 python scripts/trace_step.py > gpurun_out/trace.log 2>&1; tail -13 gpurun_out/trace.log
