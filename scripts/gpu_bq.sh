set -x
mkdir -p gpurun_out
for f in test_query_group_gpu test_sa_module_gpu; do
timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/$f.log 2>&1; echo "exit $f $?"; tail -4 gpurun_out/$f.log; done
timeout 600 python bench.py --steps 20 --warmup 5 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bq.log 2> gpurun_out/bench_bq.err; echo "bench $?"
tail -5 gpurun_out/bench_bq.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_bq.log').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['dtype'], d['e2e'], d['gpu_launches'])
    for k in d['kernels']: print(k['name'], round(k['ms'],4), {a:b for a,b in k.items() if a in ('tflops','us_per_iter','tests_per_s')})
except Exception as e: print('no bench', e)
PY
