set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sa_module_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/test_sa_module_gpu.log 2>&1; echo "exit $?"
tail -40 gpurun_out/test_sa_module_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.log 2> gpurun_out/bench_bf16.err; echo "bench $?"
tail -5 gpurun_out/bench_bf16.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_bf16.log').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['dtype'], d['e2e'])
    for k in d['kernels']: print(k['name'], round(k['ms'],4), {a:b for a,b in k.items() if a in ('tflops','us_per_iter')})
except Exception as e: print('no bench', e)
PY
