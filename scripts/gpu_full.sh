mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/test_gpu_all.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/test_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench.log') if l.startswith('{')][-1])
print(round(d['value'],1), 'frames/s', round(d['ms_per_step'],3), 'ms/step; e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'lat', round(d['config']['ms_per_step_single_in_flight'],3), d['clocks'])
for k in d['kernels']: print(' ', k['name'][:40], round(k['ms'],4), round(k['gbs'],1), 'GB/s', {a:round(b,3) for a,b in k.items() if a in ('tflops','us_per_iter','hbm_frac')})
print(d['roofline']); print(d['cpu_baseline'])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
