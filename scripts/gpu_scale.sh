mkdir -p gpurun_out
for N in "$@"; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/bench_g$N.log 2> gpurun_out/bench_g$N.err; echo "bench N=$N exit $?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_g$N.log') if l.startswith('{')][-1])
    print('N=$N', round(d['value'],1), 'frames/s', round(d['ms_per_step'],3), 'ms/step e2e', round(d['e2e']['value'],1), d['clocks'])
except Exception as e: print('no bench', e)
PY
done
