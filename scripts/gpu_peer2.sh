mkdir -p gpurun_out
N=${1:-2}
for mode in ${MODES:-peer nccl}; do
TSMDET_GATHER=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/bench_g${N}_$mode.log 2> gpurun_out/bench_g${N}_$mode.err; echo "bench N=$N $mode exit $?"
tail -4 gpurun_out/bench_g${N}_$mode.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_g${N}_$mode.log') if l.startswith('{')][-1])
    print('N=$N $mode', round(d['value'],1), 'frames/s', round(d['ms_per_step'],3), 'ms/step e2e', round(d['e2e']['value'],1), d['config'].get('gather'), d['clocks'])
except Exception as e: print('no bench', e)
PY
done
