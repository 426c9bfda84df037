mkdir -p gpurun_out
N=${1:-2}
for cfg in "6 0" "6 1" "2 0" "1 0"; do
set -- $cfg
TSMDET_BENCH_NO_GATHER=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --depth $1 --no-cpu-baseline > gpurun_out/bench_gx.log 2> gpurun_out/bench_gx.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_gx.log') if l.startswith('{')][-1])
    print('depth $1 nogather $2:', round(d['value'],1), 'frames/s', round(d['ms_per_step'],3), 'ms/step e2e', round(d['e2e']['value'],1))
except Exception as e: print('no bench', e)
PY
done
