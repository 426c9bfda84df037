mkdir -p gpurun_out
N=${1:-2}
run() {
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 6 --no-cpu-baseline > gpurun_out/bench_gx.log 2> gpurun_out/bench_gx.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_gx.log') if l.startswith('{')][-1])
    print('$1:', round(d['value'],1), 'frames/s', round(d['ms_per_step'],3), 'ms/step e2e', round(d['e2e']['value'],1))
except Exception as e: print('no bench $1', e)
PY
}
run base
TSMDET_GATHER_ASYNC=1 run async
NCCL_MAX_NCHANNELS=2 run nch2
TORCH_NCCL_HIGH_PRIORITY=1 run hiprio
NCCL_MAX_NCHANNELS=2 TORCH_NCCL_HIGH_PRIORITY=1 run nch2_hiprio
