mkdir -p gpurun_out
nproc; grep -c processor /proc/cpuinfo
run() {
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/bench_g8x.log 2> gpurun_out/bench_g8x.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_g8x.log') if l.startswith('{')][-1])
    print('$1:', round(d['value'],1), 'frames/s', round(d['ms_per_step'],3), 'ms/step e2e', round(d['e2e']['value'],1), d['clocks'])
except Exception as e: print('no bench $1', e)
PY
}
run gather
TSMDET_BENCH_NO_GATHER=1 run nogather
NCCL_MAX_NCHANNELS=2 run gather_nch2
