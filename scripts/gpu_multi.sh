mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_g$N.log 2> gpurun_out/bench_g$N.err; echo "bench N=$N exit $?"
tail -5 gpurun_out/bench_g$N.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_g$N.log') if l.startswith('{')][-1])
    print('N=$N', round(d['value'],1), 'frames/s', round(d['ms_per_step'],3), 'ms/step e2e', round(d['e2e']['value'],1), d['n_gpus'], d['scaling'], d['clocks'])
except Exception as e: print('no bench', e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_g$N.log 2> gpurun_out/bench_ref_g$N.err; echo "ref exit $?"; cat gpurun_out/bench_ref_g$N.log | cut -c1-300
