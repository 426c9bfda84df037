for c in 8 32; do
export CUDA_DEVICE_MAX_CONNECTIONS=$c
timeout 300 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('conn=$c N1', round(d['value']), round(d['e2e']['value']), d['ms_per_step'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('conn=$c N2', round(d['value']), round(d['e2e']['value']), d['ms_per_step'])"
done
