#!/bin/bash
mkdir -p gpurun_out/golden
timeout 300 python tests/golden/make_golden.py --stack-only --out gpurun_out/golden > gpurun_out/r02d_golden.txt 2>&1
cp gpurun_out/golden/stack_ops.npz tests/golden/stack_ops.npz 2>/dev/null
timeout 900 python -m pytest tests/test_sa_module_gpu.py tests/test_stack_ops_gpu.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r02d_tests.txt
for op in mlp1 mlp2 mlp3; do timeout 120 python scripts/prof.py $op --time --reps 5 2>&1 | grep ms; done > gpurun_out/r02d_prof.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
tail -n 3 gpurun_out/r02d_golden.txt; tail -n 12 gpurun_out/r02d_tests.txt; cat gpurun_out/r02d_prof.txt; tail -n 5 gpurun_out/r02d_bench.err; head -c 300 gpurun_out/r02d_bench.json
