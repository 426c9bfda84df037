#!/bin/bash
# round 2, call A: new MLP kernels + staged host path + drop-in proof, then the whole suite, smoke and the bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_sa_module_gpu.py -m gpu -q -x -s 2>&1 | tail -120 > gpurun_out/r02a_test_sa.txt
timeout 600 python -m pytest tests/test_dropin_reference_py_gpu.py -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r02a_test_dropin.txt
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_sa_module_gpu.py --deselect tests/test_dropin_reference_py_gpu.py 2>&1 | tail -60 > gpurun_out/r02a_test_rest.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02a_smoke.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -5 gpurun_out/r02a_test_sa.txt gpurun_out/r02a_test_dropin.txt gpurun_out/r02a_test_rest.txt gpurun_out/r02a_smoke.txt
head -c 1500 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
