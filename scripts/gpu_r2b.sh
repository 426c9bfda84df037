#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dropin_reference_py_gpu.py tests/test_voxel_centroid_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r02b_tests.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
tail -n 8 gpurun_out/r02b_tests.txt; tail -n 5 gpurun_out/r02b_bench.err; head -c 800 gpurun_out/r02b_bench.json
