mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_query_group_gpu.py tests/test_sa_module_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -4
timeout 500 python scripts/group_bw.py 2>&1 | tail -12 | tee gpurun_out/group_bw.log
