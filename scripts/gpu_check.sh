set -x
mkdir -p gpurun_out/golden
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/nproc.txt
timeout 600 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for f in test_fps_gpu test_query_group_gpu test_iou_nms_gpu test_sa_module_gpu; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "exit $f $?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/$f.log >> gpurun_out/summary.txt
done
timeout 600 python tests/golden/make_golden.py --out gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden $?" >> gpurun_out/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --precision fp32 --profile-kernels > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
