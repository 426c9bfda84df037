#!/bin/bash
# One entry point for everything that runs on the GPU box (replaces the one-off drivers of round 1):
#   scripts/gpu.sh [tests[:<pytest args>]] [smoke] [bench[:<bench args>]] [benchn:<gpus>] [prof:<op>] [ncu:<op>:<kernel regex>] [golden] [launches]
# e.g.  gpurun --timeout 1800 -- 'bash scripts/gpu.sh tests smoke bench'
# Every step writes gpurun_out/<tag>_<step>.txt (TAG env, default r02); steps run in order, a failing step does not
# stop the later ones.  ncu steps follow B200_PROFILING.md: the plain command must exit 0 first.
mkdir -p gpurun_out
TAG=${TAG:-r02}
for step in "$@"; do
  kind=${step%%:*}; rest=${step#*:}; [ "$rest" = "$step" ] && rest=""
  case $kind in
    tests)   timeout 2400 python -m pytest tests -m gpu -q $rest 2>&1 | tail -n 60 > gpurun_out/${TAG}_tests.txt; tail -n 4 gpurun_out/${TAG}_tests.txt ;;
    smoke)   timeout 600 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.txt 2>&1; tail -n 3 gpurun_out/${TAG}_smoke.txt ;;
    bench)   timeout 1500 python bench.py $rest > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -n 3 gpurun_out/${TAG}_bench.err; head -c 400 gpurun_out/${TAG}_bench.json; echo ;;
    benchn)  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $rest --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $rest > gpurun_out/${TAG}_bench_n$rest.json 2> gpurun_out/${TAG}_bench_n$rest.err; tail -n 3 gpurun_out/${TAG}_bench_n$rest.err; head -c 300 gpurun_out/${TAG}_bench_n$rest.json; echo ;;
    prof)    timeout 300 python scripts/prof.py $rest --time --reps 5 2>&1 | grep -E "ms:|Error" | tee -a gpurun_out/${TAG}_prof.txt ;;
    ncu)     op=${rest%%:*}; k=${rest#*:}; skip=${SKIP:-1}   # mlpN: the set-up runs the 3-layer backbone first -> SKIP=3
             timeout 300 python scripts/prof.py $op --time > gpurun_out/${TAG}_prof_$op.log 2>&1 && \
             timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/${TAG}_ncu_$op python scripts/prof.py $op > gpurun_out/${TAG}_ncu_$op.log 2>&1
             tail -n 2 gpurun_out/${TAG}_ncu_$op.log ;;
    dur)     timeout 300 python scripts/prof.py $rest > /dev/null 2>&1 && \
             timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_dur_$rest.csv python scripts/prof.py $rest > /dev/null 2>&1
             grep -E "mlp|transpose|pack|fps|bq_|nms|voxel|stack" gpurun_out/${TAG}_dur_$rest.csv | awk -F'","' '{print $5, $NF}' | tail -n 12 ;;
    launches) timeout 300 python scripts/prof.py step > gpurun_out/${TAG}_step.log 2>&1 && \
             timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/prof.py step > gpurun_out/${TAG}_launches.log 2>&1
             tail -n 2 gpurun_out/${TAG}_launches.log ;;
    golden)  mkdir -p gpurun_out/golden; timeout 600 python tests/golden/make_golden.py --out gpurun_out/golden $rest > gpurun_out/${TAG}_golden.txt 2>&1; tail -n 2 gpurun_out/${TAG}_golden.txt ;;
    *) echo "unknown step $step" ;;
  esac
done
