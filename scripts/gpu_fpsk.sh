mkdir -p gpurun_out
if [ -z "$SKIP_CHECK" ]; then
timeout 240 python scripts/fpsb_k_check.py > gpurun_out/fpsk_check.log 2>&1; echo "exit $?"
grep -v "^checked" gpurun_out/fpsk_check.log | grep -v '"T"' | tail -70
fi
for l in scripts/_prof_libtsmdet*.so; do [ -f $l ] && echo $l && TSMDET_LIB=$PWD/$l timeout 120 python scripts/fpsb_prof_run.py 2>&1 | tee -a gpurun_out/fpsk_prof.log | tail -8; done
