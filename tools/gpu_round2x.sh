mkdir -p gpurun_out/r2x
for V in "" "GOFPID_COPIER_SPIN_US=0" "GOFPID_COPIER_THREADS=7" "GOFPID_COPIER_THREADS=1" "GOFPID_COPIER_THREADS=0"; do
  echo "== $V"; env $V timeout 300 python tools/latency_breakdown.py 2>&1 | grep -E "pageable|pinned frame"
  env $V timeout 300 python tools/dropin_rate.py 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2x/latency.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "dropin or host or clip or pageable or pinned" > gpurun_out/r2x/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2x/pytest.log
