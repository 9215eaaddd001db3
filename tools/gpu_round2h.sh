mkdir -p gpurun_out/r2h
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r2h/pytest.log
python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/r2h/bench.json 2> gpurun_out/r2h/bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2h/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h/bench.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'perframe',d['e2e']['per_frame_call']['value'],'frac',d['roofline']['frac'],'dropin',d.get('dropin'))
PY
K="--steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-dropin --no-parity"
GOFPID_NO_COND=1 python bench.py --workload 1080p_knn_many_blobs $K > gpurun_out/r2h/plain_knn.log 2>&1 &&
GOFPID_NO_COND=1 ncu --set full --clock-control none --import-source on -k regex:knn_gray -s 60 -c 1 -o gpurun_out/r2h/prof_knn_gray python bench.py --workload 1080p_knn_many_blobs $K > gpurun_out/r2h/ncu_knn.log 2>&1
echo "ncu knn rc=$?"
GOFPID_NO_COND=1 python bench.py --workload ilids_720x576_bgr_defaults --streams 32 $K > gpurun_out/r2h/plain_ilids.log 2>&1 &&
GOFPID_NO_COND=1 ncu --set full --clock-control none --import-source on -k regex:mog2_lean -s 12 -c 1 -o gpurun_out/r2h/prof_mog2_lean_bgr python bench.py --workload ilids_720x576_bgr_defaults --streams 32 $K > gpurun_out/r2h/ncu_ilids.log 2>&1
echo "ncu ilids rc=$?"
