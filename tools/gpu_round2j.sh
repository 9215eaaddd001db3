mkdir -p gpurun_out/r2j
timeout 900 python -m pytest tests -m gpu -x -q -k "knn or KNN or dropin or gray_conversion" > gpurun_out/r2j/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r2j/pytest.log
python bench.py --workload 1080p_knn_many_blobs --steps 6 --warmup 3 --no-configs --no-dropin > gpurun_out/r2j/bench_knn.json 2> gpurun_out/r2j/bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2j/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j/bench_knn.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'us/frame',d['roofline']['kernel_us_per_frame'])
PY
