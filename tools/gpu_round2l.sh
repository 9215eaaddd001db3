mkdir -p gpurun_out/r2l
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -x -q -k "mog2 and not generic or headline or 4k" > gpurun_out/r2l/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r2l/pytest.log
python bench.py --steps 20 --warmup 5 --no-configs --no-dropin --no-cpu-baseline > gpurun_out/r2l/bench.json 2> gpurun_out/r2l/bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2l/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l/bench.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'floor',d['e2e']['host_link_floor']['value'],'frac',d['roofline']['frac'],'us/frame',d['roofline']['kernel_us_per_frame'])
PY
