mkdir -p gpurun_out/r2i
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r2i/pytest.log
python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/r2i/bench.json 2> gpurun_out/r2i/bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2i/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i/bench.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'perframe',d['e2e']['per_frame_call']['value'],'frac',d['roofline']['frac'],'us/frame',d['roofline']['kernel_us_per_frame'],'dropin',d.get('dropin'))
PY
timeout 200 python tools/stage_floor.py 3 2>&1 | tail -3
