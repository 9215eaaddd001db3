mkdir -p gpurun_out/r2p
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -x -q -k "mog2 or headline or 4k or ilids or bgr or dropin or clip" > gpurun_out/r2p/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r2p/pytest.log
for NC in 24 72; do echo "== clip $NC frames"; FLOOR_NCLIP=$NC timeout 300 python tools/stage_floor.py 3 2>&1 | tail -3; done > gpurun_out/r2p/floor.txt 2>&1
cat gpurun_out/r2p/floor.txt
python bench.py --steps 20 --warmup 5 --no-dropin --no-cpu-baseline > gpurun_out/r2p/bench.json 2> gpurun_out/r2p/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p/bench.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'us/frame',d['roofline']['kernel_us_per_frame'])
for k,v in d['configs'].items():
    r=v['roofline']; print('%-34s value %8.0f e2e %8.0f frac %.3f kernel_us/frame %.2f'%(k,v['value'],v['e2e']['value'],r['frac'],r['kernel_us_per_frame']))
PY
