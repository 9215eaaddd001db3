mkdir -p gpurun_out/r2s
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r2s/pytest.log
python bench.py --steps 20 --warmup 5 --no-dropin > gpurun_out/r2s/bench.json 2> gpurun_out/r2s/bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2s/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s/bench.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'floor',d['e2e']['host_link_floor']['value'],'frac',d['roofline']['frac'],'us/frame',d['roofline']['kernel_us_per_frame'])
for k,v in d['configs'].items():
    r=v['roofline']; print('%-34s value %8.0f e2e %8.0f floor %8.0f frac %.3f kernel_us/frame %.2f'%(k,v['value'],v['e2e']['value'],v['e2e']['host_link_floor']['value'],r['frac'],r['kernel_us_per_frame']))
PY
