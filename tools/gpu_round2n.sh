mkdir -p gpurun_out/r2n
for NP in 0 1; do echo "== clip 72 frames, GOFPID_NO_PREFETCH=$NP"; GOFPID_NO_PREFETCH=$NP FLOOR_NCLIP=72 timeout 300 python tools/stage_floor.py 3 2>&1 | tail -3; done > gpurun_out/r2n/floor_prefetch.txt 2>&1
cat gpurun_out/r2n/floor_prefetch.txt
python bench.py --steps 20 --warmup 5 --no-configs --no-dropin --no-cpu-baseline > gpurun_out/r2n/bench.json 2> gpurun_out/r2n/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n/bench.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'us/frame',d['roofline']['kernel_us_per_frame'])
PY
