set -x
mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -x -q -k "mog2 or clip or headline or streams or 1080p or 4k or ilids" > gpurun_out/r2b/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 6 gpurun_out/r2b/pytest.log
timeout 300 python tools/stage_floor.py 0,3,4 > gpurun_out/r2b/stage_floor.txt 2>&1
cat gpurun_out/r2b/stage_floor.txt
GOFPID_MOG2_GENERIC=1 timeout 300 python tools/stage_floor.py 3 > gpurun_out/r2b/stage_floor_generic.txt 2>&1
cat gpurun_out/r2b/stage_floor_generic.txt
for v in c9 m3; do GOFPID_LIB=libgofpid_b200_$v.so timeout 300 python tools/stage_floor.py 3 > gpurun_out/r2b/stage_floor_$v.txt 2>&1; cat gpurun_out/r2b/stage_floor_$v.txt; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-configs --no-dropin > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2b/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b/bench.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'kernel_us_per_frame',d['roofline']['kernel_us_per_frame'])
PY
