mkdir -p gpurun_out/r3b
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -x -q -k "morph or headline or 4k or fd or blobs or dropin or clip" > gpurun_out/r3b/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r3b/pytest.log
for NC in 24 72; do echo "== clip $NC frames"; FLOOR_NCLIP=$NC timeout 300 python tools/stage_floor.py 3 2>&1 | tail -3; done
K="--steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-dropin"
python bench.py $K > gpurun_out/r3b/bench.json 2> gpurun_out/r3b/bench.err; echo "headline rc=$?"
python - <<'PY'
import json
for n in ("bench",):
    d=json.loads(open("gpurun_out/r3b/%s.json"%n).read().strip().splitlines()[-1])
    print(n, "value %.0f e2e %.0f frac %.3f kernel_us/frame %.2f parity %s"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_us_per_frame"], d.get("parity_checked")))
PY
