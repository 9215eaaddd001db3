mkdir -p gpurun_out/r2z
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r2z/pytest.log
K="--steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-dropin"
python bench.py $K > gpurun_out/r2z/bench.json 2> gpurun_out/r2z/bench.err; echo "headline rc=$?"
python bench.py --workload 1080p_knn_many_blobs $K > gpurun_out/r2z/bench_knn.json 2> gpurun_out/r2z/bench_knn.err; echo "knn rc=$?"
python - <<'PY'
import json
for n in ("bench","bench_knn"):
    try:
        d=json.loads(open("gpurun_out/r2z/%s.json"%n).read().strip().splitlines()[-1])
        print(n, "value %.0f e2e %.0f frac %.3f kernel_us/frame %.2f parity %s"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_us_per_frame"], d.get("parity_checked")))
    except Exception as e: print(n, "failed", e)
PY
for NC in 24 72; do echo "== clip $NC frames"; FLOOR_NCLIP=$NC timeout 300 python tools/stage_floor.py 3 2>&1 | tail -3; done
