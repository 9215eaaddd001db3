mkdir -p gpurun_out/r3c
timeout 900 python -m pytest tests -m gpu -x -q -k "knn" > gpurun_out/r3c/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r3c/pytest.log
K="--steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-dropin"
python bench.py --workload 1080p_knn_many_blobs $K > gpurun_out/r3c/bench_knn.json 2> gpurun_out/r3c/bench_knn.err; echo "knn rc=$?"
python - <<'PY'
import json
for n in ("bench_knn",):
    d=json.loads(open("gpurun_out/r3c/%s.json"%n).read().strip().splitlines()[-1])
    print(n, "value %.0f e2e %.0f frac %.3f kernel_us/frame %.2f parity %s"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_us_per_frame"], d.get("parity_checked")))
PY
