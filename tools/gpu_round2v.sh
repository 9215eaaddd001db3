mkdir -p gpurun_out/r2v
timeout 1200 python -m pytest tests -m gpu -x -q -k "mog2 or bgr or ilids or colo or fullsize or clip" > gpurun_out/r2v/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r2v/pytest.log
K="--steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-dropin"
python bench.py --workload 1080p_bgr_mog2_open_close_blobs $K > gpurun_out/r2v/bench_bgr.json 2> gpurun_out/r2v/bench_bgr.err; echo "bgr rc=$?"
python bench.py --workload ilids_720x576_bgr_defaults --streams 256 $K > gpurun_out/r2v/bench_ilids.json 2> gpurun_out/r2v/bench_ilids.err; echo "ilids rc=$?"
python - <<'PY'
import json
for n in ("bgr","ilids"):
    try:
        d=json.loads(open("gpurun_out/r2v/bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "value %.0f e2e %.0f frac %.3f kernel_us/frame %.2f parity %s"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_us_per_frame"], d.get("parity_checked")))
    except Exception as e: print(n, "failed", e)
PY
timeout 300 python tools/detect_profile.py > gpurun_out/r2v/detect_profile.txt 2>&1; head -40 gpurun_out/r2v/detect_profile.txt
