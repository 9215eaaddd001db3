mkdir -p gpurun_out/r3n
timeout 900 python -m pytest tests -m gpu -x -q -k "fd or blur or knn or dropin or stages or fused or framediff" > gpurun_out/r3n/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3n/pytest.log
K="--steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-dropin"
for WL in 320x240_fd_open_blobs 1080p_bgr_knn; do
python bench.py --workload $WL $K > gpurun_out/r3n/b.json 2> gpurun_out/r3n/b.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r3n/b.json").read().strip().splitlines()[-1])
print("$WL value %.0f e2e %.0f frac %.3f kernel_us/frame %.2f parity %s stage %s"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_us_per_frame"], d.get("parity_checked"), d.get("stage_ms")))
PY
done
