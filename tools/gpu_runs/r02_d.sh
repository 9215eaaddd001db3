mkdir -p gpurun_out/r3d
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3d/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r3d/pytest.log
( time python bench.py > gpurun_out/r3d/bench.json 2> gpurun_out/r3d/bench.err ) 2>&1 | grep real; echo "bench rc=$?"
( time python bench.py --impl reference > gpurun_out/r3d/bench_ref.json 2> gpurun_out/r3d/bench_ref.err ) 2>&1 | grep real
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3d/smoke.log 2>&1; echo "smoke rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3d/bench.json").read().strip().splitlines()[-1])
print("headline value %.0f e2e %.0f floor %.0f frac %.3f kernel_us/frame %.2f parity %s cpu %.0f"%(d["value"], d["e2e"]["value"], d["e2e"]["host_link_floor"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_us_per_frame"], d.get("parity_checked"), d["cpu_baseline"]["value"]))
print("dropin", d.get("dropin"))
for k,c in d.get("configs",{}).items():
    try: print(k, "value %.0f e2e %.0f frac %.3f us/frame %.2f"%(c["value"], c["e2e"]["value"], c["roofline"]["frac"], c["roofline"]["kernel_us_per_frame"]))
    except Exception as e: print(k, c)
r=json.loads(open("gpurun_out/r3d/bench_ref.json").read().strip().splitlines()[-1])
print("reference", r["value"], r.get("cpu_baseline"))
PY
