mkdir -p gpurun_out/r3k
timeout 900 python -m pytest tests -m gpu -x -q -k "clip or sequence or host or dropin or groups or fullsize" > gpurun_out/r3k/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3k/pytest.log
for V in "" "GOFPID_DEPTH=16"; do echo "== $V"; env $V timeout 200 python tools/e2e_clip_floor.py 2>&1 | tail -5; done | tee gpurun_out/r3k/e2e_clip_floor.txt
K="--steps 5 --warmup 3 --no-cpu-baseline --no-dropin"
python bench.py $K > gpurun_out/r3k/bench.json 2> gpurun_out/r3k/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3k/bench.json").read().strip().splitlines()[-1])
print("headline value %.0f e2e %.0f floor %.0f per-frame %.0f frac %.3f"%(d["value"], d["e2e"]["value"], d["e2e"]["host_link_floor"]["value"], d["e2e"]["per_frame_call"]["value"], d["roofline"]["frac"]))
for k,c in d.get("configs",{}).items():
    print(k, "value %.0f e2e %.0f floor %s"%(c["value"], c["e2e"]["value"], c["e2e"].get("host_link_floor",{}).get("value")))
PY
