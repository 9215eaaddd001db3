mkdir -p gpurun_out/r3k
K="--no-cpu-baseline --no-dropin --no-configs"
for i in 1 2; do
python bench.py $K > gpurun_out/r3k/bench_$i.json 2> gpurun_out/r3k/bench_$i.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r3k/bench_$i.json").read().strip().splitlines()[-1])
print("headline value %.0f e2e %.0f floor %.0f per-frame %.0f frac %.3f"%(d["value"], d["e2e"]["value"], d["e2e"]["host_link_floor"]["value"], d["e2e"]["per_frame_call"]["value"], d["roofline"]["frac"]))
PY
done
