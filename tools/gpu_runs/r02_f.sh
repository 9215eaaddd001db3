mkdir -p gpurun_out/r3f
K="--steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-dropin --no-parity"
export GOFPID_NO_COND=1
python bench.py $K > gpurun_out/r3f/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 300 --csv --log-file gpurun_out/r3f/launches_headline.csv python bench.py $K > gpurun_out/r3f/ncu2.log 2>&1
echo "ncu launches rc=$?"
