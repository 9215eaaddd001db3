mkdir -p gpurun_out/r3h
K="--steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-dropin --no-parity"
export GOFPID_NO_COND=1
python bench.py --workload ilids_720x576_bgr_defaults --streams 256 $K > gpurun_out/r3h/plain_ilids.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mog2_lean -s 20 -c 1 -o gpurun_out/r3h/prof_mog2_lean_bgr256 python bench.py --workload ilids_720x576_bgr_defaults --streams 256 $K > gpurun_out/r3h/ncu_ilids.log 2>&1
echo "ncu ilids rc=$?"
