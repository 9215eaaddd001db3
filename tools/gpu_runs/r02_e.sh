mkdir -p gpurun_out/r3e
K="--steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-dropin --no-parity"
export GOFPID_NO_COND=1
python bench.py $K > gpurun_out/r3e/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"mog2_lean|morph_bits|trace_kernel|nest_kernel|runs_extract|results_to_host" -s 120 -c 9 -o gpurun_out/r3e/prof_headline python bench.py $K > gpurun_out/r3e/ncu1.log 2>&1
echo "ncu headline rc=$?"
