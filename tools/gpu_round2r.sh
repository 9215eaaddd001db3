mkdir -p gpurun_out/r2r
K="--steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-dropin --no-parity"
export GOFPID_NO_COND=1
python bench.py $K > gpurun_out/r2r/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"mog2_lean|morph_bits|trace_kernel" -s 120 -c 6 -o gpurun_out/r2r/prof_headline python bench.py $K > gpurun_out/r2r/ncu1.log 2>&1
echo "ncu headline rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 300 --csv --log-file gpurun_out/r2r/launches_headline.csv python bench.py $K > gpurun_out/r2r/ncu2.log 2>&1
echo "ncu launches rc=$?"
python bench.py --workload 1080p_knn_many_blobs $K > gpurun_out/r2r/plain_knn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn_gray -s 60 -c 1 -o gpurun_out/r2r/prof_knn_gray python bench.py --workload 1080p_knn_many_blobs $K > gpurun_out/r2r/ncu_knn.log 2>&1
echo "ncu knn rc=$?"
python bench.py --workload ilids_720x576_bgr_defaults --streams 256 $K > gpurun_out/r2r/plain_ilids.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mog2_lean -s 6 -c 1 -o gpurun_out/r2r/prof_mog2_lean_bgr256 python bench.py --workload ilids_720x576_bgr_defaults --streams 256 $K > gpurun_out/r2r/ncu_ilids.log 2>&1
echo "ncu ilids rc=$?"
