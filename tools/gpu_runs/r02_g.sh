mkdir -p gpurun_out/r3g
K="--steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-dropin --no-parity"
export GOFPID_NO_COND=1
python bench.py --workload 1080p_knn_many_blobs $K > gpurun_out/r3g/plain_knn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn_gray -s 60 -c 1 -o gpurun_out/r3g/prof_knn_gray python bench.py --workload 1080p_knn_many_blobs $K > gpurun_out/r3g/ncu_knn.log 2>&1
echo "ncu knn rc=$?"
