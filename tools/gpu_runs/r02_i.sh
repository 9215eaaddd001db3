mkdir -p gpurun_out/r3i
K="--steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-dropin"
export GOFPID_NO_COND=1
python bench.py --workload 1080p_bgr_knn $K > gpurun_out/r3i/plain_knn_bgr.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s 30 -c 1 -o gpurun_out/r3i/prof_knn_bgr python bench.py --workload 1080p_bgr_knn $K --no-parity > gpurun_out/r3i/ncu_knn_bgr.log 2>&1
echo "ncu knn bgr rc=$?"
tail -c 600 gpurun_out/r3i/plain_knn_bgr.log
