mkdir -p gpurun_out/r2f
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -x -q -k "blob or clip or headline or small or graph or device or dropin or noise or 4k or knn" > gpurun_out/r2f/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 5 gpurun_out/r2f/pytest.log
for MC in 8 32; do for G in 3 4 8; do
  echo "== MAX_CONNECTIONS=$MC G=$G small(warp tracer)"; CUDA_DEVICE_MAX_CONNECTIONS=$MC GOFPID_DEPTH=16 timeout 200 python tools/stage_floor.py $G 2>&1 | grep -E "full"
done; done > gpurun_out/r2f/floor.txt 2>&1
cat gpurun_out/r2f/floor.txt
