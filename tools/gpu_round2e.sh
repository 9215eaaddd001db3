mkdir -p gpurun_out/r2e
for MC in 8 32; do for G in 3 8; do for L in 0 1; do
  echo "== MAX_CONNECTIONS=$MC G=$G B_LARGE=$L"; CUDA_DEVICE_MAX_CONNECTIONS=$MC GOFPID_DEPTH=16 GOFPID_B_LARGE=$L timeout 200 python tools/stage_floor.py $G 2>&1 | grep -E "full"
done; done; done > gpurun_out/r2e/floor_connections.txt 2>&1
cat gpurun_out/r2e/floor_connections.txt
