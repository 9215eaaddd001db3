set -x
mkdir -p gpurun_out/r2d
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 6 gpurun_out/r2d/pytest.log
for G in 3 4 6 8; do
  echo "== small B, G=$G"; GOFPID_DEPTH=16 timeout 200 python tools/stage_floor.py $G 2>&1 | grep -E "full|morph"
  echo "== large B, G=$G"; GOFPID_DEPTH=16 GOFPID_B_LARGE=1 timeout 200 python tools/stage_floor.py $G 2>&1 | grep full
done > gpurun_out/r2d/floor_small_vs_large.txt 2>&1
cat gpurun_out/r2d/floor_small_vs_large.txt
for G in 4 8; do
  echo "== small B, 7 CTAs/SM, G=$G"; GOFPID_DEPTH=16 GOFPID_MOG2_SMEM_PAD=5000 timeout 200 python tools/stage_floor.py $G 2>&1 | grep -E "full|only"
  echo "== small B, morph 128 threads, G=$G"; GOFPID_DEPTH=16 GOFPID_LIB=libgofpid_b200_mb128.so timeout 200 python tools/stage_floor.py $G 2>&1 | grep -E "full|morph"
done > gpurun_out/r2d/floor_variants.txt 2>&1
cat gpurun_out/r2d/floor_variants.txt
