mkdir -p gpurun_out/r2o
for NC in 24 72; do echo "== clip $NC frames, model evict_last"; GOFPID_NO_PREFETCH=1 FLOOR_NCLIP=$NC timeout 300 python tools/stage_floor.py 3 2>&1 | tail -3; done > gpurun_out/r2o/floor.txt 2>&1
cat gpurun_out/r2o/floor.txt
timeout 600 python -m pytest tests/test_gpu_pipeline.py -x -q -k "mog2_frame_groups and not generic" 2>&1 | tail -2
