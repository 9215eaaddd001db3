mkdir -p gpurun_out/r2m
for NC in 24 72; do echo "== clip $NC frames"; FLOOR_NCLIP=$NC timeout 300 python tools/stage_floor.py 3 2>&1 | tail -3; done > gpurun_out/r2m/floor_clip.txt 2>&1
cat gpurun_out/r2m/floor_clip.txt
