mkdir -p gpurun_out/r2u
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -x -q -k "morph or headline or 4k or fd or blobs or dropin" > gpurun_out/r2u/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r2u/pytest.log
for NC in 24 72; do echo "== clip $NC frames"; FLOOR_NCLIP=$NC timeout 300 python tools/stage_floor.py 3 2>&1 | tail -3; done
