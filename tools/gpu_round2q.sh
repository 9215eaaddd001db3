mkdir -p gpurun_out/r2q
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -x -q -k "mog2 or headline or ilids or bgr" > gpurun_out/r2q/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r2q/pytest.log
K="--steps 6 --warmup 3 --no-configs --no-dropin --no-cpu-baseline --no-parity"
for v in "" libgofpid_b200_b6.so libgofpid_b200_b7.so; do
  for wl in "ilids_720x576_bgr_defaults --streams 64" "1080p_bgr_mog2_open_close_blobs"; do
    GOFPID_LIB=$v python bench.py --workload $wl $K 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('lib=$v', d['config']['workload'], 'value %.0f frac %.3f us/frame %.1f'%(d['value'], d['roofline']['frac'], d['roofline']['kernel_us_per_frame']))"
  done
done > gpurun_out/r2q/variants.txt 2>&1
cat gpurun_out/r2q/variants.txt
python bench.py --steps 20 --warmup 5 --no-configs --no-dropin --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('headline value %.0f e2e %.0f frac %.3f us/frame %.2f'%(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_us_per_frame']))"
