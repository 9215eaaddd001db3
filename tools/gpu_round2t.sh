mkdir -p gpurun_out/r2t
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2t/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r2t/pytest.log
K="--no-dropin --no-cpu-baseline --no-parity --no-configs"
for wl in "ilids_720x576_bgr_defaults --streams 256 --steps 4" "320x240_fd_open_blobs --steps 6" "1080p_mog2_open_close_blobs --steps 20"; do
python bench.py --workload $wl --warmup 3 $K 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); e=d['e2e']; print(d['config']['workload'], 'value %.0f e2e %.0f floor %.0f'%(d['value'], e['value'], e['host_link_floor']['value']), e.get('clip_call',{}).get('value'), e.get('per_frame_call',{}).get('value'))"
done
