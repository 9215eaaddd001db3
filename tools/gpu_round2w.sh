mkdir -p gpurun_out/r2w
timeout 300 python tools/latency_breakdown.py 2>&1 | tee gpurun_out/r2w/latency.txt
