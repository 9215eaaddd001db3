mkdir -p gpurun_out/r2g
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r2g/pytest.log
python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/r2g/bench.json 2> gpurun_out/r2g/bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2g/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g/bench.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'perframe',d['e2e']['per_frame_call']['value'],'frac',d['roofline']['frac'],'dropin',d.get('dropin'))
PY
python tools/pcie_bw.py --sizes 2,4,6,8,12,16,24,32 > gpurun_out/r2g/pcie_sizes.txt 2>&1; cat gpurun_out/r2g/pcie_sizes.txt
python tools/pcie_bw.py --wc --sizes 2,4,6,8,12,16,24,32 > gpurun_out/r2g/pcie_sizes_wc.txt 2>&1; cat gpurun_out/r2g/pcie_sizes_wc.txt
