mkdir -p gpurun_out/r2k
nvidia-smi topo -m > gpurun_out/r2k/topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > gpurun_out/r2k/lscpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4 2; do
  $TR --nproc-per-node $N --master-port $((29500+N)) tools/pcie_bw.py --sizes 2,6,24 2>&1 | grep ranks= >> gpurun_out/r2k/pcie_ranks.txt
done
$TR --nproc-per-node 8 --master-port 29611 tools/pcie_bw.py --sizes 6 --wc 2>&1 | grep ranks= >> gpurun_out/r2k/pcie_ranks.txt
$TR --nproc-per-node 8 --master-port 29612 tools/pcie_bw.py --sizes 6 --affinity 2>&1 | grep ranks= >> gpurun_out/r2k/pcie_ranks.txt
cat gpurun_out/r2k/pcie_ranks.txt
for N in 8 4; do
  $TR --nproc-per-node $N --master-port $((29700+N)) bench.py --gpus $N --steps 20 --warmup 5 --no-dropin > gpurun_out/r2k/bench_n$N.json 2> gpurun_out/r2k/bench_n$N.err; echo "bench N=$N rc=$?"
  tail -c 400 gpurun_out/r2k/bench_n$N.err
done
python - <<'PY'
import json
for N in (8,4):
    try:
        d=json.loads(open('gpurun_out/r2k/bench_n%d.json'%N).read().strip().split('\n')[-1])
        e=d['e2e']
        print('N',N,'value',round(d['value']),'per gpu',round(d['value']/N),'e2e',round(e['value']),'floor',round(e['host_link_floor']['value']),'e2e/floor',round(e['host_link_floor']['e2e_over_floor'],3))
        c=d['configs']['ilids_720x576_bgr_defaults']
        print('   ilids256 value',round(c['value']),'e2e',round(c['e2e']['value']),'floor',round(c['e2e']['host_link_floor']['value']),'frac',round(c['roofline']['frac'],3),'streams/gpu',c['config']['streams_per_gpu'])
    except Exception as ex: print('N',N,'failed',ex)
PY
