mkdir -p gpurun_out/r3q
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=2
( time $TR --nproc-per-node $N --master-port $((29700+N)) bench.py --gpus $N > gpurun_out/r3q/bench_n$N.json 2> gpurun_out/r3q/bench_n$N.err ) 2>&1 | grep real; echo "bench N=$N rc=$?"
tail -c 300 gpurun_out/r3q/bench_n$N.err
( time $TR --nproc-per-node $N --master-port $((29800+N)) bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/r3q/bench_ref_n$N.json 2> gpurun_out/r3q/bench_ref_n$N.err ) 2>&1 | grep real
python - <<'PY'
import json
for N in (2,):
    try:
        d=json.loads(open('gpurun_out/r3q/bench_n%d.json'%N).read().strip().split('\n')[-1])
        e=d['e2e']
        print('N',N,'value',round(d['value']),'per gpu',round(d['value']/N),'e2e',round(e['value']),'floor',round(e['host_link_floor']['value']),'e2e/floor',round(e['host_link_floor']['e2e_over_floor'],3))
        for k,c in d['configs'].items():
            print('  ',k,'value',round(c['value']),'e2e',round(c['e2e']['value']),'frac',round(c['roofline']['frac'],3))
    except Exception as ex: print('N',N,'failed',ex)
    try:
        r=json.loads(open('gpurun_out/r3q/bench_ref_n%d.json'%N).read().strip().split('\n')[-1]); print('ref', r.get('value'), r.get('n_gpus'))
    except Exception as ex: print('ref failed', ex)
PY
