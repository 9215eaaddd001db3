set -x
mkdir -p gpurun_out/r2c
python bench.py --steps 20 --warmup 5 --no-dropin > gpurun_out/r2c/bench.json 2> gpurun_out/r2c/bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2c/bench.err
python tools/pcie_bw.py > gpurun_out/r2c/pcie_bw.txt 2>&1; cat gpurun_out/r2c/pcie_bw.txt
python tools/pcie_bw.py --wc > gpurun_out/r2c/pcie_bw_wc.txt 2>&1; cat gpurun_out/r2c/pcie_bw_wc.txt
GOFPID_NO_COND=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-dropin --no-parity > gpurun_out/r2c/plain.log 2>&1 &&
GOFPID_NO_COND=1 ncu --set full --clock-control none --import-source on -k regex:mog2_lean -s 30 -c 1 -o gpurun_out/r2c/prof_mog2_lean python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-dropin --no-parity > gpurun_out/r2c/ncu1.log 2>&1
echo "ncu rc=$?"
GOFPID_NO_COND=1 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 300 --csv --log-file gpurun_out/r2c/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-dropin --no-parity > gpurun_out/r2c/ncu2.log 2>&1
echo "ncu2 rc=$?"
