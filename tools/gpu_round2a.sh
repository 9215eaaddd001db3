set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
tail -5 gpurun_out/r2a/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2a/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r2a/bench_ref.json 2>&1; echo "ref rc=$?"
timeout 300 python tools/stage_floor.py 0,3 > gpurun_out/r2a/stage_floor.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/r2a/memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/r2a/racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/r2a/memcheck.log gpurun_out/r2a/racecheck.log
