# round 2: full GPU suite + the new bench.py
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -n "passed\|failed\|FAILED\|Error" gpurun_out/pytest_gpu.log | tail -12
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
tail -c 6000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
