timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v6b.json 2> gpurun_out/bench_v6b.err; cat gpurun_out/bench_v6b.json; tail -5 gpurun_out/bench_v6b.err
