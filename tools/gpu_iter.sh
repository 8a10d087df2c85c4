# quick iteration on the GPU box: parity tests, kernel timing (both variants), bench line, optional extras via $1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python tools/profile_trace.py --tf 1 --spp 32 --launches 4 > gpurun_out/time_tf.log 2>&1
python tools/profile_trace.py --tf 0 --spp 32 --launches 4 > gpurun_out/time_notf.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -4 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/time_tf.log; tail -n 2 gpurun_out/time_notf.log; cat gpurun_out/bench_n1.json
if [ -n "$1" ]; then bash -c "$1"; fi
