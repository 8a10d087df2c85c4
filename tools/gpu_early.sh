# early-rejection round: parity tests, early-rejection rates (counting build), A/B sweep of the variants in tools/_sweep, bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python tools/profile_trace.py --tf 1 --spp 16 --launches 2 --count 1 > gpurun_out/count_tf.log 2>&1
python tools/profile_trace.py --tf 0 --spp 16 --launches 2 --count 1 > gpurun_out/count_notf.log 2>&1
tail -n 2 gpurun_out/count_tf.log gpurun_out/count_notf.log
timeout 900 python tools/sweep.py run > gpurun_out/sweep_early.txt 2>&1
cat gpurun_out/sweep_early.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
