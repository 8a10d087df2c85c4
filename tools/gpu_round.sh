mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -s 1 -c 1 -o gpurun_out/prof_tf python tools/profile_trace.py --tf 1 --spp 4 --launches 2 > gpurun_out/prof_tf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -s 1 -c 1 -o gpurun_out/prof_notf python tools/profile_trace.py --tf 0 --spp 4 --launches 2 > gpurun_out/prof_notf.log 2>&1
python tools/profile_trace.py --tf 1 --spp 16 --launches 4 > gpurun_out/time_tf.log 2>&1
python tools/profile_trace.py --tf 0 --spp 16 --launches 4 > gpurun_out/time_notf.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_n1.json; cat gpurun_out/bench_ref.json; tail -2 gpurun_out/time_tf.log gpurun_out/time_notf.log
