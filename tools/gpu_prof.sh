# full ncu captures of the production tracking kernel (third launch, 16 spp) + launch list of the bench command
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -s 2 -c 1 -o gpurun_out/prof_tf python tools/profile_trace.py --tf 1 --spp 32 --launches 3 > gpurun_out/prof_tf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -s 2 -c 1 -o gpurun_out/prof_notf python tools/profile_trace.py --tf 0 --spp 32 --launches 3 > gpurun_out/prof_notf.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
