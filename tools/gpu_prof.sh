timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -s 1 -c 1 -o gpurun_out/prof4_tf python tools/profile_trace.py --tf 1 --spp 16 --launches 2 > gpurun_out/prof4_tf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -s 1 -c 1 -o gpurun_out/prof4_notf python tools/profile_trace.py --tf 0 --spp 16 --launches 2 > gpurun_out/prof4_notf.log 2>&1
python tools/profile_trace.py --tf 1 --spp 16 --launches 3 --count 1
python tools/profile_trace.py --tf 0 --spp 16 --launches 3 --count 1
