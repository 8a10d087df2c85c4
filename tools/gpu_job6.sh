# round 2 (session 2): LUT encode + build timing, pool kernel v3 profile on the non-TF configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_brick.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/pytest_brick.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_brick.log
tail -15 gpurun_out/pytest_brick.log
timeout 300 python tools/gpu_build_timing.py 1024 > gpurun_out/build_timing.log 2>&1; cat gpurun_out/build_timing.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_range|k_scan|k_brick|k_make|k_linear" -c 60 --csv --log-file gpurun_out/launches_build.csv python tools/gpu_build_timing.py 1024 > gpurun_out/launches_build.log 2>&1
for sc in c3 c1; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_pool -s 2 -c 1 -o gpurun_out/prof_pool3_$sc -f python tools/profile_trace.py --scene $sc --spp 32 --launches 3 > gpurun_out/prof_pool3_$sc.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_pool3_$sc.ncu-rep > gpurun_out/sum_pool3_$sc.txt 2>&1
  python tools/ncu_lines.py gpurun_out/prof_pool3_$sc.ncu-rep k_trace 120 > gpurun_out/lines_pool3_$sc.txt 2>&1
  python tools/ncu_opcodes.py gpurun_out/prof_pool3_$sc.ncu-rep k_trace > gpurun_out/ops_pool3_$sc.txt 2>&1
done
head -12 gpurun_out/sum_pool3_c3.txt
