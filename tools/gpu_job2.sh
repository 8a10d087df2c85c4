# round 2: pool-kernel profile (c1 non-TF, c2 TF) + the GPU tests that did not run yet
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for sc in c1 c2; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_pool -s 2 -c 1 -o gpurun_out/prof_pool_$sc python tools/profile_trace.py --scene $sc --spp 32 --launches 3 > gpurun_out/prof_pool_$sc.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_pool_$sc.ncu-rep > gpurun_out/sum_pool_$sc.txt 2>&1
  python tools/ncu_lines.py gpurun_out/prof_pool_$sc.ncu-rep k_trace 90 > gpurun_out/lines_pool_$sc.txt 2>&1
done
tail -8 gpurun_out/pytest_gpu.log
