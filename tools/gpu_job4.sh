# round 2: pool kernel v3 (register-resident SEGMENT loop)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py -m gpu -q -s -k "ray_pool or running_mean or T2 or T3 or scheduling or brick_mask or cached" > gpurun_out/pytest_pool.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_pool.log
tail -6 gpurun_out/pytest_pool.log
for sc in c1 c2 c3 c4; do for k in 3 0; do timeout 300 python tools/profile_trace.py --scene $sc --kernel $k --spp 32 --launches 4 --json 1 2>&1 | grep "JSON\|rror"; done; done > gpurun_out/kernels.log
cat gpurun_out/kernels.log | cut -c1-160
python tools/debug_k12.py > gpurun_out/debug_k12.log 2>&1; cat gpurun_out/debug_k12.log | head -40
for sc in c1 c2; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_pool -s 2 -c 1 -o gpurun_out/prof_pool_$sc python tools/profile_trace.py --scene $sc --spp 32 --launches 3 > gpurun_out/prof_pool_$sc.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_pool_$sc.ncu-rep > gpurun_out/sum_pool_$sc.txt 2>&1
  python tools/ncu_lines.py gpurun_out/prof_pool_$sc.ncu-rep k_trace 70 > gpurun_out/lines_pool_$sc.txt 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3.csv python tools/profile_trace.py --scene c3 --spp 4 --launches 1 > gpurun_out/launches_c3.log 2>&1
