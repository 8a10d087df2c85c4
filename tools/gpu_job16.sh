# final captures of round 2 (code frozen): ncu --set full of the tracking kernel on c1 / c2 / c3, traffic stamped with the source hash,
# launch list of the bench command, bench N = 1
mkdir -p gpurun_out
for sc in c3 c1 c2; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_pool -s 2 -c 1 -o gpurun_out/prof_pool4_$sc -f python tools/profile_trace.py --scene $sc --spp 32 --launches 3 > gpurun_out/prof_pool4_$sc.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_pool4_$sc.ncu-rep > gpurun_out/sum_pool4_$sc.txt 2>&1
  python tools/ncu_lines.py gpurun_out/prof_pool4_$sc.ncu-rep k_trace 100 > gpurun_out/lines_pool4_$sc.txt 2>&1
  python tools/ncu_opcodes.py gpurun_out/prof_pool4_$sc.ncu-rep k_trace > gpurun_out/ops_pool4_$sc.txt 2>&1
done
timeout 900 python tools/capture_traffic.py gpurun_out/traffic_latest.json 2>&1 | tail -6
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --configs none --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
python tools/launch_table.py gpurun_out/launches_bench.csv 1 2>&1 | tail -30
