# GPU-box job (gpurun): ncu --set full of the tracking kernel on c1 / c2 / c3, traffic stamped with the source hash, bench N = 1
mkdir -p gpurun_out
for sc in c3 c1 c2; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_pool -s 2 -c 1 -o gpurun_out/prof_pool_$sc -f python tools/profile_trace.py --scene $sc --spp 32 --launches 3 > gpurun_out/prof_pool_$sc.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_pool_$sc.ncu-rep > gpurun_out/sum_pool_$sc.txt 2>&1
  python tools/ncu_lines.py gpurun_out/prof_pool_$sc.ncu-rep k_trace 100 > gpurun_out/lines_pool_$sc.txt 2>&1
  python tools/ncu_opcodes.py gpurun_out/prof_pool_$sc.ncu-rep k_trace > gpurun_out/ops_pool_$sc.txt 2>&1
done
timeout 900 python tools/capture_traffic.py gpurun_out/traffic_latest.json 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"] / 1e9, "e2e", d["e2e"]["value"] / 1e9, "frac", d["roofline"]["frac"], d["roofline"]["bound"], "traffic", d["roofline"]["traffic"])
for k, c in d.get("configs", {}).items():
    print(k, c.get("value"), c.get("unit"), c.get("roofline", {}).get("frac"), c.get("roofline", {}).get("traffic"), c.get("brick_build", {}).get("ms"))
PY
for sc in c1 c2 c3; do grep -E "time_duration|thread_inst_executed_per|inst_executed.sum|issue_active" gpurun_out/sum_pool_$sc.txt | sed 's/  */ /g'; done
