# GPU-box job (gpurun): smoke(), the whole GPU suite, the reference arm, bench N = 1
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -5
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-400 gpurun_out/bench_ref.json
( time timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>&1 | grep real; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"] / 1e9, "e2e", d["e2e"]["value"] / 1e9, "frac", d["roofline"]["frac"], d["roofline"]["bound"], "traffic", d["roofline"]["traffic"], d["roofline"]["traffic_source"][:60])
for k, c in d.get("configs", {}).items():
    print(k, c.get("value"), c.get("unit"), c.get("roofline", {}).get("frac"), c.get("roofline", {}).get("traffic"), c.get("brick_build", {}).get("ms"))
print("cpu", d.get("cpu_baseline"))
print("clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
PY
