# full GPU suite + bench after the brick-build rewrite
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -12
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"] / 1e9, "e2e", d["e2e"]["value"] / 1e9, "frac", d["roofline"]["frac"], d["roofline"]["bound"])
for k, c in d.get("configs", {}).items():
    print(k, c.get("value"), c.get("unit"), c.get("roofline", {}).get("frac"), c.get("brick_build", {}).get("ms"))
print("cpu", d.get("cpu_baseline"))
PY
tail -3 gpurun_out/bench_n1.err
