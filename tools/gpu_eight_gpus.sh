# GPU-box job (gpurun --gpus 8): bench at N = 4 and N = 8 (weak value, strong object, per-config splits)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n$n exit $?"
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_n$n.json") if l.startswith("{")][-1])
print("N=$n value", d["value"] / 1e9, "e2e", d["e2e"]["value"] / 1e9, "collective_ms", d.get("collective_ms"), "ms_per_step", d["ms_per_step"])
print("  strong", {k: d["strong"][k] for k in ("value", "ms_per_frame", "collective_ms")} if d.get("strong") else None)
for k, c in d.get("configs", {}).items():
    print("  ", k, c.get("value"), c.get("unit"), c.get("collective_ms"))
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
