# 2 GPUs: the single-process multi-GPU test (vrb_reduce / vrb_copy_rows) + bench at N = 2 (weak value, strong object, collective_ms)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_host.py -m gpu -q -k "multi_gpu" 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_n2.json") if l.startswith("{")][-1])
print("N=2 value", d["value"] / 1e9, "e2e", d["e2e"]["value"] / 1e9, "collective_ms", d.get("collective_ms"), "scaling", d["scaling"])
print("strong", d.get("strong"))
for k, c in d.get("configs", {}).items():
    print(k, c.get("value"), c.get("unit"), c.get("partition"), c.get("collective_ms"))
PY
tail -3 gpurun_out/bench_n2.err
