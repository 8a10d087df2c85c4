# lean scaling check (run with gpurun --gpus N): bench at 1, 2, 4[, 8] ranks + the single-process multi-GPU host test
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_s$n.json 2> gpurun_out/bench_s$n.err
    tail -n 2 gpurun_out/bench_s$n.err
  fi
done
timeout 300 python -m pytest tests/test_gpu_host.py -m gpu -x -q -k "multi_gpu" > gpurun_out/pytest_multi.log 2>&1; tail -2 gpurun_out/pytest_multi.log
cat gpurun_out/bench_s*.json | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print(d['n_gpus'], 'value %.2f G' % (d['value'] / 1e9), 'e2e %.2f G' % (d['e2e']['value'] / 1e9), 'ms/step %.3f' % d['ms_per_step'])
"
