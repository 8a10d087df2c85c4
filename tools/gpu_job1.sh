# round 2, first GPU call: parity suite + old-vs-new kernel timing on C1..C4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for sc in c2 c1 c3 c4; do for k in 3 0; do
  timeout 600 python tools/profile_trace.py --scene $sc --kernel $k --spp 32 --launches 4 --json 1 2>&1 | grep "JSON\|rror" >> gpurun_out/kernels.log
done; done
timeout 300 python tools/profile_trace.py --scene c3 --kernel 0 --spp 4 --launches 1 --count 1 --json 1 2>&1 | grep "JSON\|rror" >> gpurun_out/kernels.log
timeout 300 python tools/profile_trace.py --scene c1 --kernel 0 --spp 4 --launches 1 --count 1 --json 1 2>&1 | grep "JSON\|rror" >> gpurun_out/kernels.log
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/kernels.log
