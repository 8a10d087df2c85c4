# round 2: pool-kernel v2 (incremental counts): bit-identity tests, parameter sweep on c1/c2/c3/c4, datagen scripts
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py -m gpu -q -s -k "ray_pool or persistent_kernel or running_mean or ieee or T3 or scheduling" > gpurun_out/pytest_pool.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_pool.log
tail -4 gpurun_out/pytest_pool.log
for sc in c1 c2 c3 c4; do timeout 300 python tools/profile_trace.py --scene $sc --kernel 3 --spp 32 --launches 4 --json 1 2>&1 | grep "JSON\|rror"; done > gpurun_out/sweep.log
timeout 1500 python tools/sweep.py run --scenes c1,c2,c3,c4 --spp 32 --launches 4 >> gpurun_out/sweep.log 2>&1
cat gpurun_out/sweep.log | cut -c1-200
timeout 1500 python -m pytest tests/test_gpu_host.py -m gpu -q -s -k "datagen" > gpurun_out/pytest_datagen.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_datagen.log
tail -15 gpurun_out/pytest_datagen.log
