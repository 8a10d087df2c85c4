mkdir -p gpurun_out
echo "== current lib (pool edits + build rewrite)"; timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -q -k "scheduling_options or cached_tile_order or ray_pool or brick_mask" 2>&1 | tail -4
echo "== old lib 423401c"; VRB200_LIB=$PWD/tools/_build/old_423401c.so timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -q -k "scheduling_options or cached_tile_order or ray_pool or brick_mask" 2>&1 | tail -4
echo "== cpu baseline worker"; timeout 300 python bench.py --cpu-baseline-worker; echo "rc $?"
echo "== sweep"; timeout 1500 python tools/sweep.py run --scenes c1,c3,c2,c4 --spp 32 --launches 4 > gpurun_out/sweep2.log 2>&1; cut -c1-40,41-44,100-200 gpurun_out/sweep2.log | sed 's/"res.*best_ms/ms/' 
