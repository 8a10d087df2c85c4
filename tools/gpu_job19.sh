mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py -m gpu -q -k "ray_pool or emission or scheduling or T2 or T3 or brick_mask or cached" 2>&1 | tail -3
for sc in c2 c4 c1 c3; do
  timeout 300 python tools/profile_trace.py --scene $sc --spp 32 --launches 6 --json 1 2>&1 | grep "JSON\|rror" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('JSON'):
        d=json.loads(l[5:]); print(d['scene'], round(d['gsamples_per_s'],3), 'Gsamples/s', round(d['best_ms'],3), 'ms')
    else: print(l[:200])
"
done
