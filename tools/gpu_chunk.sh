timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for c in 1 2 4 8 16; do
  for tf in 1 0; do
    echo "chunk=$c tf=$tf $(VRB200_CHUNK=$c python tools/profile_trace.py --tf $tf --spp 16 --launches 4 | tail -1)"
  done
done
echo "spp64 chunk=4 $(VRB200_CHUNK=4 python tools/profile_trace.py --tf 1 --spp 64 --launches 3 | tail -1)"
echo "spp64 chunk=64 $(VRB200_CHUNK=64 python tools/profile_trace.py --tf 1 --spp 64 --launches 3 | tail -1)"
echo "spp1 $(python tools/profile_trace.py --tf 1 --spp 1 --launches 5 | tail -1)"
