# timing of the tracking kernel inside ONE box (boxes differ by a few percent)
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pass in 16 8 32; do
  for tf in 1 0; do
    echo "pass=$pass tf=$tf spp16 $(VRB200_PASS=$pass python tools/profile_trace.py --tf $tf --spp 16 --launches 5 | tail -1)"
  done
done
echo "lpt=0 tf=1 spp16 $(VRB200_LPT=0 python tools/profile_trace.py --tf 1 --spp 16 --launches 5 | tail -1)"
echo "lpt=0 tf=0 spp16 $(VRB200_LPT=0 python tools/profile_trace.py --tf 0 --spp 16 --launches 5 | tail -1)"
echo "tf=1 spp64 $(python tools/profile_trace.py --tf 1 --spp 64 --launches 4 | tail -1)"
echo "tf=1 spp64 pass32 $(VRB200_PASS=32 python tools/profile_trace.py --tf 1 --spp 64 --launches 4 | tail -1)"
echo "tf=1 spp1 $(python tools/profile_trace.py --tf 1 --spp 1 --launches 6 | tail -1)"
echo "tf=0 spp1 $(python tools/profile_trace.py --tf 0 --spp 1 --launches 6 | tail -1)"
