# A/B timing of the tracking kernel variants inside ONE box (boxes differ by a few percent)
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
for lpt in 0 1; do
  for tf in 1 0; do
    echo "lpt=$lpt tf=$tf spp16 $(VRB200_LPT=$lpt python tools/profile_trace.py --tf $tf --spp 16 --launches 5 | tail -1)"
  done
done
done
echo "lpt=0 tf=1 spp64 $(VRB200_LPT=0 python tools/profile_trace.py --tf 1 --spp 64 --launches 4 | tail -1)"
echo "lpt=1 tf=1 spp64 $(VRB200_LPT=1 python tools/profile_trace.py --tf 1 --spp 64 --launches 4 | tail -1)"
echo "lpt=0 tf=1 spp1 $(VRB200_LPT=0 python tools/profile_trace.py --tf 1 --spp 1 --launches 6 | tail -1)"
echo "lpt=1 tf=1 spp1 $(VRB200_LPT=1 python tools/profile_trace.py --tf 1 --spp 1 --launches 6 | tail -1)"
