# two-rays-per-lane kernel: parity test + timing against the one-ray kernel
python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "two_rays or persistent_kernel" 2>&1 | tail -5
for k in 0 3; do for tf in 1 0; do echo "kernel $k tf $tf"; VRB200_KERNEL=$k python tools/profile_trace.py --tf $tf --spp 32 --launches 4 --kernel $k | tail -2; done; done
