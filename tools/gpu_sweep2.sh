for rep in 1 2; do echo "== rep $rep"; python tools/sweep.py run --kernel ${KERNEL:-0}; done
