echo "ulimit -s: $(ulimit -s)"; 
echo "== OMP_STACKSIZE=64M"; OMP_STACKSIZE=64M timeout 300 python bench.py --cpu-baseline-worker 2>&1 | tail -2 | cut -c1-200; echo "rc $?"
echo "== as is"; timeout 300 python bench.py --cpu-baseline-worker 2>&1 | tail -2 | cut -c1-200
echo "== taskset 8 cores"; taskset -c 0-7 timeout 300 python bench.py --cpu-baseline-worker 2>&1 | tail -2 | cut -c1-200
