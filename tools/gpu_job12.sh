python tools/debug_sched.py 2>&1 | tail -30
echo "== worker with faulthandler"; timeout 300 python -X faulthandler bench.py --cpu-baseline-worker 2>&1 | tail -25
echo "== nproc $(nproc)"; python -c "import os; print(len(os.sched_getaffinity(0)))"
echo "== glsl tests on this box"; timeout 600 python -m pytest tests/test_glsl_ref.py -q -x 2>&1 | tail -5
