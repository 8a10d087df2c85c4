# e2e leg of bench.py with 2 / 3 / 4 pipelined contexts
for L in ${LANES:-2 3 4}; do VRB_E2E_LANES=$L python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes', d['e2e'].get('lanes'), 'value %.2f G' % (d['value'] / 1e9), 'e2e %.2f G' % (d['e2e']['value'] / 1e9))"; done
