"""Table of a tools/sweep.py run log: variant x scene -> Gsamples/s.   python tools/sweep_table.py gpurun_out/sweep.log"""
import json
import re
import sys

rows = {}
for l in open(sys.argv[1]):
    m = re.match(r"(\S+)\.so\s+(c\d)\s+(\{.*\})", l)
    if m:
        rows.setdefault(m.group(1), {})[m.group(2)] = json.loads(m.group(3))["gsamples_per_s"]
    elif "FAILED" in l:
        print(l[:200].rstrip())
print(f"{'variant':12s} {'c1':>7s} {'c3':>7s} {'c2':>7s} {'c4':>7s}")
for k, v in rows.items():
    print(f"{k:12s} {v.get('c1', 0):7.3f} {v.get('c3', 0):7.3f} {v.get('c2', 0):7.2f} {v.get('c4', 0):7.2f}")
