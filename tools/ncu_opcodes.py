"""Aggregate executed warp-instructions per SASS opcode for one kernel of an ncu report.
    python tools/ncu_opcodes.py report.ncu-rep [kernel-substring]"""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]; ksub = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
ops = collections.Counter(); thr = collections.Counter()
cols = None; active = False; done = False
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        if active: break
        active = ksub in r[1]; continue
    if r[0] == "Address": cols = r; continue
    if not active or cols is None: continue
    d = dict(zip(cols, r))
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", d["Source"])
    if not m: continue
    op = m.group(2).split(".")[0]
    full = ".".join(m.group(2).split(".")[:2]) if op in ("MUFU", "F2I", "I2F", "F2F", "FRND", "I2FP", "F2IP") else op
    try:
        n = int(d["Instructions Executed"]); t = int(d["Thread Instructions Executed"])
    except ValueError:
        continue
    ops[full] += n; thr[full] += t
tot = sum(ops.values())
print(f"total warp-inst {tot:,}")
for op, n in ops.most_common(45):
    print(f"{op:14s} {100*n/tot:6.2f}%  thr/inst {thr[op]/max(n,1):5.1f}")
