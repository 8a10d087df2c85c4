"""Summarise an ncu report per CUDA source line: instructions executed, thread efficiency, stall samples.
    python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top-N]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
ksub = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines = {}
fpath = func = None
seen_funcs = []
cols = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        func = r[1]
        if func not in seen_funcs:
            seen_funcs.append(func)
    elif r[0] == "Line No":
        cols = r
    elif r[0] not in ("", "Kernel Name") and cols and r[0].isdigit():
        if ksub not in (func or "") or func != [f for f in seen_funcs if ksub in f][0]:
            continue
        d = dict(zip(cols, r))
        try:
            inst = int(d["Instructions Executed"]); thr = int(d["Thread Instructions Executed"]); smp = int(d["# Samples"])
        except ValueError:
            continue
        key = (fpath, int(r[0]))
        a = lines.setdefault(key, [0, 0, 0, r[1].strip()[:110]])
        a[0] += inst; a[1] += thr; a[2] += smp
tot_i = sum(v[0] for v in lines.values()); tot_t = sum(v[1] for v in lines.values()); tot_s = sum(v[2] for v in lines.values())
print(f"kernel: {[f for f in seen_funcs if ksub in f][0][:90]}")
print(f"total warp-inst {tot_i:,}  thread-inst {tot_t:,}  avg active threads {tot_t/max(tot_i,1):.2f}  stall samples {tot_s:,}")
print(f"{'file:line':28s} {'inst%':>6s} {'thr/inst':>8s} {'smp%':>6s}  source")
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f+':'+str(ln):28s} {100*v[0]/tot_i:6.2f} {v[1]/max(v[0],1):8.2f} {100*v[2]/max(tot_s,1):6.2f}  {v[3]}")
