"""Static SASS size per CUDA source line for one kernel (code-size hot spots).
    python tools/sass_lines.py <lib.so> <kernel-mangled-substring> [top]"""
import collections, re, subprocess, sys, tempfile, os, glob
lib, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cubin = glob.glob(os.path.join(d, "*.cubin"))[0]
txt = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
cnt = collections.Counter(); cur = None; active = False
for ln in txt:
    if ln.startswith(".text."):
        active = ksub in ln; continue
    if not active: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        m2 = re.search(r'inlined at "([^"]+)", line (\d+)', ln)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln) and cur: cnt[cur] += 1
tot = sum(cnt.values()); print("total instructions", tot, "=", tot * 16 / 1024, "KB")
for k, v in cnt.most_common(top): print(f"{v:5d} {100*v/tot:5.1f}%  {k[0]}:{k[1]}")
