"""DRAM traffic per launch of the tracking kernel on the BASELINE configs, stamped with the hash of the CUDA sources it was
taken from (bench.py reports `roofline.traffic` from this file and ignores it when the sources changed).

    python tools/capture_traffic.py [out.json]        # on a GPU box; runs ncu (metrics pass) once per scene

Writes {source_sha16, c1, c2, c3, c4: {kernel, workload, dram_read_bytes, dram_write_bytes, gpu_time_ns, l2_hit_pct, source}}."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv_saved, sys.argv = sys.argv, ["bench"]
import bench  # noqa: E402  (source_hash)

out_path = sys.argv_saved[1] if len(sys.argv_saved) > 1 else os.path.join(ROOT, "gpurun_out", "traffic_latest.json")
METRICS = "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct"
res = {"source_sha16": bench.source_hash()}
for scene, what in (("c1", "configs[0], 1024x1024"), ("c2", "configs[1], 1920x1080"), ("c3", "configs[2], 1024^3 fBm, 1920x1080"), ("c4", "configs[3], 512x512x1800 CT + Turbo LUT, 1920x1080")):
    cmd = ["ncu", "--metrics", METRICS, "--clock-control", "none", "-k", "regex:k_trace_pool", "-s", "2", "-c", "1", "--csv",
           sys.executable, os.path.join(ROOT, "tools", "profile_trace.py"), "--scene", scene, "--spp", "32", "--launches", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    rows = [row for row in csv.reader(io.StringIO(r.stdout)) if len(row) > 10 and "k_trace_pool" in row[4]]
    m = {row[-3]: float(row[-1].replace(",", "")) for row in rows}
    if not m:
        print("no capture for", scene, r.stderr[-300:])
        continue
    res[scene] = {"kernel": rows[0][4].split("(")[0], "workload": what + ", 32 spp per launch",
                  "dram_read_bytes": int(m["dram__bytes_read.sum"]), "dram_write_bytes": int(m["dram__bytes_write.sum"]),
                  "gpu_time_ns": int(m["gpu__time_duration.sum"]), "l2_hit_pct": m["lts__t_sector_hit_rate.pct"],
                  "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, third launch of tools/profile_trace.py --scene %s --spp 32 (tools/capture_traffic.py)" % scene}
    print(scene, res[scene], flush=True)
json.dump(res, open(out_path, "w"), indent=1)
print("wrote", out_path)
