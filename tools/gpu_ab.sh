timeout 600 python -m pytest tests/test_gpu_brick.py -x -q 2>&1 | tail -3
python tools/gpu_build_timing.py 1024
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_range|k_brick|k_scan|k_make|k_linear" -c 40 --csv --log-file gpurun_out/build_launches.csv python tools/gpu_build_timing.py 1024 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/build_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-39:]:
    print(r[4][:40].ljust(42), r[-3], r[-1])
PY
