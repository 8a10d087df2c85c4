# GPU-box job (gpurun): brick-build tests, timing of the 1024^3 build, per-kernel launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_brick.py tests/test_gpu_configs.py tests/test_nvdb.py -m gpu -q > gpurun_out/pytest_brick.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_brick.log
grep -v "^$" gpurun_out/pytest_brick.log | tail -8
timeout 300 python tools/gpu_build_timing.py 1024 2>&1 | tail -3

timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_range|k_scan|k_brick|k_make|k_linear" -c 60 --csv --log-file gpurun_out/launches_build.csv python tools/gpu_build_timing.py 1024 > gpurun_out/launches_build.log 2>&1
python tools/launch_table.py gpurun_out/launches_build.csv 5
