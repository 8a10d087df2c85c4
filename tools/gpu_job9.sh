mkdir -p gpurun_out
VRB200_ENCODE_LUT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_brick_encode_lut|k_range_xy" -s 4 -c 2 -o gpurun_out/prof_build -f python tools/gpu_build_timing.py 1024 > gpurun_out/prof_build.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_build.ncu-rep > gpurun_out/sum_build.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_build.ncu-rep k_brick_encode_lut 40 > gpurun_out/lines_encode.txt 2>&1
cat gpurun_out/sum_build.txt | head -120
