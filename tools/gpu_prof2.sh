# full ncu captures of the production tracking kernel (third launch, 16 spp), with per-line, per-opcode and pipe summaries
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -s 2 -c 1 -o gpurun_out/prof_tf python tools/profile_trace.py --tf 1 --spp 32 --launches 3 > gpurun_out/prof_tf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -s 2 -c 1 -o gpurun_out/prof_notf python tools/profile_trace.py --tf 0 --spp 32 --launches 3 > gpurun_out/prof_notf.log 2>&1
for v in tf notf; do
  python tools/ncu_summary.py gpurun_out/prof_$v.ncu-rep > gpurun_out/sum_$v.txt 2>&1
  python tools/ncu_opcodes.py gpurun_out/prof_$v.ncu-rep k_trace > gpurun_out/ops_$v.txt 2>&1
  python tools/ncu_lines.py gpurun_out/prof_$v.ncu-rep k_trace 70 > gpurun_out/lines_$v.txt 2>&1
done
grep -i "pipe\|issue_active\|thread_inst\|time_dur" gpurun_out/sum_notf.txt; head -30 gpurun_out/ops_notf.txt
