# ncu capture of the two-rays-per-lane kernel (third launch, 32 spp)
mkdir -p gpurun_out
for v in 1 0; do
  n=$([ $v = 1 ] && echo tf || echo notf)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_duo -s 2 -c 1 -o gpurun_out/prof_duo_$n python tools/profile_trace.py --tf $v --spp 32 --launches 3 --kernel 3 > gpurun_out/prof_duo_$n.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_duo_$n.ncu-rep > gpurun_out/sum_duo_$n.txt 2>&1
  python tools/ncu_lines.py gpurun_out/prof_duo_$n.ncu-rep k_trace 60 > gpurun_out/lines_duo_$n.txt 2>&1
done
grep -h "time_dur\|inst_executed.sum\|issue_active\|thread_inst_executed_per\|warps_active\|pipe_alu.avg\|pipe_xu.avg\|pipe_lsu.avg\|registers" gpurun_out/sum_duo_tf.txt gpurun_out/sum_duo_notf.txt
