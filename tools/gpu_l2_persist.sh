# L2 access-policy window over records + majorant tables: Gsamples/s and L2 hit rate / DRAM bytes with 0 / 16 / 32 / 64 MiB persisting
mkdir -p gpurun_out
: > gpurun_out/l2_persist.log
for sc in c3 c4 c1; do for mb in 0 16 32 64; do
  timeout 300 python tools/profile_trace.py --scene $sc --spp 32 --launches 5 --l2-persist $mb --json 1 2>&1 | grep "JSON\|rror" >> gpurun_out/l2_persist.log
done; done
for sc in c3 c4; do for mb in 0 32; do
  timeout 600 ncu --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:k_trace_pool -s 2 -c 1 --csv python tools/profile_trace.py --scene $sc --spp 32 --launches 3 --l2-persist $mb 2>/dev/null | grep "k_trace_pool" | awk -F'","' -v tag="$sc l2_persist=$mb" '{print tag, $(NF-2), $(NF-1), $NF}' >> gpurun_out/l2_persist.log
done; done
cat gpurun_out/l2_persist.log | cut -c1-220
