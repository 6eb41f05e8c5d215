#!/bin/bash
# development aid: DRAM bytes and speed of K1 per launch order (WT_FP_PHASES) at C4, one float4 group and 8 slices
mkdir -p gpurun_out
for ph in 1 2 4; do
  WT_FP_PHASES=$ph python tools/microbench.py 2048,1800,4 2048,1800,16 > gpurun_out/order_${ph}.log 2>&1
  WT_FP_PHASES=$ph ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
     --clock-control none -k regex:fp_kernel -c 3 --csv --log-file gpurun_out/order_${ph}.csv python tools/microbench.py 2048,1800,4 > /dev/null 2>&1
done
cat gpurun_out/order_*.log
grep -h "dram__bytes\|gpu__time" gpurun_out/order_*.csv | cut -d, -f5,13-15
