#!/bin/bash
# ncu full capture of K1/K2 at C4 (16 images) for two builds of the library (development aid)
set -e
mkdir -p gpurun_out
for tag in r1 new; do
  if [ $tag = r1 ]; then LIB=tools/bin/libwt_r1.so; else LIB=whitomo_b200/libwhitomo_b200.so; fi
  python tools/ab_kernels.py $LIB 2048,1800,16 > gpurun_out/plain_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k 'regex:fp_kernel|bp_kernel' -s 3 -c 5 -f -o gpurun_out/ab_$tag python tools/ab_kernels.py $LIB 2048,1800,16 > gpurun_out/ncu_$tag.log 2>&1
  tail -2 gpurun_out/ncu_$tag.log
done
