#!/bin/bash
# bench line, ncu launch list and one ncu --set full capture of the three hot kernels at the bench configuration
# usage (on the GPU box): bash tools/profile_bench.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
tail -c 600 gpurun_out/bench_${tag}.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --slices 4 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:fp_kernel|bp_kernel|barrier_update' -s 12 -c 3 \
    -f -o gpurun_out/prof_${tag} python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_${tag}.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv
