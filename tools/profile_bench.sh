#!/bin/bash
# bench line, ncu launch list and ncu --set full captures of the hot kernels at the bench configuration
# usage (on the GPU box): bash tools/profile_bench.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
tail -c 400 gpurun_out/bench_${tag}.json; echo
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch_${tag}.log 2>&1
# timed steps 1-2: K2 (update epilogue), K1 (white epilogue), K2, K1  (synthesis + init + 3 warm-up steps are skipped)
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain2_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:fp_kernel|bp_kernel' -s 8 -c 4 \
    -f -o gpurun_out/prof_${tag} python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_${tag}.log 2>&1
ncu --set full --clock-control none -k 'regex:barrier_update_kernel' -s 3 -c 1 \
    -f -o gpurun_out/prof_${tag}_k3 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_${tag}_k3.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv
ncu -i gpurun_out/prof_${tag}_k3.ncu-rep --page raw --csv | tail -n +3 >> gpurun_out/prof_${tag}_raw.csv
ls -la gpurun_out/prof_${tag}*
