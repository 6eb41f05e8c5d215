#!/bin/bash
# one gpurun call: A/B of the single-/two-image kernels against the previous build, GPU suite, smoke, bench line, config tables
mkdir -p gpurun_out
if [ -f tools/bin/libwt_old.so ]; then
python tools/ab_compare.py tools/bin/libwt_old.so whitomo_b200/libwhitomo_b200.so 256,180,1 512,360,2 1024,720,1 2048,1800,1 2048,1800,2 2048,1800,1,exact 4096,2048,1 > gpurun_out/ab_pairlines.jsonl 2> gpurun_out/ab_pairlines.err; cut -c1-200 gpurun_out/ab_pairlines.jsonl
fi
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench_r2b.json; echo
python tools/config_table.py > gpurun_out/config_table_r2b.json 2> gpurun_out/config_table_r2b.err; echo "config rc=$?"
WT_POSITIONS=exact python tools/config_table.py --gpu-only > gpurun_out/config_table_r2b_exact.json 2>/dev/null; echo "config exact rc=$?"
