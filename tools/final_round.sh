mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/profile_bench.sh r2 2>&1 | tail -5
python tools/config_table.py > gpurun_out/config_table_r2.json 2> gpurun_out/config_table_r2.err; echo "config rc=$?"
WT_POSITIONS=exact python tools/config_table.py --gpu-only > gpurun_out/config_table_r2_exact.json 2>/dev/null; echo "config exact rc=$?"
timeout 1500 python tools/parity_fullsize.py --configs C2,C3,C4,C5 --out gpurun_out/parity_fullsize.json > gpurun_out/parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/parity.log
