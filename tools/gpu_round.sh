#!/bin/bash
# one gpurun call: GPU test suite, full-size parity table, a short bench (logs under gpurun_out/)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/parity_fullsize.py --configs ${PARITY_CONFIGS:-C2,C3,C4} --out gpurun_out/parity_fullsize.json > gpurun_out/parity.log 2>&1; echo "parity rc=$?"
tail -5 gpurun_out/parity.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
