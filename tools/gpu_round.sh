#!/bin/bash
# one gpurun call: GPU test suite, full-size parity table, a short bench (logs under gpurun_out/)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -${PYTEST_TAIL:-40} gpurun_out/pytest_gpu.log
if [ -n "${PARITY_CONFIGS:-}" ]; then
timeout 1500 python tools/parity_fullsize.py --configs ${PARITY_CONFIGS} --out gpurun_out/parity_fullsize.json > gpurun_out/parity.log 2>&1; echo "parity rc=$?"
tail -5 gpurun_out/parity.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
