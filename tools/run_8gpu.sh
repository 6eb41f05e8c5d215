#!/bin/bash
# 8-GPU evidence run: bench.py under torchrun (slice-sharded C4 step + the C5 angle-split record) and the C4 stack job
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/c4_job.py --slices $((16*N)) --batch 8 --n-iter 6 --n-biter 3 --alpha 0.5 > gpurun_out/c4_job_${N}gpu.json 2> gpurun_out/c4_job_${N}gpu.err; echo "job rc=$?"
tail -c 300 gpurun_out/c4_job_${N}gpu.err
