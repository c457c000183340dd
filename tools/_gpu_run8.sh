#!/bin/bash
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r02_bench_n${N}_f6.json 2> gpurun_out/r02_bench_n${N}_f6.err
grep '^{' gpurun_out/r02_bench_n${N}_f6.json | tail -c 400
tail -3 gpurun_out/r02_bench_n${N}_f6.err
