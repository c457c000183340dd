timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r01h_bench_n2.json 2> gpurun_out/r01h_bench_n2.err; echo rc=$?
tail -c 400 gpurun_out/r01h_bench_n2.err
