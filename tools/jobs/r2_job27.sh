#!/bin/bash
# round 2, GPU job 27 (4 GPUs): the bench line under torchrun at N=4 and, on two of the GPUs, at N=2
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3 ) > gpurun_out/j27_bench_4gpu.json 2> gpurun_out/j27_bench_4gpu.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/j27_bench_2gpu.json 2> gpurun_out/j27_bench_2gpu.err
