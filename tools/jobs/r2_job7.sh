#!/bin/bash
# round 2, GPU job 7 (2 GPUs): the bench line under torchrun at N=2 (C4 ensemble, shared rain upload over NCCL), and the
# C2 end-to-end path standalone with phase timing
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 ) > gpurun_out/j7_bench_2gpu.json 2> gpurun_out/j7_bench_2gpu.err
( timeout 600 python bench.py --workload c2 --no-extras --steps 5 --warmup 3 ) > gpurun_out/j7_bench_c2.json 2> gpurun_out/j7_bench_c2.err
( WMF_TIMING=1 timeout 600 python bench.py --workload c2 --no-extras --steps 2 --warmup 1 ) > gpurun_out/j7_bench_c2_timing.json 2> gpurun_out/j7_bench_c2_timing.err
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 --workload c2 ) > gpurun_out/j7_bench_c2_2gpu.json 2> gpurun_out/j7_bench_c2_2gpu.err
ls -la gpurun_out > gpurun_out/j7_ls.txt
