#!/bin/bash
# N-GPU weak-scaling line of bench.py (one rank per GPU, index replicated, queries sharded, no data-path collective)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err | cut -c1-300; cut -c1-260 gpurun_out/bench_n$N.json
