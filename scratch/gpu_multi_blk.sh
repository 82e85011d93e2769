#!/bin/bash
mkdir -p gpurun_out/mb
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/mb/multi.log 2>&1
tail -5 gpurun_out/mb/multi.log
grep -a "ok:" gpurun_out/mb/multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/mb/bench2.log 2>&1
tail -2 gpurun_out/mb/bench2.log | cut -c1-1500
