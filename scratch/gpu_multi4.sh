#!/bin/bash
mkdir -p gpurun_out/m4
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/m4/multi.log 2>&1
tail -3 gpurun_out/m4/multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/m4/bench4.log 2> gpurun_out/m4/bench4.err
python scratch/show_bench.py gpurun_out/m4/bench4.log | head -3
