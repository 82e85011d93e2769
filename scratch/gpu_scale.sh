#!/bin/bash
# scaling sanity: bench at N GPUs (64M), as the driver launches it
tag=${1:-s}
n=${2:-4}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L | wc -l
if [ "$n" == "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 2 --warmup 1 --no-cpu-baseline > $out/bench_x1.json 2> $out/bench_x1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 3 --warmup 3 > $out/bench_x$n.json 2> $out/bench_x$n.err
fi
python scratch/show_bench.py $out/bench_x$n.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $out/bench_x$n.err | tail -5 | cut -c1-300
