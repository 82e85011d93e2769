#!/bin/bash
tag=${1:-m}
n=${2:-2}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 | tee $out/pytest.txt
for np_ in 16777216 67108864; do
echo "== bench $np_ x$n"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --particles $np_ --steps 2 --warmup 1 --no-cpu-baseline > $out/bench_${np_}_x$n.json 2> $out/bench_${np_}_x$n.err; python scratch/show_bench.py $out/bench_${np_}_x$n.json; grep -v "OMP_NUM\|^\*\*\*\|^$" $out/bench_${np_}_x$n.err | tail -4 | cut -c1-300
done
