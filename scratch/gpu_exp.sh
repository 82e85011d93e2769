#!/bin/bash
tag=${1:-x1}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 1500 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -25 | cut -c1-250 | tee $out/pytest.txt
