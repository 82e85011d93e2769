#!/bin/bash
tag=${1:-x1}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 1500 python -m pytest tests/test_gpu_configs.py -m gpu -q 2>&1 | tail -60 | cut -c1-300 | tee $out/pytest.txt
