#!/bin/bash
tag=${1:-x1}
out=gpurun_out/$tag
mkdir -p $out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/pytest.txt
