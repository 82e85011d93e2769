#!/bin/bash
tag=${1:-x1}
out=gpurun_out/$tag
mkdir -p $out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/pytest.txt
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 0 2>/dev/null | tail -1 | cut -c1-300
