#!/bin/bash
tag=${1:-x1}
out=gpurun_out/$tag
mkdir -p $out
echo "== memcheck whole suite (small cases)"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mainloop.py tests/test_snapshot.py -m gpu -x -q -k "not million and not sphere_vs_oracle_tree" > $out/memcheck_all.log 2>&1; echo "rc=$?"; grep -E "passed|failed|ERROR SUMMARY" $out/memcheck_all.log | tail -3
