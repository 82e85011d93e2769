#!/bin/bash
# compute-sanitizer over a subset of the GPU parity / main-loop tests (small particle counts)
tag=${1:-san2}
out=gpurun_out/$tag
mkdir -p $out
SUB='gravity_forces_match_reference or full_step_vs_oracle or prewalk or list_overflow or stack_overflow or inactive or block_timesteps_match_reference or mainloop_matches_reference or mirror'
echo "== memcheck"; timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mainloop.py -m gpu -x -q -k "$SUB" > $out/memcheck.log 2>&1; echo "rc=$?"; grep -E "passed|failed|ERROR SUMMARY" $out/memcheck.log | tail -3
echo "== racecheck"; timeout 2400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gravity_forces_match_reference or prewalk or full_step_vs_oracle" > $out/racecheck.log 2>&1; echo "rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" $out/racecheck.log | tail -3
