#!/bin/bash
# final 1-GPU evidence set of round 2: sanitizers on the kernels changed since the last run, then the profile set
tag=${1:-f1}
out=gpurun_out/$tag
mkdir -p $out
SUB='gravity_forces_match_reference or full_step_vs_oracle or prewalk or list_overflow or stack_overflow or inactive or block_timesteps_match_reference or mainloop_matches_reference or mirror or deferred or restock or tree'
echo "== memcheck"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mainloop.py -m gpu -x -q -k "$SUB" > $out/memcheck.log 2>&1; echo "rc=$?"; grep -E "passed|failed|ERROR SUMMARY" $out/memcheck.log | tail -3
echo "== racecheck"; timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(gravity_forces_match_reference and (fastmono or quad)) or full_step_vs_oracle or kdtree_matches_reference or sharded_tree_build" > $out/racecheck.log 2>&1; echo "rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" $out/racecheck.log | tail -3
bash scratch/gpu_profile_r2.sh $tag
