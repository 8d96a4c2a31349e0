#!/bin/bash
# round 2, first GPU call: GPU tests, kernel times at 5 Mbp / 3 Gbp, ncu of both hot kernels
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/r2_pytest1.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r2_pytest1.log
for g in 5000000 3000000000; do
  timeout 600 python tools/tune.py 10000000 $g 0 > $out/r2_tune_$g.json 2> $out/r2_tune_$g.err; cat $out/r2_tune_$g.json
done
timeout 600 tools/gpu_profile.sh r2a 5000000 0 > $out/profile_r2a.log 2>&1
timeout 900 tools/gpu_profile.sh r2a_3g 3000000000 0 > $out/profile_r2a_3g.log 2>&1
cat $out/ncu_k_extend_greedy_r2a.txt | grep -E "duration|thread_inst_executed_per|registers|dram__bytes|issue_active|warps_active|long_scoreboard"
cat $out/ncu_k_extend_greedy_r2a_3g.txt | grep -E "duration|thread_inst_executed_per|registers|dram__bytes|issue_active|warps_active|long_scoreboard"
