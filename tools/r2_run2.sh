#!/bin/bash
# round 2, second GPU call: parity of the CTA round machine, kernel times, ncu
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/r2_pytest2.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r2_pytest2.log
for g in 5000000 3000000000; do
  for minb in 3 4 2; do
    BGR_EXT_MINB=$minb timeout 600 python tools/tune.py 10000000 $g 0 2> $out/r2_tune2.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('g',d['genome'],'minb',d['ext_minb'],'scan',round(d['scan_ms'],3),'ext',round(d['ext_ms'],3))"
  done
  BGR_L2_FETCH=64 timeout 600 python tools/tune.py 10000000 $g 0 2> $out/r2_tune2.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('L2FETCH64 g',d['genome'],'scan',round(d['scan_ms'],3),'ext',round(d['ext_ms'],3))"
done
timeout 600 tools/gpu_profile.sh r2b 5000000 0 > $out/profile_r2b.log 2>&1
timeout 900 tools/gpu_profile.sh r2b_3g 3000000000 0 > $out/profile_r2b_3g.log 2>&1
for f in r2b r2b_3g; do grep -E "duration|thread_inst_executed_per|registers|dram__bytes|issue_active|warps_active|long_scoreboard|barrier|inst_executed.sum|lts__t_sectors.sum" $out/ncu_k_extend_greedy_$f.txt; done
