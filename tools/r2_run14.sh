#!/bin/bash
out=gpurun_out; mkdir -p $out
pr() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 g',d['genome'],'mode',d['mode'],'scan',round(d['scan_ms'],3),'ext',round(d['ext_ms'],3))"; }
for lz in 24 28 32; do
  BGR_EXT_LAZY=$lz timeout 600 python tools/tune.py 10000000 3000000000 0 2> $out/r2_tune14.err | pr lazy$lz
  BGR_EXT_LAZY=$lz timeout 600 python tools/tune.py 10000000 5000000 0 2> $out/r2_tune14.err | pr lazy$lz
done
for lz in 8 24 32; do BGR_EXT_LAZY=$lz timeout 600 python tools/tune.py 10000000 100000000 1 2> $out/r2_tune14.err | pr exh_lazy$lz; done
timeout 600 tools/gpu_profile.sh r2i 5000000 0 > $out/profile_r2i.log 2>&1
grep -E "duration|thread_inst_executed_per|inst_executed.sum|issue_active|long_scoreboard" $out/ncu_k_extend_greedy_r2i.txt
python tools/ncu_lines.py $out/prof_ext_r2i.ncu-rep k_extend_greedyILi5 25 2>&1 | head -32
