#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/r2_pytest4.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2_pytest4.log
pr() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 g',d['genome'],'scan',round(d['scan_ms'],3),'ext',round(d['ext_ms'],3), d['index']['device_bytes'])"; }
for g in 100000000 1000000000 3000000000; do
  timeout 600 python tools/tune.py 10000000 $g 0 2> $out/r2_tune4.err | pr default
done
BGR_BLOOM_MB=128 timeout 600 python tools/tune.py 10000000 3000000000 0 2> $out/r2_tune4.err | pr bloom128MB
BGR_BLOOM_MB=32 timeout 600 python tools/tune.py 10000000 3000000000 0 2> $out/r2_tune4.err | pr bloom32MB
BGR_BLOOM_MB=64 timeout 600 python tools/tune.py 10000000 1000000000 0 2> $out/r2_tune4.err | pr bloom64MB
timeout 900 tools/gpu_profile.sh r2d_3g 3000000000 0 > $out/profile_r2d_3g.log 2>&1
grep -E "duration|thread_inst_executed_per|registers|dram__bytes|issue_active|warps_active|long_scoreboard|inst_executed.sum|lts__t_sectors.sum|lts__throughput|l1tex__throughput|pipe_alu|pipe_fma|lts__t_sector_hit" $out/ncu_k_scan_tiles_r2d_3g.txt
