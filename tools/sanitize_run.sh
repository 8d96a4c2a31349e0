#!/bin/bash
out=gpurun_out; mkdir -p $out
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py 30000 > $out/r02_sanitizer_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|^ok|=========.*(Invalid|Race|hazard|Barrier)" $out/r02_sanitizer_$tool.txt | head -8
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hbm_random tools/hbm_random.cu && /tmp/hbm_random 12 > $out/r02_hbm_random.txt; /tmp/hbm_random 1 >> $out/r02_hbm_random.txt; head -3 $out/r02_hbm_random.txt
