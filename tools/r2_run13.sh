#!/bin/bash
out=gpurun_out; mkdir -p $out
python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print('L2', p.L2_cache_size)
PY
pr() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 g',d['genome'],'scan',round(d['scan_ms'],3),'ext',round(d['ext_ms'],3), d['index']['device_bytes'])"; }
for g in 100000000 1000000000 3000000000; do
  timeout 600 python tools/tune.py 10000000 $g 0 2> $out/r2_tune13.err | pr persist
  BGR_L2_PERSIST=0 timeout 600 python tools/tune.py 10000000 $g 0 2> $out/r2_tune13.err | pr nopersist
done
BGR_BLOOM_MB=32 timeout 600 python tools/tune.py 10000000 3000000000 0 2> $out/r2_tune13.err | pr persist_bloom32
BGR_BLOOM_MB=48 timeout 600 python tools/tune.py 10000000 3000000000 0 2> $out/r2_tune13.err | pr persist_bloom48
for lz in 4 16 24; do BGR_EXT_LAZY=$lz timeout 600 python tools/tune.py 10000000 3000000000 0 2> $out/r2_tune13.err | pr lazy$lz; done
for lz in 4 16 24; do BGR_EXT_LAZY=$lz timeout 600 python tools/tune.py 10000000 5000000 0 2> $out/r2_tune13.err | pr lazy$lz; done
