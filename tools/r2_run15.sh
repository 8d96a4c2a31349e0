#!/bin/bash
out=gpurun_out; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2_pytest15.log 2>&1 ) 2>&1 | grep real; tail -3 $out/r2_pytest15.log
pr() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 g',d['genome'],'mode',d['mode'],'scan',round(d['scan_ms'],3),'ext',round(d['ext_ms'],3))"; }
timeout 600 python tools/tune.py 10000000 3000000000 0 2> $out/r2_tune15.err | pr greedy
timeout 600 python tools/tune.py 10000000 5000000 0 2> $out/r2_tune15.err | pr greedy
timeout 600 python tools/tune.py 10000000 100000000 1 2> $out/r2_tune15.err | pr exhaustive
timeout 600 python tools/tune.py 10000000 3000000000 1 2> $out/r2_tune15.err | pr exhaustive
