#!/bin/bash
out=gpurun_out; mkdir -p $out
( time timeout 1200 python -m pytest tests/test_gpu_large_graph.py -x -q > $out/r2_pytest8.log 2>&1 ) 2>&1 | grep real; tail -5 $out/r2_pytest8.log
( time timeout 1500 python tools/parity_large.py 3e9 200000 > $out/parity_3Gbp.json 2> $out/parity_3Gbp.err ) 2>&1 | grep real
cat $out/parity_3Gbp.json; tail -3 $out/parity_3Gbp.err
