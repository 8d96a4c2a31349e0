#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_gpu_seam.py tests/test_gpu_bench_workload.py -x -q > $out/r2_pytest10.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r2_pytest10.log
BGR_PROBE_ENVS="BGR_TEXT=0;BGR_GPUS=2;BGR_RANGE_BYTES=33554432;BGR_RANGE_BYTES=67108864" BGR_PROBE_ONLY_ENVS=1 timeout 900 python tools/cli_probe.py --reads 40000000 > $out/r2_cli_probe10.txt 2>&1
cat $out/r2_cli_probe10.txt
