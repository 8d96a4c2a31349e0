#!/bin/bash
# two GPUs: the multi-device command line, and the bench at N=2
out=gpurun_out; mkdir -p $out
nvidia-smi -L; nproc; free -g | head -2
timeout 600 python -m pytest tests/test_gpu_cli.py -x -q -k "several" > $out/r2_pytest11.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2_pytest11.log
BGR_PROBE_ENVS="BGR_GPUS=1" BGR_PROBE_ONLY_ENVS=1 timeout 900 python tools/cli_probe.py --reads 40000000 > $out/r2_cli_probe11.txt 2>&1
cat $out/r2_cli_probe11.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_r2h_2gpu.json 2> $out/bench_r2h_2gpu.err ) 2>&1 | grep real
tail -3 $out/bench_r2h_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2h_2gpu.json'))
print('N',d['n_gpus'],'value',d['value']/1e6,'e2e',d['e2e']['value']/1e6,{k:v['value']/1e6 for k,v in d['e2e']['forms'].items()})
print('kernel_ms',d['kernel_ms_per_pass'],'index',d['index'])
PY
