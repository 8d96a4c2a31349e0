#!/bin/bash
out=gpurun_out; mkdir -p $out
nvidia-smi -L | wc -l; nproc; free -g | head -2
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > $out/bench_r2h_8gpu.json 2> $out/bench_r2h_8gpu.err ) 2>&1 | grep real
tail -3 $out/bench_r2h_8gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2h_8gpu.json'))
print('N',d['n_gpus'],'value',d['value']/1e6,'e2e',d['e2e']['value']/1e6,{k:v['value']/1e6 for k,v in d['e2e']['forms'].items()})
print('kernel_ms',d['kernel_ms_per_pass'],'index',d['index'], d.get('rank0_numa_binding'))
PY
BGR_PROBE_ENVS="BGR_GPUS=1" BGR_PROBE_ONLY_ENVS=1 timeout 600 python tools/cli_probe.py --reads 40000000 2>&1 | grep -E "====|mapping done|index built|busy"
