#!/bin/bash
out=gpurun_out; mkdir -p $out
nproc; free -g | head -2
( time python bench.py --impl reference --steps 5 --warmup 3 > $out/bench_r2g_ref.json 2> $out/bench_r2g_ref.err ) 2>&1 | grep real
tail -c 1500 $out/bench_r2g_ref.json; echo
( time python bench.py --steps 20 --warmup 5 > $out/bench_r2g.json 2> $out/bench_r2g.err ) 2>&1 | grep real
tail -5 $out/bench_r2g.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2g.json'))
print('value',d['value']/1e6,'e2e',d['e2e']['value']/1e6,{k:v['value']/1e6 for k,v in d['e2e']['forms'].items()})
print('kernel_ms',d['kernel_ms_per_pass'],'timed',d['timed_region_s'])
r=d['roofline']; print('roofline dominant',r['kernel'],r['frac'],r['model_frac'],r['traffic'])
for k,v in r['kernels'].items(): print(k,{x:v[x] for x in ('frac','model_frac','necessary_bytes_per_read','traffic','kernel_ms')})
print('cpu',d.get('cpu_baseline')); print('cli',d.get('cli')); print('outcome',d['outcome']); print('index',d['index'])
PY
