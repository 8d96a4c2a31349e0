#!/bin/bash
# per-function view of the greedy kernel (one ncu capture at 5 Mbp) + kernel times at 5 Mbp / 3 Gbp
out=gpurun_out; mkdir -p $out
pr() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 g',d['genome'],'mode',d['mode'],'scan',round(d['scan_ms'],3),'ext',round(d['ext_ms'],3))"; }
timeout 600 python tools/tune.py 10000000 5000000 0 2> /dev/null | pr greedy
timeout 600 python tools/tune.py 10000000 100000000 0 2> /dev/null | pr greedy
timeout 600 python tools/tune.py 10000000 3000000000 0 2> /dev/null | pr greedy
ncu --set full --clock-control none --import-source on -k regex:k_extend_greedy -s 3 -c 1 -f -o /tmp/prof_ext python tools/tune.py 10000000 5000000 0 > /dev/null 2>&1
python tools/ncu_lines.py /tmp/prof_ext.ncu-rep k_extend_greedyILi5 24 --by-function > $out/funcs_k_extend_greedy_5m.txt 2>&1
cat $out/funcs_k_extend_greedy_5m.txt
