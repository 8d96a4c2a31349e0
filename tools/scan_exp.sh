#!/bin/bash
# Experiment: what the index loads make HBM deliver at 3 Gbp. Variants of the library built with -DBGR_LD_EXP=1
# (key buckets / entry sides loaded with .L2::64B) / =3 (the filter blocks too), and the stock library under
# cudaLimitMaxL2FetchGranularity 32 / 128: kernel times, and a narrow ncu pass (DRAM bytes, L2 sectors) of both kernels.
out=gpurun_out; mkdir -p $out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_ltcfabric.sum"
for v in base ld1 ld3 gran32 gran128; do
  lib=""; gran=""
  case $v in ld*) lib=$PWD/gpurun_exp/libexp_$v.so;; gran*) gran=${v#gran};; esac
  BGR_TUNE_LIB=$lib BGR_TUNE_L2GRAN=$gran python tools/tune.py 10000000 3000000000 0 2>$out/scanexp_$v.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant $v', 'scan', round(d['scan_ms'],3), 'ext', round(d['ext_ms'],3), d['counters']['aligned'])"
  BGR_TUNE_LIB=$lib BGR_TUNE_L2GRAN=$gran ncu --metrics $M --clock-control none -k regex:'k_scan_tiles|k_extend_greedy' -s 6 -c 2 --csv --log-file $out/scanexp_ncu_$v.csv python tools/tune.py 10000000 3000000000 0 > /dev/null 2>&1
  python - $out/scanexp_ncu_$v.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; i_n = h.index("Metric Name"); i_v = h.index("Metric Value"); i_k = h.index("Kernel Name")
for r in rows[1:]:
    print("   ", r[i_k][:24], r[i_n], r[i_v])
PY
done
