#!/bin/bash
# Everything profiles/ holds for one version tag, in one GPU-box call: ncu captures of both hot kernels at two graph
# sizes, the ncu launch list of the bench command, both arms of the default bench (configs[4], 3 Gbp) and the bench
# lines of the other BASELINE.json configs.   Usage (on the GPU box): tools/round_runs.sh TAG
tag=${1:-r02}; out=gpurun_out; mkdir -p $out
tools/gpu_profile.sh ${tag}_3g 3000000000 0 > $out/profile_${tag}_3g.log 2>&1
tools/gpu_profile.sh ${tag}_5m 5000000 0 > $out/profile_${tag}_5m.log 2>&1
tools/gpu_profile.sh ${tag}_100m_exh 100000000 1 > $out/profile_${tag}_100m_exh.log 2>&1
# per-source-line view of the greedy kernel (needs the report and the library it was taken with: done here), then
# the reports go: gpurun brings back at most 64 MB
python tools/ncu_lines.py $out/prof_ext_${tag}_3g.ncu-rep k_extend_greedyILi5 60 > $out/lines_k_extend_greedy_${tag}_3g.txt 2>&1
python tools/ncu_lines.py $out/prof_ext_${tag}_5m.ncu-rep k_extend_greedyILi5 60 > $out/lines_k_extend_greedy_${tag}_5m.txt 2>&1
rm -f $out/*.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench_$tag.csv \
    python bench.py --steps 2 --warmup 1 --passes 2 --no-cli --no-cpu-baseline > $out/launches_bench_$tag.log 2>&1
python bench.py --impl reference --steps 5 --warmup 3 > $out/bench_reference_arm_$tag.json 2> $out/bench_reference_arm_$tag.err
python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err
python bench.py --genome 5000000 --label "configs[1]" --steps 5 --warmup 3 --cli-reads 10000000 > $out/bench_c1_$tag.json 2> $out/bench_c1_$tag.err
for m in 0 2 5; do
  extra="--no-cpu-baseline"; [ $m = 2 ] && extra=""
  python bench.py --genome 100000000 --errors $m --label "configs[2] (-m $m)" --steps 3 --warmup 3 --no-cli $extra \
    > $out/bench_c2_m${m}_$tag.json 2> $out/bench_c2_m${m}_$tag.err
done
python bench.py --genome 100000000 --mode exhaustive --sub-rate 0.02 --label "configs[3]" --steps 3 --warmup 3 --no-cli \
  > $out/bench_c3_$tag.json 2> $out/bench_c3_$tag.err
for f in $out/bench_$tag.json $out/bench_c*_$tag.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print({k: d.get(k) for k in ("value", "ms_per_step")}, d.get("kernel_ms_per_pass"), (d.get("e2e") or {}).get("value"),
          {k: (None if v["frac"] is None else round(v["frac"], 3)) for k, v in (d.get("roofline") or {}).get("kernels", {}).items()},
          (d.get("cpu_baseline") or {}).get("value"), (d.get("cli") or {}).get("reads_per_s_mapping"))
except Exception as e:
    print("unreadable:", e)
PY
done
