#!/bin/bash
# Everything profiles/ holds for one version tag, in one GPU-box call: ncu captures of both hot kernels at two
# graph sizes, the ncu launch list of the bench command, and the bench lines of the BASELINE.json configs.
# Usage (on the GPU box): tools/round_runs.sh TAG
tag=${1:-v9}; out=gpurun_out; mkdir -p $out
tools/gpu_profile.sh $tag 5000000 0 > $out/profile_$tag.log 2>&1
tools/gpu_profile.sh ${tag}_1g 1000000000 0 >> $out/profile_$tag.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cli --no-cpu-baseline > $out/launches_bench_$tag.log 2>&1
python bench.py --impl reference > $out/bench_reference_arm_$tag.json 2> $out/bench_reference_arm_$tag.err
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err
for m in 0 2 5; do
  extra="--no-cpu-baseline"; [ $m = 2 ] && extra=""
  python bench.py --genome 100000000 --errors $m --label "configs[2] (-m $m)" --steps 3 --warmup 3 --no-cli $extra \
    > $out/bench_c3_m${m}_$tag.json 2> $out/bench_c3_m${m}_$tag.err
done
python bench.py --genome 100000000 --mode exhaustive --sub-rate 0.02 --label "configs[3]" --steps 3 --warmup 3 --no-cli \
  > $out/bench_c4_$tag.json 2> $out/bench_c4_$tag.err
python bench.py --genome 1000000000 --label "configs[4] scaled to 1 Gbp" --steps 3 --warmup 3 --no-cli \
  > $out/bench_c5s_$tag.json 2> $out/bench_c5s_$tag.err
python bench.py --genome 3000000000 --label "configs[4] (one 10 M-read shard of the 1 B)" --steps 3 --warmup 3 \
  --no-cpu-baseline --no-cli > $out/bench_c5_3Gbp_$tag.json 2> $out/bench_c5_3Gbp_$tag.err
for f in $out/bench_*_$tag.json $out/bench_$tag.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print({k: d.get(k) for k in ("value", "ms_per_step")}, d.get("kernel_ms"), (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("upload"),
          (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"), d.get("cli"))
except Exception as e:
    print("unreadable:", e)
PY
done
