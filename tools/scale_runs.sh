#!/bin/bash
# Bench lines for the BASELINE.json configs other than configs[1] (which is the default bench.py run).
# Usage (on the GPU box): tools/scale_runs.sh OUTDIR [c3 c4 c5 ...]
out=${1:-gpurun_out}; shift
mkdir -p "$out"
what=${@:-c3 c4 c5}
for w in $what; do
  case $w in
    c3)  # configs[2]: 100 Mbp graph, mismatch sweep -m 0..5 (10 M reads per step stand for the 100 M)
      for m in 0 1 2 3 4 5; do
        extra="--no-cpu-baseline"; [ $m = 2 ] && extra=""
        python bench.py --genome 100000000 --errors $m --label "configs[2] (-m $m)" --steps 3 --warmup 3 --no-cli $extra \
          > "$out/bench_c3_m$m.json" 2> "$out/bench_c3_m$m.err"
      done ;;
    c4)  # configs[3]: same graph, exhaustive aligner, 2 % error
      python bench.py --genome 100000000 --mode exhaustive --sub-rate 0.02 --label "configs[3]" --steps 3 --warmup 3 --no-cli \
        > "$out/bench_c4.json" 2> "$out/bench_c4.err" ;;
    c5)  # configs[4]: 3 Gbp graph, index resident in HBM; one 10 M-read shard per step
      python bench.py --genome 3000000000 --label "configs[4] (one 10 M-read shard of the 1 B)" --steps 3 --warmup 3 \
        --no-cpu-baseline --no-cli > "$out/bench_c5.json" 2> "$out/bench_c5.err" ;;
    c5s)  # 1 Gbp variant of configs[4] (index beyond L2 as well), with the CPU baseline
      python bench.py --genome 1000000000 --label "configs[4] scaled to 1 Gbp" --steps 3 --warmup 3 --no-cli \
        > "$out/bench_c5s.json" 2> "$out/bench_c5s.err" ;;
  esac
done
tail -n 3 "$out"/bench_c*.err
cat "$out"/bench_c*.json
