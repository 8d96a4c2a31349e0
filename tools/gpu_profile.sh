#!/bin/bash
# ncu --set full captures of the two hot kernels (one launch each, after warm-up launches), summarised
# into profiles/. Usage (on the GPU box): tools/gpu_profile.sh TAG [GENOME_BP] [MODE]
tag=${1:-v5}; genome=${2:-5000000}; mode=${3:-0}
out=gpurun_out; mkdir -p $out
scan=k_scan_tiles; ext=k_extend_greedy; [ "$mode" != 0 ] && ext=k_extend_exhaustive
ncu --set full --clock-control none --import-source on -k regex:$scan -s 3 -c 1 -f -o $out/prof_scan_$tag \
    python tools/tune.py 10000000 $genome $mode > $out/prof_scan_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$ext -s 3 -c 1 -f -o $out/prof_ext_$tag \
    python tools/tune.py 10000000 $genome $mode > $out/prof_ext_$tag.log 2>&1
python tools/ncu_summary.py $out/prof_scan_$tag.ncu-rep $out/ncu_${scan}_$tag.txt --traffic-key "k_scan_tiles@$genome" --traffic-json $out/roofline_traffic.json
python tools/ncu_summary.py $out/prof_ext_$tag.ncu-rep $out/ncu_${ext}_$tag.txt --traffic-key "$ext@$genome" --traffic-json $out/roofline_traffic.json
tail -2 $out/prof_scan_$tag.log
