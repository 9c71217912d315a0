#!/usr/bin/env bash
# ncu --set full of the fp16 single-pass matcher at the c1 (2k x 2k, persistent CTAs) and c4 (20k x 20k, plain CTAs) shapes, each
# after the same command has run to completion without ncu.  Usage (under gpurun): bash tools/gpu_ncu_f16.sh [tag]
tag="${1:-r02}"
out=gpurun_out
mkdir -p "$out"
for wl in c1 c4; do
  CMD="python bench.py --workload $wl --steps 1 --warmup 3 --passes 1 --no-cpu --no-variants --no-api-rate --allow-no-clocks"
  $CMD > "$out/plain_${wl}.log" 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:match_f32_tc_kernel -s 6 -c 1 -f -o "$out/prof_match_f16_${wl}_${tag}" $CMD > "$out/ncu_full_${wl}.log" 2>&1
  echo "$wl ncu rc=$?"
  ncu -i "$out/prof_match_f16_${wl}_${tag}.ncu-rep" --page raw --csv > "$out/${tag}_ncu_raw_match_f16_${wl}.csv" 2>/dev/null
  ncu -i "$out/prof_match_f16_${wl}_${tag}.ncu-rep" --page details --csv > "$out/${tag}_ncu_details_match_f16_${wl}.csv" 2>/dev/null
  rm -f "$out/prof_match_f16_${wl}_${tag}.ncu-rep"
done
