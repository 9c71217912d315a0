#!/usr/bin/env bash
# The default bench line and the single workloads, one JSON file each.  Usage (under gpurun): bash tools/gpu_bench_all.sh [tag]
tag="${1:-r02}"
out=gpurun_out
mkdir -p "$out"
timeout 600 python bench.py --steps 5 --warmup 3 > "$out/bench_default_${tag}.json" 2> "$out/bench_default.err"; echo "default rc=$?"
for wl in ${WORKLOADS:-c2tc c2r c1 c4 c5}; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 > "$out/bench_${wl}_${tag}.json" 2> "$out/bench_${wl}.err"; echo "$wl rc=$?"
done
