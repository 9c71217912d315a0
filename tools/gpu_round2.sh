#!/usr/bin/env bash
# Round-2 GPU visit: full parity suite, default bench (c3 headline + c2), single workloads, ncu launch list + full capture of the
# dominant kernels.  Usage (under gpurun): bash tools/gpu_round2.sh [tag]
tag="${1:-r02}"
out=gpurun_out
mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$out/pytest_gpu_${tag}.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu_${tag}.log"
tail -4 "$out/pytest_gpu_${tag}.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 > "$out/bench_default_${tag}.json" 2> "$out/bench_default.err"; echo "default rc=$?"
for wl in ${WORKLOADS:-c2tc c2r c1 c4 c5}; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 > "$out/bench_${wl}_${tag}.json" 2> "$out/bench_${wl}.err"; echo "$wl rc=$?"
done
C3="python bench.py --workload c3 --steps 1 --warmup 3 --passes 1 --no-cpu --allow-no-clocks"
$C3 > "$out/plain_c3.log" 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$out/launches_c3_${tag}.csv" $C3 > "$out/ncu_c3.log" 2>&1
$C3 > "$out/plain_c3b.log" 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:match_f32_tc_kernel -s 4 -c 1 -f -o "$out/prof_match_tc_${tag}" $C3 > "$out/ncu_full_c3.log" 2>&1
C2T="python bench.py --workload c2tc --steps 1 --warmup 3 --passes 1 --no-cpu --allow-no-clocks"
$C2T > "$out/plain_c2tc.log" 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:match_f32_tc_kernel -s 8 -c 1 -f -o "$out/prof_match_bits_${tag}" $C2T > "$out/ncu_full_c2tc.log" 2>&1
for f in "$out"/bench_*_${tag}.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2), "| roof", round(r["achieved"],2), r["unit"], "frac", round(r["frac"],4), "| clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
true
