#!/usr/bin/env bash
# One GPU-box visit: parity tests, smoke, benches, then (only if the plain runs exit 0) ncu launch lists / captures.
# Usage (under gpurun): bash tools/gpu_round.sh [tag]
tag="${1:-r01}"
out=gpurun_out
mkdir -p "$out"
timeout 120 python tools/tc_debug.py > "$out/tc_debug.log" 2>&1
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke rc=$?" >> "$out/smoke.log"
timeout 400 python bench.py --steps 5 --warmup 3 > "$out/bench_c2_${tag}.json" 2> "$out/bench_c2.err"; echo "c2 rc=$?"
timeout 400 python bench.py --workload c3 --steps 3 --warmup 3 > "$out/bench_c3_${tag}.json" 2> "$out/bench_c3.err"; echo "c3 rc=$?"
timeout 400 python bench.py --workload c1 --steps 3 --warmup 3 > "$out/bench_c1_${tag}.json" 2> "$out/bench_c1.err"; echo "c1 rc=$?"
# --- ncu: launch lists (cold, serialised: shares only) and one full capture per dominant kernel
C2="python bench.py --steps 1 --warmup 3 --pairs 250 --unique 20 --no-cpu"
C3="python bench.py --workload c3 --steps 1 --warmup 3 --pairs 8 --unique 8 --chunk 8 --e2e-chunk 8 --no-cpu"
$C2 > "$out/plain_c2.log" 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_c2_${tag}.csv" $C2 > "$out/ncu_c2.log" 2>&1
$C3 > "$out/plain_c3.log" 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_c3_${tag}.csv" $C3 > "$out/ncu_c3.log" 2>&1
$C2 > "$out/plain_c2b.log" 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:match_u8_kernel -s 3 -c 1 -f -o "$out/prof_match_u8_${tag}" $C2 > "$out/ncu_full_c2.log" 2>&1
$C3 > "$out/plain_c3b.log" 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:match_f32_tc_kernel -s 3 -c 1 -f -o "$out/prof_match_tc_${tag}" $C3 > "$out/ncu_full_c3.log" 2>&1
tail -3 "$out/pytest_gpu.log"; cat "$out/tc_debug.log" | tail -4; cat "$out/smoke.log" | tail -2
for f in "$out"/bench_c?_${tag}.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2), "| roof", r["kernel"][:30], round(r["achieved"],2), r["unit"], "frac", round(r["frac"],4), "share", round(r["share_of_step"] or 0,3), "| stages", {k:round(v,3) for k,v in d["stages_ms_per_launch"].items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
ls -la "$out" | tail -30
