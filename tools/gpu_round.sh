#!/usr/bin/env bash
# GPU-box visit: full parity suite, all five workloads, front-end benches, ncu launch list + full captures of the
# dominant kernels.  Usage (under gpurun): bash tools/gpu_round.sh [tag]
tag="${1:-r01k}"
out=gpurun_out
mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$out/pytest_gpu_${tag}.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu_${tag}.log"
tail -6 "$out/pytest_gpu_${tag}.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python tools/orb_bench.py > "$out/orb_bench_${tag}.json" 2> "$out/orb_bench.err"; cut -c1-400 "$out/orb_bench_${tag}.json"
timeout 300 python tools/sift_bench.py > "$out/sift_bench_${tag}.json" 2> "$out/sift_bench.err"; cut -c1-400 "$out/sift_bench_${tag}.json"
for wl in c2 c2r c3 c1 c4 c5; do
  timeout 500 python bench.py --workload $wl --steps 3 --warmup 3 > "$out/bench_${wl}_${tag}.json" 2> "$out/bench_${wl}.err"; echo "$wl rc=$?"
done
timeout 200 python tools/r2d2_bench.py > "$out/r2d2_bench_${tag}.json" 2>/dev/null; timeout 200 python tools/r2d2_e2e.py > "$out/r2d2_e2e_${tag}.json" 2>/dev/null
timeout 300 python tools/seq_bench.py > "$out/seq_bench_${tag}.log" 2>&1
C2="python bench.py --steps 1 --warmup 3 --pairs 250 --unique 20 --no-cpu"
$C2 > "$out/plain_c2.log" 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_c2_${tag}.csv" $C2 > "$out/ncu_c2.log" 2>&1
$C2 > "$out/plain_c2b.log" 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:match_u8_kernel -s 3 -c 1 -f -o "$out/prof_match_u8_${tag}" $C2 > "$out/ncu_full_c2.log" 2>&1
for f in "$out"/bench_c?_${tag}.json "$out"/bench_c2r_${tag}.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2), "| roof", round(r["achieved"],2), r["unit"], "frac", round(r["frac"],4), "| stages", {k:round(v,3) for k,v in d["stages_ms_per_launch"].items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
cat "$out/r2d2_bench_${tag}.json" | cut -c1-300; cat "$out/r2d2_e2e_${tag}.json"; tail -4 "$out/seq_bench_${tag}.log" | cut -c1-300
true
