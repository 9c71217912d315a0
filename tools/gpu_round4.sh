#!/usr/bin/env bash
# GPU-box visit D: parity after the single-pass K-extension, SIFT workloads, keyframe-loop profile.
tag="${1:-r01f}"
out=gpurun_out
mkdir -p "$out"
timeout 120 python tools/tc_debug.py > "$out/tc_debug.log" 2>&1; tail -4 "$out/tc_debug.log"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$out/pytest_gpu_${tag}.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu_${tag}.log"
tail -15 "$out/pytest_gpu_${tag}.log"
for wl in c1 c4 c3; do
  timeout 500 python bench.py --workload $wl --steps 3 --warmup 3 > "$out/bench_${wl}_${tag}.json" 2> "$out/bench_${wl}.err"; echo "$wl rc=$?"
done
timeout 300 python tools/seq_bench.py > "$out/seq_bench_${tag}.log" 2>&1; echo "seq rc=$?"; tail -6 "$out/seq_bench_${tag}.log"
for f in "$out"/bench_c?_${tag}.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "| roof", round(r["achieved"],2), r["unit"], "frac", round(r["frac"],4), "| stages", {k:round(v,3) for k,v in d["stages_ms_per_launch"].items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
for wl in c1 c3 c4; do [ -s "$out/bench_${wl}.err" ] && { echo "== $wl stderr"; tail -5 "$out/bench_${wl}.err"; }; done
true
