#!/usr/bin/env bash
# GPU-box visit C: parity tests, e2e dense vs sampled depth, keyframe-loop frames/s.  Usage: bash tools/gpu_round3.sh [tag]
tag="${1:-r01e}"
out=gpurun_out
mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > "$out/pytest_gpu_${tag}.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu_${tag}.log"
tail -15 "$out/pytest_gpu_${tag}.log"
for dm in sampled dense; do
  timeout 500 python bench.py --workload c2 --steps 5 --warmup 3 --e2e-depth $dm > "$out/bench_c2_${dm}_${tag}.json" 2> "$out/bench_c2_${dm}.err"; echo "c2 $dm rc=$?"
done
timeout 500 python bench.py --workload c3 --steps 3 --warmup 3 > "$out/bench_c3_${tag}.json" 2> "$out/bench_c3.err"; echo "c3 rc=$?"
timeout 300 python tools/seq_bench.py > "$out/seq_bench_${tag}.log" 2>&1; echo "seq rc=$?"; cat "$out/seq_bench_${tag}.log" | tail -5
for f in "$out"/bench_c2_sampled_${tag}.json "$out"/bench_c2_dense_${tag}.json "$out"/bench_c3_${tag}.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["e2e"].get("depth"), "h2d MB", round(d["e2e"]["h2d_bytes_per_step"]/1e6,1), "| stages", {k:round(v,3) for k,v in d["stages_ms_per_launch"].items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
for f in "$out"/bench_c2_sampled.err "$out"/bench_c2_dense.err; do [ -s $f ] && tail -5 $f; done
true
