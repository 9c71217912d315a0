#!/usr/bin/env bash
# GPU-box visit B: full-size parity tests, all five workloads, ncu captures of the PnP scoring and dense
# back-projection kernels.  Usage (under gpurun): bash tools/gpu_round2.sh [tag]
tag="${1:-r01d}"
out=gpurun_out
mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > "$out/pytest_gpu_${tag}.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu_${tag}.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke rc=$?" >> "$out/smoke.log"
for wl in c2 c3 c1 c4 c5; do
  timeout 500 python bench.py --workload $wl --steps 3 --warmup 3 > "$out/bench_${wl}_${tag}.json" 2> "$out/bench_${wl}.err"; echo "$wl rc=$?"
done
C2="python bench.py --steps 1 --warmup 3 --pairs 250 --unique 20 --no-cpu"
C3="python bench.py --workload c3 --steps 1 --warmup 3 --pairs 8 --unique 8 --chunk 8 --e2e-chunk 8 --no-cpu"
$C2 > "$out/plain_c2.log" 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_c2_${tag}.csv" $C2 > "$out/ncu_c2.log" 2>&1
$C2 > "$out/plain_c2b.log" 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 3 -c 1 -f -o "$out/prof_score_c2_${tag}" $C2 > "$out/ncu_full_score_c2.log" 2>&1
$C3 > "$out/plain_c3b.log" 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 3 -c 1 -f -o "$out/prof_score_c3_${tag}" $C3 > "$out/ncu_full_score_c3.log" 2>&1
$C2 > "$out/plain_c2c.log" 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:backproject_dense -s 1 -c 1 -f -o "$out/prof_dense_${tag}" $C2 > "$out/ncu_full_dense.log" 2>&1
$C2 > "$out/plain_c2d.log" 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:match_u8_kernel -s 3 -c 1 -f -o "$out/prof_match_u8_${tag}" $C2 > "$out/ncu_full_c2.log" 2>&1
tail -12 "$out/pytest_gpu_${tag}.log"; tail -2 "$out/smoke.log"
for f in "$out"/bench_c?_${tag}.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2), "| roof", round(r["achieved"],2), r["unit"], "frac", round(r["frac"],4), "share", round(r["share_of_step"] or 0,3), "| stages", {k:round(v,3) for k,v in d["stages_ms_per_launch"].items()})
    for k in d.get("kernels",[]): print("    ", k.get("kernel","")[:40], round(k.get("achieved",0),2), k.get("unit"), "frac", round(k.get("frac",0),4), k.get("error",""))
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1].replace(".json","").rsplit("_",1)[0]+".err").read()[-1500:] if False else "")
PY
done
for wl in c1 c2 c3 c4 c5; do [ -s "$out/bench_${wl}.err" ] && { echo "== $wl stderr"; tail -5 "$out/bench_${wl}.err"; }; done
true
