#!/usr/bin/env bash
# Short GPU-box visit after a change to the Hamming matcher: full parity suite, smoke, the headline workload (c2),
# its ncu launch list and one full capture of match_u8_kernel.  Usage (under gpurun): bash tools/gpu_round_c2.sh [tag]
tag="${1:-r01l}"
out=gpurun_out
mkdir -p "$out"
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider > "$out/pytest_gpu_${tag}.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu_${tag}.log"
tail -3 "$out/pytest_gpu_${tag}.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --workload c2 --steps 3 --warmup 3 > "$out/bench_c2_${tag}.json" 2> "$out/bench_c2.err"; echo "c2 rc=$?"
C2="python bench.py --steps 1 --warmup 3 --pairs 250 --unique 20 --no-cpu"
$C2 > "$out/plain_c2.log" 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_c2_${tag}.csv" $C2 > "$out/ncu_c2.log" 2>&1
$C2 > "$out/plain_c2b.log" 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:match_u8_kernel -s 3 -c 1 -f -o "$out/prof_match_u8_${tag}" $C2 > "$out/ncu_full_c2.log" 2>&1
ncu -i "$out/prof_match_u8_${tag}.ncu-rep" --page raw --csv > "$out/ncu_raw_match_u8_${tag}.csv" 2>/dev/null
ncu -i "$out/prof_match_u8_${tag}.ncu-rep" --page details --csv > "$out/ncu_details_match_u8_${tag}.csv" 2>/dev/null
python - "$out/bench_c2_${tag}.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print("c2 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2),
      "| binding", r.get("binding_pipe"), "| stages", {k:round(v,3) for k,v in d["stages_ms_per_launch"].items()})
PY
true
