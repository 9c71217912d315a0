#!/usr/bin/env bash
# compute-sanitizer over every kernel family at small shapes (tools/sanitize_cases.py).  Run on the GPU box:
#   bash tools/sanitize.sh [tag]      -> gpurun_out/sanitizer/<tool>_<group>_<tag>.log + summary_<tag>.txt
# memcheck / synccheck / initcheck / racecheck; a tool that reports an error leaves a non-zero exit code in the summary.
tag="${1:-r02}"
out=gpurun_out/sanitizer
mkdir -p "$out"
CS=/usr/local/cuda/bin/compute-sanitizer
groups="${GROUPS_OVERRIDE:-match_f32 match_u8 pipeline geometry seq orb sift conv}"
tools="${TOOLS_OVERRIDE:-memcheck synccheck racecheck initcheck}"
: > "$out/summary_${tag}.txt"
for tool in $tools; do
  for grp in $groups; do
    log="$out/${tool}_${grp}_${tag}.log"
    extra=""
    [ "$tool" = initcheck ] && extra="--track-unused-memory no"
    timeout 900 $CS --tool $tool $extra --error-exitcode 66 --print-limit 20 python tools/sanitize_cases.py $grp > "$log" 2>&1
    rc=$?
    errs=$(grep -c "^========= \(Invalid\|Race\|Uninitialized\|Barrier\|Error\|Hazard\|Potential\|Program hit\)" "$log")
    echo "$tool $grp rc=$rc reports=$errs : $(grep -h 'ERROR SUMMARY\|RACECHECK SUMMARY' "$log" | tail -1)" | tee -a "$out/summary_${tag}.txt"
  done
done
