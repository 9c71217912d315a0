cat > /tmp/f16x3_one.py <<'PY'
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import torch, vo_b200
from vo_b200 import ops
import tc_perf
tc_perf.run(10000, 10000, 16, ops.VO_PREC_F16X3, ops.VO_METRIC_COSINE, reps=2)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:match_f32_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_match_f16x3_r01i python /tmp/f16x3_one.py > gpurun_out/ncu_f16x3.log 2>&1; tail -2 gpurun_out/ncu_f16x3.log
VO_TC_DEBUG=1 timeout 100 python /tmp/f16x3_one.py 2>&1 | tail -3 | cut -c1-400
