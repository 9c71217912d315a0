#!/usr/bin/env bash
# Multi-GPU scaling lines (builder-run; the driver computes its own SCALE file from the default workload).
# Usage under gpurun --gpus N:  bash tools/scale_run.sh N tag [workloads...]
N="$1"; tag="$2"; shift 2
out=gpurun_out; mkdir -p "$out"
port=29600
for wl in "${@:-default}"; do
  port=$((port + 1))
  VO_BENCH_WORKLOAD=$wl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus "$N" --steps 5 --warmup 3 > "$out/scale_${wl}_n${N}_${tag}.json" 2> "$out/scale_${wl}_n${N}.err"
  echo "$wl N=$N rc=$?"; tail -c 300 "$out/scale_${wl}_n${N}.err" | tail -2
done
port=$((port + 1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $port tools/h2d_wall.py 2>/dev/null | tail -1 > "$out/h2d_wall_n${N}_${tag}.json"
cat "$out/h2d_wall_n${N}_${tag}.json"
