#!/usr/bin/env bash
# N GPUs, trimmed: cfg2 / cfg4 / cfg3 with the peer-memory exchange + one timeline
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run() {
  tag=$1; shift
  ( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline "$@" ) > gpurun_out/r02f${N}_$tag.json 2> gpurun_out/r02f${N}_$tag.err
  rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02f${N}_$tag.json")); c = d.get("exchange_check") or {}
    print("== N=$N $tag rc=$rc ms", round(d["ms_per_step"], 4), "samples/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["launch_mode"], "identical", c.get("grads_bit_identical_across_ranks"), "maxdiff", c.get("max_rel_diff_vs_mean_of_local_grads"))
except Exception as e:
    print("== N=$N $tag rc=$rc parse failed", e)
PY
}
run cfg2_peer --workload cfg2
( DMC_PROF_GPUS=$N timeout 120 python tools/prof_step_dp.py peer ) > gpurun_out/r02f${N}_timeline_peer.txt 2>&1
run cfg4_peer --workload cfg4
if [ "$N" -ge 8 ]; then run cfg3_peer --workload cfg3; fi
echo done
