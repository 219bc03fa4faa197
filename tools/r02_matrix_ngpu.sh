#!/usr/bin/env bash
# Measurement matrix, N GPUs of one box (N = number of visible GPUs): cfg2 (headline), cfg3 (ResNet-50, BASELINE configs[2]: 8 GPUs) and
# cfg4 (Swin-t, configs[3]: 2/4/8 GPUs) with the peer-memory exchange, cfg2 also with the NCCL fp32 exchange; + a timeline.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run() {
  tag=$1; shift
  ( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline "$@" ) > gpurun_out/r02n${N}_$tag.json 2> gpurun_out/r02n${N}_$tag.err
  rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02n${N}_$tag.json")); c = d.get("exchange_check") or {}
    print("== N=$N $tag rc=$rc ms", round(d["ms_per_step"], 4), "samples/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["launch_mode"], "identical", c.get("grads_bit_identical_across_ranks"), "maxdiff", c.get("max_rel_diff_vs_mean_of_local_grads"))
except Exception as e:
    print("== N=$N $tag rc=$rc parse failed", e)
PY
  grep -i "capture failed\|Error" gpurun_out/r02n${N}_$tag.err | head -2 | cut -c1-200
}
run cfg2_peer --workload cfg2
run cfg2_nccl_f32 --workload cfg2 --exchange nccl --grad-compress none
run cfg4_peer --workload cfg4
if [ "$N" -ge 8 ]; then run cfg3_peer --workload cfg3; fi
( DMC_PROF_GPUS=$N timeout 300 python tools/prof_step_dp.py peer ) > gpurun_out/r02n${N}_timeline_peer.txt 2>&1
echo "timeline rc=$?"
( timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r02n${N}_n1.json 2> gpurun_out/r02n${N}_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02n${N}_n1.json")); print("== same box, 1 GPU: ms", round(d["ms_per_step"], 4), "samples/s", round(d["value"]))
PY
echo done
