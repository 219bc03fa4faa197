#!/usr/bin/env bash
# round 2, GPU call 15 (2 GPUs): exchange kernel on reserved SMs vs shared SMs
mkdir -p gpurun_out
run() {
  tag=$1; shift
  ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline "$@" ) > gpurun_out/r02q_$tag.json 2> gpurun_out/r02q_$tag.err
  rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02q_$tag.json")); c = d.get("exchange_check") or {}
    print("== $tag rc=$rc ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["launch_mode"], "identical", c.get("grads_bit_identical_across_ranks"), "maxdiff", c.get("max_rel_diff_vs_mean_of_local_grads"))
except Exception as e:
    print("== $tag rc=$rc parse failed", e)
PY
  grep -i "capture failed" gpurun_out/r02q_$tag.err | head -2 | cut -c1-200
}
run c32_r16
DMC_XRANK_CTAS=16 DMC_XRANK_RESERVE_SMS=8 run c16_r8
DMC_XRANK_CTAS=16 DMC_XRANK_RESERVE_SMS=16 run c16_r16
DMC_XRANK_CTAS=48 DMC_XRANK_RESERVE_SMS=24 run c48_r24
DMC_XRANK_CTAS=148 DMC_XRANK_RESERVE_SMS=0 run c148_r0
DMC_XRANK_CTAS=32 DMC_XRANK_RESERVE_SMS=16 DMC_PEER_FLUSH_MB=64 run c32_r16_oneflush
( timeout 300 python tools/prof_step_dp.py ) > gpurun_out/r02q_prof_peer.txt 2>&1
echo done
