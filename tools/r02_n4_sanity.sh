#!/usr/bin/env bash
mkdir -p gpurun_out
( timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02f4_cfg2_peer.json 2> gpurun_out/r02f4_cfg2_peer.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02f4_cfg2_peer.json")); c = d.get("exchange_check") or {}
    print("== N=4 cfg2 ms", round(d["ms_per_step"], 4), "samples/s", round(d["value"]), d["launch_mode"], "identical", c.get("grads_bit_identical_across_ranks"), "maxdiff", c.get("max_rel_diff_vs_mean_of_local_grads"), d["config"]["grad_allreduce"][:50])
except Exception as e:
    print("== parse failed", e)
PY
