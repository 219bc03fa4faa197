#!/usr/bin/env bash
# round 2, GPU call 12 (2 GPUs): exchange kernel with the max-smem carveout (co-residency with the dgrad), tests, timeline
mkdir -p gpurun_out
run() {
  tag=$1; shift
  ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline "$@" ) > gpurun_out/r02p_$tag.json 2> gpurun_out/r02p_$tag.err
  rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02p_$tag.json")); c = d.get("exchange_check") or {}
    print("== $tag rc=$rc ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["launch_mode"], "identical", c.get("grads_bit_identical_across_ranks"), "maxdiff", c.get("max_rel_diff_vs_mean_of_local_grads"), d["config"]["grad_allreduce"][:40])
except Exception as e:
    print("== $tag rc=$rc parse failed", e)
PY
  grep -i "capture failed" gpurun_out/r02p_$tag.err | head -2 | cut -c1-200
}
run peer
run peer2
( timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r02p_n1.json 2> gpurun_out/r02p_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02p_n1.json")); print("== n1 ms", round(d["ms_per_step"], 4))
PY
( timeout 300 python tools/prof_step_dp.py ) > gpurun_out/r02p_prof_peer.txt 2>&1
echo "prof rc=$?"
echo done
