#!/usr/bin/env bash
# round 2, GPU call 16 (2 GPUs): late weight-norm backward + CTA count by buffer size
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short -p no:cacheprovider ) > gpurun_out/r02r_test_gpu_multi.log 2>&1
echo "== test_gpu_multi rc=$?"; tail -n 3 gpurun_out/r02r_test_gpu_multi.log | cut -c1-300
run() {
  tag=$1; shift
  ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline "$@" ) > gpurun_out/r02r_$tag.json 2> gpurun_out/r02r_$tag.err
  rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02r_$tag.json")); c = d.get("exchange_check") or {}
    print("== $tag rc=$rc ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["launch_mode"], "identical", c.get("grads_bit_identical_across_ranks"), "maxdiff", c.get("max_rel_diff_vs_mean_of_local_grads"))
except Exception as e:
    print("== $tag rc=$rc parse failed", e)
PY
  grep -i "capture failed" gpurun_out/r02r_$tag.err | head -2 | cut -c1-200
}
run peer
DMC_LATE_WN_BWD=0 run peer_earlywn
DMC_XRANK_BYTES_PER_CTA_LOG2=14 run peer_morectas
DMC_XRANK_BYTES_PER_CTA_LOG2=18 run peer_fewctas
( timeout 300 python tools/prof_step_dp.py ) > gpurun_out/r02r_prof_peer.txt 2>&1
echo done
