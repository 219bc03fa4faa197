#!/usr/bin/env bash
# 2 GPUs: hierarchical barrier validation
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short -p no:cacheprovider ) > gpurun_out/r02s_test_gpu_multi.log 2>&1
echo "== test_gpu_multi rc=$?"; tail -n 3 gpurun_out/r02s_test_gpu_multi.log | cut -c1-300
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r02s_peer.json 2> gpurun_out/r02s_peer.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02s_peer.json")); c = d.get("exchange_check") or {}
    print("== peer ms", round(d["ms_per_step"], 4), d["launch_mode"], "identical", c.get("grads_bit_identical_across_ranks"), "maxdiff", c.get("max_rel_diff_vs_mean_of_local_grads"))
except Exception as e:
    print("== peer parse failed", e)
PY
echo done
