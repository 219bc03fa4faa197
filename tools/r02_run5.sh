#!/usr/bin/env bash
# round 2, GPU call 5 (2 GPUs): why the peer exchange fails under graph capture
mkdir -p gpurun_out
( timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/xrank_probe.py ) > gpurun_out/r02e_probe.txt 2>&1
echo "probe rc=$?"; grep -v -i "warn\|enable_symm" gpurun_out/r02e_probe.txt | tail -12 | cut -c1-300
run() {
  tag=$1; shift
  ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 30 --warmup 5 "$@" ) > gpurun_out/r02e_$tag.json 2> gpurun_out/r02e_$tag.err
  rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02e_$tag.json")); print("== $tag rc=$rc ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["launch_mode"], d.get("exchange_check"), d["config"]["grad_allreduce"][:50])
except Exception as e:
    print("== $tag rc=$rc parse failed", e)
PY
  grep -i "capture failed\|Error" gpurun_out/r02e_$tag.err | head -3 | cut -c1-300
}
run peer_eager --graph 0
DMC_REDUCER_PDL=0 run peer_nopdl
run peer
run nccl_f32_eager --exchange nccl --grad-compress none --graph 0
echo done
