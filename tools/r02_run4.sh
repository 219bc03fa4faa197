#!/usr/bin/env bash
# round 2, GPU call 4 (2 GPUs): the peer-memory all-reduce kernel against NCCL, then the 2-GPU step with both transports
mkdir -p gpurun_out
nvidia-smi -L
( timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/xrank_probe.py ) > gpurun_out/r02d_probe.txt 2>&1
echo "probe rc=$?"; grep -v -i "warn" gpurun_out/r02d_probe.txt | tail -20
run() {
  tag=$1; shift
  ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 30 --warmup 5 "$@" ) > gpurun_out/r02d_$tag.json 2> gpurun_out/r02d_$tag.err
  rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02d_$tag.json")); print("== $tag rc=$rc ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d.get("exchange_check"), d["config"]["grad_allreduce"][:60])
except Exception as e:
    print("== $tag rc=$rc parse failed", e)
PY
  grep -v -i "warn" gpurun_out/r02d_$tag.err | tail -3
}
run peer
run nccl_bf16 --exchange nccl --grad-compress bf16
run nccl_f32 --exchange nccl --grad-compress none
( timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r02d_n1.json 2> gpurun_out/r02d_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02d_n1.json")); print("== n1 ms", round(d["ms_per_step"], 4))
PY
( timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_modules.py -q -m gpu --tb=line -p no:cacheprovider ) > gpurun_out/r02d_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/r02d_tests.log | cut -c1-250
echo done
