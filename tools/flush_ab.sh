#!/usr/bin/env bash
# N-GPU A/B of the small-gradient flush threshold of GradAllReduce (MB); usage: tools/flush_ab.sh 1 64 ...
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for mb in "$@"; do
  f=gpurun_out/flush_n${NG}_$mb
  ( DMC_REDUCER_FLUSH_MB=$mb timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29543 \
      bench.py --gpus $NG --steps 30 --warmup 5 --no-cpu-baseline ) > $f.json 2> $f.err
  echo "== flush ${mb} MB rc=$?"; python - <<PY
import json
try:
    d = json.load(open("$f.json")); print("ms", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("parse failed", e)
PY
done
