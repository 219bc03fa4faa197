#!/usr/bin/env bash
# A/B of the gradient exchange (fp32 vs bf16) -- 1 GPU: new tests + the reducer path on a 1-rank NCCL group;
# N >= 2 GPUs: 2-GPU tests + torchrun bench at N = number of visible GPUs.  Usage: tools/compress_ab.sh [tests] [bench]
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for s in ${*:-tests bench}; do
case $s in
tests)
  if [ "$NG" -ge 2 ]; then
    ( timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short -p no:cacheprovider ) > gpurun_out/test_compress_multi.log 2>&1
    echo "multi rc=$?"; tail -n 15 gpurun_out/test_compress_multi.log
  else
    ( timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py -q -m gpu --tb=short -p no:cacheprovider \
        -k "weightnorm or split_and_cast or bf16_gradient_exchange" ) > gpurun_out/test_compress.log 2>&1
    echo "tests rc=$?"; tail -n 30 gpurun_out/test_compress.log
  fi ;;
bench)
  for c in none bf16; do
    if [ "$NG" -ge 2 ]; then
      ( DMC_BENCH_GRAD_COMPRESS=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 \
          bench.py --gpus $NG --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/cmp_n${NG}_$c.json 2> gpurun_out/cmp_n${NG}_$c.err
      f=gpurun_out/cmp_n${NG}_$c
    else
      ( DMC_BENCH_FORCE_DP=1 DMC_BENCH_GRAD_COMPRESS=$c timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/cmp_dp1_$c.json 2> gpurun_out/cmp_dp1_$c.err
      f=gpurun_out/cmp_dp1_$c
    fi
    echo "== $f rc=$?"; python - <<PY
import json
try:
    d = json.load(open("$f.json")); print("ms", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["config"]["grad_allreduce"][:60])
except Exception as e:
    print("parse failed", e)
PY
    grep -v -i "warn" $f.err | tail -n 6
  done ;;
esac
done
