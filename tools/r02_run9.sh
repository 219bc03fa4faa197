#!/usr/bin/env bash
# round 2, GPU call 9 (1 GPU): bounded teacher pass + GELU' saved in forward: tests, then A/B
mkdir -p gpurun_out
for f in test_gpu_modules test_gpu_baseline_dims test_gpu_kernels test_gpu_gemm test_gpu_dropin; do
  ( timeout 900 python -m pytest tests/$f.py -q -m gpu --tb=short -p no:cacheprovider ) > gpurun_out/r02i_$f.log 2>&1
  echo "== $f rc=$?"; tail -n 4 gpurun_out/r02i_$f.log | cut -c1-300
done
run() {
  tag=$1; shift
  ( timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" ) > gpurun_out/r02i_$tag.json 2> gpurun_out/r02i_$tag.err
  rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02i_$tag.json")); print("== $tag rc=$rc ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4))
except Exception as e:
    print("== $tag rc=$rc parse failed", e)
PY
}
run default
run default_again
DMC_GELU_DG=0 run nogeludg
DMC_BOUNDED_TEACHER_STATS=0 run nobounded
DMC_WN_AFTER_FIRST_GEMM=0 run wnfirst
( timeout 300 python tools/prof_step.py bf16 ) > gpurun_out/r02i_prof_default.txt 2>&1
echo done
