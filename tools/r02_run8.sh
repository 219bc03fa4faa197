#!/usr/bin/env bash
# round 2, GPU call 8 (1 GPU): tests for the fused normalize backward / BN head / executed drop-in, then scheduling A/B
mkdir -p gpurun_out
for f in test_gpu_modules test_gpu_dropin test_gpu_baseline_dims test_gpu_kernels test_gpu_gemm; do
  ( timeout 900 python -m pytest tests/$f.py -q -m gpu --tb=short -p no:cacheprovider ) > gpurun_out/r02h_$f.log 2>&1
  echo "== $f rc=$?"; tail -n 8 gpurun_out/r02h_$f.log | cut -c1-300
done
run() {
  tag=$1; shift
  ( timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" ) > gpurun_out/r02h_$tag.json 2> gpurun_out/r02h_$tag.err
  rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02h_$tag.json")); print("== $tag rc=$rc ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4))
except Exception as e:
    print("== $tag rc=$rc parse failed", e)
PY
}
run default
run default_again
DMC_DEFER_JOINS=0 run nodefer
DMC_FUSE_NORMALIZE_BWD=0 run nofuse_nbwd
run hiprio --hiprio 1
run hiprio_overlap0 --hiprio 1 --overlap 0
run overlap0 --overlap 0
DMC_EARLY_TEACHER_STATS=0 run noearly
( timeout 300 python tools/prof_step.py bf16 ) > gpurun_out/r02h_prof_default.txt 2>&1
echo done
