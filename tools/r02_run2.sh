#!/usr/bin/env bash
# round 2, GPU call 2: module tests with the new scheduling, then an A/B matrix of scheduling switches + timelines
mkdir -p gpurun_out
for f in test_gpu_modules test_gpu_baseline_dims; do
  ( timeout 900 python -m pytest tests/$f.py -q -m gpu --tb=line -p no:cacheprovider ) > gpurun_out/r02b_$f.log 2>&1
  echo "== $f rc=$?"; tail -n 12 gpurun_out/r02b_$f.log | cut -c1-300
done
run() {  # tag, bench args..., env via leading VAR=val handled by caller
  tag=$1; shift
  ( timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" ) > gpurun_out/r02b_$tag.json 2> gpurun_out/r02b_$tag.err
  rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02b_$tag.json")); print("== $tag rc=$rc ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["roofline"]["timing"][:12])
except Exception as e:
    print("== $tag rc=$rc parse failed", e)
PY
}
run default
run default_again
run overlap0 --overlap 0
DMC_EARLY_TEACHER_STATS=0 run overlap0_noearly --overlap 0
DMC_EARLY_TEACHER_STATS=0 run overlap1_noearly
DMC_DEFER_JOINS=0 run nodefer
DMC_WN_AFTER_FIRST_GEMM=0 run wnfirst
DMC_WGRAD_BF16=1 run wgrad_bf16
DMC_GEMM_SMEM_RESERVE_KB=8 run reserve8
DMC_GEMM_SMEM_RESERVE_KB=8 run reserve8_overlap0 --overlap 0
DMC_PDL=0 run nopdl
( timeout 300 python tools/prof_step.py bf16 ) > gpurun_out/r02b_prof_default.txt 2>&1
( DMC_BENCH_OVERLAP=0 timeout 300 python tools/prof_step.py bf16 ) > gpurun_out/r02b_prof_overlap0.txt 2>&1
echo done
