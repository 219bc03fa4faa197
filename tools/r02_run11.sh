#!/usr/bin/env bash
# round 2, GPU call 11 (1 GPU): max-smem carveout for every kernel (co-residency), bounded teacher statistics in the GEMM epilogue
mkdir -p gpurun_out
run() {
  tag=$1; shift
  ( timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" ) > gpurun_out/r02k_$tag.json 2> gpurun_out/r02k_$tag.err
  rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02k_$tag.json")); print("== $tag rc=$rc ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4))
except Exception as e:
    print("== $tag rc=$rc parse failed", e)
PY
}
run default
DMC_MAX_SMEM_CARVEOUT=0 run nocarveout
run default_again
DMC_TEACHER_EPILOGUE_STATS=1 run epistats
DMC_POLITE_CTAS=148 run polite148
DMC_POLITE_CTAS=296 run polite296
DMC_POLITE_CTAS=444 run polite444
run hiprio --hiprio 1
DMC_BOUNDED_TEACHER_STATS=0 run nobounded
( timeout 300 python tools/prof_step.py bf16 ) > gpurun_out/r02k_prof_default.txt 2>&1
( DMC_TEACHER_EPILOGUE_STATS=1 timeout 300 python tools/prof_step.py bf16 ) > gpurun_out/r02k_prof_epistats.txt 2>&1
( DMC_TEACHER_EPILOGUE_STATS=1 timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_baseline_dims.py -q -m gpu --tb=line -p no:cacheprovider ) > gpurun_out/r02k_tests_epistats.log 2>&1
echo "== tests with epilogue stats rc=$?"; tail -n 8 gpurun_out/r02k_tests_epistats.log | cut -c1-250
echo done
