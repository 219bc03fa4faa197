#!/usr/bin/env bash
# round 2, GPU call 3: fp32 accumulator banks (parity at BASELINE dims), polite streaming kernels, teacher epilogue statistics
mkdir -p gpurun_out
for f in test_gpu_baseline_dims test_gpu_modules test_gpu_gemm test_gpu_kernels; do
  ( timeout 900 python -m pytest tests/$f.py -q -m gpu --tb=line -p no:cacheprovider ) > gpurun_out/r02c_$f.log 2>&1
  echo "== $f rc=$?"; tail -n 12 gpurun_out/r02c_$f.log | cut -c1-300
done
run() {
  tag=$1; shift
  ( timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" ) > gpurun_out/r02c_$tag.json 2> gpurun_out/r02c_$tag.err
  rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02c_$tag.json")); print("== $tag rc=$rc ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["roofline"]["timing"][:12])
except Exception as e:
    print("== $tag rc=$rc parse failed", e)
PY
}
run default
run default_again
DMC_POLITE_CTAS=0 run impolite
DMC_POLITE_CTAS=296 run polite296
DMC_TEACHER_EPILOGUE_STATS=0 run noepistats
DMC_DEFER_JOINS=0 run nodefer
DMC_WGRAD_BF16=0 run wgrad_f32
run overlap0 --overlap 0
DMC_WN_AFTER_FIRST_GEMM=0 run wnfirst
run fp32 --mode fp32 --steps 10
( timeout 300 python tools/prof_step.py bf16 ) > gpurun_out/r02c_prof_default.txt 2>&1
( DMC_BENCH_OVERLAP=0 timeout 300 python tools/prof_step.py bf16 ) > gpurun_out/r02c_prof_overlap0.txt 2>&1
echo done
