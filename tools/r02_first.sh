#!/usr/bin/env bash
# First GPU call of the next round (1 GPU, ~3 min of box time): confirm the tree, then the A/B runs and captures that
# round 1 could no longer afford.  Everything lands in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 600 -- tools/r02_first.sh
mkdir -p gpurun_out
tools/gpu_round.sh tests smoke
# experimental tests (switches that are off by default)
( DMC_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_modules.py -q -m gpu --tb=short -p no:cacheprovider -k experimental ) > gpurun_out/test_experimental.log 2>&1
echo "experimental rc=$?"; tail -n 5 gpurun_out/test_experimental.log
run() {  # tag, env...
  tag=$1; shift
  ( env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r02_$tag.json 2> gpurun_out/r02_$tag.err
  echo "== $tag rc=$?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_$tag.json")); print("ms", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("parse failed", e)
PY
}
run default DMC_NOP=1
run wgrad_bf16 DMC_WGRAD_BF16=1          # dW stored bf16 between wgrad and weight-norm backward (-64 MB / step)
run default2 DMC_NOP=1                   # run-to-run spread
for n in 296 592 2368; do                # teacher statistics pass: CTA count (default 1184 = 8 per SM); it is latency-bound (26 % DRAM)
  run teacher_ctas_$n DMC_TEACHER_TARGET_CTAS=$n
done
# where the MLP backward goes: one ncu full capture of the layer-2 dgrad (GELU' epilogue) and the layer-3 shapes
tools/ncu_src.sh mlp_dgrad2 3 python tools/gemm_bench.py mlp_dgrad2
