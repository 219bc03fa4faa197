# Structural cost of the data-parallel step machinery on ONE GPU (1-rank NCCL group: the all-reduces move nothing).
run() { tag=$1; shift; ( env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ${EXTRA:-} ) > gpurun_out/dp_$tag.json 2> gpurun_out/dp_$tag.err; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/dp_$tag.json") if l.startswith("{")][-1]
    print("$tag", round(d["ms_per_step"],4), round(d["value"]))
except Exception as e: print("$tag failed", e); print(open("gpurun_out/dp_$tag.err").read()[-600:])
PY
}
run plain A=1
run dp1 DMC_BENCH_FORCE_DP=1
EXTRA="--reserve-sms 0" run dp1_reserve0 DMC_BENCH_FORCE_DP=1
run dp1_skipall DMC_BENCH_FORCE_DP=1 DMC_REDUCER_SKIP=all
