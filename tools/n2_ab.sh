# A/B of the 2-GPU step (run under gpurun --gpus 2): reducer variants, same box.
run() { tag=$1; shift; ( env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/n2_$tag.json 2> gpurun_out/n2_$tag.err; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/n2_$tag.json") if l.startswith("{")][-1]
    print("$tag", round(d["ms_per_step"],4), round(d["value"]), round(d["e2e"]["value"]))
except Exception as e: print("$tag failed", e)
PY
}
run default A=1
run pdl DMC_REDUCER_PDL=1
run auxwn DMC_REDUCER_AUXWN=1
run pdl_auxwn DMC_REDUCER_PDL=1 DMC_REDUCER_AUXWN=1
