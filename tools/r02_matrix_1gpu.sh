#!/usr/bin/env bash
# Measurement matrix, 1 GPU: every BASELINE.json config (cfg1..cfg5 incl. the cfg5 out_dim sweep 4096..262144), bf16-GEMM mode,
# plus the fp32 parity mode on cfg2.  One JSON line per run in gpurun_out/r02m_*.json; summary table printed at the end.
mkdir -p gpurun_out
run() {
  tag=$1; shift
  ( timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" ) > gpurun_out/r02m_$tag.json 2> gpurun_out/r02m_$tag.err
  echo "== $tag rc=$?"
}
run cfg1 --workload cfg1
run cfg2 --workload cfg2
run cfg3 --workload cfg3
run cfg4 --workload cfg4
for k in 4096 8192 16384 32768 65536 131072 262144; do
  run cfg5_K$k --workload cfg5 --out-dim $k
done
run cfg2_fp32 --workload cfg2 --mode fp32 --steps 10
python - <<'PY'
import glob, json
print("| run | ms/step | samples/s | e2e samples/s | % of step roofline | dominant op (frac of its bound) |")
print("|---|---|---|---|---|---|")
for f in sorted(glob.glob("gpurun_out/r02m_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print("|", f, "| failed:", e, "|"); continue
    r = d["roofline"]
    print(f"| {f.split('r02m_')[1][:-5]} | {d['ms_per_step']:.4f} | {d['value']:.0f} | {d['e2e']['value']:.0f} | {100 * r['step']['frac_of_step_roofline']:.1f} | {r['kernel']} ({r['frac']:.2f} {r['bound']}) |")
PY
