#!/usr/bin/env bash
# One gpurun call: diagnostics + the GPU test suite + bench, each step bounded by its own timeout.
# Everything worth reading lands in gpurun_out/.  Usage: tools/gpu_round.sh [diag] [tests] [smoke] [bench] [ncu]
mkdir -p gpurun_out
STEPS="${*:-tests smoke bench}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
for s in $STEPS; do
case $s in
diag)
  ( timeout 900 python tools/diag_gemm.py ) > gpurun_out/diag_gemm.log 2>&1
  echo "diag rc=$?" >> gpurun_out/diag_gemm.log; tail -n 40 gpurun_out/diag_gemm.log ;;
tests)
  for f in test_gpu_gemm test_gpu_kernels test_gpu_modules; do
    ( timeout 900 python -m pytest tests/$f.py -q -m gpu --tb=short -p no:cacheprovider ) > gpurun_out/$f.log 2>&1
    echo "$f rc=$?" >> gpurun_out/$f.log
    echo "== $f"; tail -n 12 gpurun_out/$f.log
  done ;;
smoke)
  ( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1
  echo "smoke rc=$?" >> gpurun_out/smoke.log; echo "== smoke"; tail -n 6 gpurun_out/smoke.log ;;
bench)
  ( timeout 900 python bench.py --steps 30 --warmup 5 ) > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
  echo "== bench bf16 graph rc=$?"; tail -c 6000 gpurun_out/bench_bf16.json; tail -n 5 gpurun_out/bench_bf16.err
  ( timeout 600 python bench.py --steps 30 --warmup 5 --graph 0 --no-cpu-baseline ) > gpurun_out/bench_bf16_eager.json 2> gpurun_out/bench_bf16_eager.err
  echo "== bench bf16 eager rc=$?"; tail -c 3000 gpurun_out/bench_bf16_eager.json; tail -n 5 gpurun_out/bench_bf16_eager.err
  ( timeout 600 python bench.py --steps 10 --warmup 3 --mode fp32 --no-cpu-baseline ) > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
  echo "== bench fp32 rc=$?"; tail -c 3000 gpurun_out/bench_fp32.json; tail -n 5 gpurun_out/bench_fp32.err
  ( timeout 600 python bench.py --impl reference --steps 5 --warmup 2 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "== bench reference rc=$?"; tail -c 2000 gpurun_out/bench_ref.json; tail -n 5 gpurun_out/bench_ref.err ;;
ab)
  for fl in 0 1 2 3; do
    ( DMC_GEMM_FLAGS=$fl timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/bench_flags$fl.json 2> gpurun_out/bench_flags$fl.err
    echo "== bench flags=$fl rc=$?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_flags$fl.json"))
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]))
    for k in d["roofline"]["kernels"]:
        print("   %-26s %.4f ms  x%.0f %s" % (k["kernel"], k["ms_per_step"], k["calls_per_step"], ("%.0f GB/s" % k["GBps"]) if "GBps" in k else ""))
except Exception as e:
    print("parse failed", e)
PY
  done ;;
multi)
  NG=$(nvidia-smi -L | wc -l)
  for n in 2 4 8; do
    if [ $n -le $NG ]; then
      ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 20 --warmup 5 ) > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
      echo "== bench N=$n rc=$?"; tail -c 1500 gpurun_out/bench_n$n.json; grep -v Warning gpurun_out/bench_n$n.err | tail -n 8
    fi
  done ;;
profstep)
  ( timeout 600 python tools/prof_step.py bf16 ) > gpurun_out/prof_step_bf16.txt 2>&1
  echo "== prof_step rc=$?"; grep -v Warning gpurun_out/prof_step_bf16.txt | tail -30 ;;
gemmbench)
  ( timeout 600 python tools/gemm_bench.py ) > gpurun_out/gemm_bench.log 2>&1
  echo "== gemm_bench rc=$?"; cat gpurun_out/gemm_bench.log ;;
ncustats)
  ( timeout 300 python tools/gemm_bench.py last_fwd_student_bound_stats > gpurun_out/ncustats_plain.log 2>&1 ) &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm_stats \
      python tools/gemm_bench.py last_fwd_student_bound_stats > gpurun_out/ncustats.log 2>&1
  echo "== ncu stats gemm rc=$?"; cat gpurun_out/ncustats_plain.log; ls -la gpurun_out/prof_gemm_stats.ncu-rep ;;
ncugemm)
  ( timeout 300 python tools/gemm_bench.py last_fwd_student > gpurun_out/ncugemm_plain.log 2>&1 ) &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm_last_fwd \
      python tools/gemm_bench.py last_fwd_student > gpurun_out/ncugemm.log 2>&1
  echo "== ncu gemm rc=$?"; tail -n 5 gpurun_out/ncugemm.log; ls -la gpurun_out/*.ncu-rep ;;
ncufull)
  ( timeout 300 python bench.py --steps 2 --warmup 3 --graph 0 --overlap 0 --no-cpu-baseline > gpurun_out/ncufull_plain.log 2>&1 ) &&
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|ce_fused_kernel|ema_kernel|teacher_pass_kernel|weightnorm_fwd_kernel|weightnorm_bwd_kernel" -s 150 -c 24 -f -o /tmp/prof_step_kernels \
      python bench.py --steps 2 --warmup 3 --graph 0 --overlap 0 --no-cpu-baseline > gpurun_out/ncufull.log 2>&1
  echo "== ncu full rc=$?"; tail -n 3 gpurun_out/ncufull.log | cut -c1-300
  # the report is too big to travel back whole: export the raw page here, keep the report only if small
  ncu -i /tmp/prof_step_kernels.ncu-rep --page raw --csv > gpurun_out/prof_step_kernels_raw.csv 2>/dev/null
  ncu -i /tmp/prof_step_kernels.ncu-rep --page details --csv > gpurun_out/prof_step_kernels_details.csv 2>/dev/null
  sz=$(stat -c %s /tmp/prof_step_kernels.ncu-rep); echo "report bytes: $sz"
  if [ "$sz" -lt 40000000 ]; then cp /tmp/prof_step_kernels.ncu-rep gpurun_out/; fi
  ls -la gpurun_out/ | head -30 ;;
ncuce)
  ( timeout 300 python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncuce_plain.log 2>&1 ) &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"ce_fwd_kernel|ce_bwd_kernel|teacher_pass_kernel|ema_kernel|weightnorm_fwd_kernel" -s 25 -c 6 -f -o gpurun_out/prof_loss \
      python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncuce.log 2>&1
  echo "== ncu ce rc=$?"; tail -n 5 gpurun_out/ncuce.log; ls -la gpurun_out/*.ncu-rep ;;
ncu)
  ( timeout 600 python bench.py --steps 2 --warmup 3 --graph 0 --overlap 0 --no-cpu-baseline > gpurun_out/ncu_plain.log 2>&1 ) &&
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --graph 0 --overlap 0 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
  echo "== ncu launches rc=$?"; tail -n 3 gpurun_out/ncu.log; wc -l gpurun_out/launches.csv ;;
esac
done
