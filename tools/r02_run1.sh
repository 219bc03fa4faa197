#!/usr/bin/env bash
# round 2, GPU call 1: BASELINE-dimension parity tests + module tests (shadows, StepGraph schedules), bench with the CUPTI roofline, graph timeline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_baseline_dims test_gpu_modules test_gpu_kernels; do
  ( timeout 900 python -m pytest tests/$f.py -q -m gpu --tb=short -p no:cacheprovider ) > gpurun_out/r02_$f.log 2>&1
  echo "== $f rc=$?"; tail -n 25 gpurun_out/r02_$f.log
done
( timeout 600 python bench.py --steps 30 --warmup 5 ) > gpurun_out/r02_bench0.json 2> gpurun_out/r02_bench0.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/r02_bench0.json; tail -n 5 gpurun_out/r02_bench0.err
( DMC_TEACHER_SHADOWS=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r02_bench0_noshadow.json 2> gpurun_out/r02_bench0_noshadow.err
echo "bench noshadow rc=$?"; head -c 400 gpurun_out/r02_bench0_noshadow.json
( timeout 300 python tools/prof_step.py bf16 ) > gpurun_out/r02_prof0.txt 2>&1
echo "prof rc=$?"
( timeout 600 python bench.py --impl reference --steps 5 --warmup 2 ) > gpurun_out/r02_ref0.json 2> gpurun_out/r02_ref0.err
echo "ref rc=$?"; tail -c 1500 gpurun_out/r02_ref0.json
