#!/usr/bin/env bash
# round 2, GPU call 1: new BASELINE-dimension parity tests, bench with the CUPTI roofline, graph timeline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_baseline_dims.py -q -m gpu --tb=short -p no:cacheprovider -x ) > gpurun_out/r02_test_baseline_dims.log 2>&1
echo "baseline_dims rc=$?"; tail -n 30 gpurun_out/r02_test_baseline_dims.log
( timeout 600 python bench.py --steps 30 --warmup 5 ) > gpurun_out/r02_bench0.json 2> gpurun_out/r02_bench0.err
echo "bench rc=$?"; tail -c 5000 gpurun_out/r02_bench0.json; tail -n 5 gpurun_out/r02_bench0.err
( timeout 300 python tools/prof_step.py bf16 ) > gpurun_out/r02_prof0.txt 2>&1
echo "prof rc=$?"
( timeout 600 python bench.py --impl reference --steps 5 --warmup 2 ) > gpurun_out/r02_ref0.json 2> gpurun_out/r02_ref0.err
echo "ref rc=$?"; tail -c 1500 gpurun_out/r02_ref0.json
