#!/usr/bin/env bash
# round 2, final 1-GPU evidence: the whole GPU test suite, smoke, bench (bf16 + fp32 + reference arm), timeline, ncu launch list and
# one ncu --set full capture of the step's main kernels (exported as CSV pages; clocks recorded beside them).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv > gpurun_out/r02z_gpu.txt 2>&1
( timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider ) > gpurun_out/r02z_tests.log 2>&1
echo "== pytest -m gpu rc=$?"; tail -n 5 gpurun_out/r02z_tests.log | cut -c1-250
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r02z_smoke.log 2>&1; echo "== smoke rc=$?"; tail -n 3 gpurun_out/r02z_smoke.log
( timeout 600 python bench.py --steps 50 --warmup 10 ) > gpurun_out/r02z_bench_bf16.json 2> gpurun_out/r02z_bench_bf16.err; echo "== bench rc=$?"; head -c 700 gpurun_out/r02z_bench_bf16.json; echo
( timeout 600 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/r02z_bench_ref.json 2> gpurun_out/r02z_bench_ref.err; echo "== ref rc=$?"; head -c 300 gpurun_out/r02z_bench_ref.json; echo
( timeout 300 python tools/prof_step.py bf16 ) > gpurun_out/r02z_timeline.txt 2>&1
( timeout 300 python bench.py --steps 2 --warmup 3 --graph 0 --overlap 0 --no-cpu-baseline > gpurun_out/r02z_ncu_plain.log 2>&1 ) &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches.csv \
    python bench.py --steps 2 --warmup 3 --graph 0 --overlap 0 --no-cpu-baseline > gpurun_out/r02z_ncu.log 2>&1
echo "== ncu launches rc=$?"; wc -l gpurun_out/r02z_launches.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|ce_fused_kernel|ema2_kernel|teacher_pass_kernel|weightnorm_fwd_kernel|weightnorm_bwd_kernel|splitk_normalize_bwd_kernel" -s 160 -c 26 -f -o /tmp/r02z_step_kernels \
    python bench.py --steps 2 --warmup 3 --graph 0 --overlap 0 --no-cpu-baseline > gpurun_out/r02z_ncufull.log 2>&1
echo "== ncu full rc=$?"
ncu -i /tmp/r02z_step_kernels.ncu-rep --page raw --csv > gpurun_out/r02z_step_kernels_raw.csv 2>/dev/null
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv >> gpurun_out/r02z_gpu.txt 2>&1
ls -la gpurun_out/r02z_* | head -20
echo done
