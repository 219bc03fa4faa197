#!/usr/bin/env bash
# One gpurun call: diagnostics + the GPU test suite, each step bounded by its own timeout.
# Everything worth reading lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 900 python tools/diag_gemm.py ) > gpurun_out/diag_gemm.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag_gemm.log
for f in test_gpu_gemm test_gpu_kernels test_gpu_modules; do
  ( timeout 900 python -m pytest tests/$f.py -q -m gpu --tb=short -p no:cacheprovider ) > gpurun_out/$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/$f.log
done
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -n 30 gpurun_out/diag_gemm.log
for f in test_gpu_gemm test_gpu_kernels test_gpu_modules smoke; do echo "== $f"; tail -n 15 gpurun_out/$f.log; done
