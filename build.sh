#!/usr/bin/env bash
# Builds libdinomc.so (all CUDA kernels + the C ABI) for sm_100a, in-tree.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")" && pwd)"
PKG="$ROOT/self-supervised-learning-for-aerial-image-segmentation_b200"
SRC="$PKG/csrc"
OUT="$PKG/libdinomc.so"
BUILD="$PKG/build"
EXTRA=()
if [[ -n "${DMC_VARIANT:-}" ]]; then    # experiment build: DMC_VARIANT=name DMC_VARIANT_FLAGS="-D..." -> libdinomc_<name>.so (load with DMC_LIB=)
  OUT="$PKG/libdinomc_${DMC_VARIANT}.so"; BUILD="$PKG/build_${DMC_VARIANT}"; EXTRA=(${DMC_VARIANT_FLAGS:-})
fi
if [[ -n "${DMC_TRACE:-}" ]]; then      # debug build with the GEMM pipeline trace compiled in (tools/gemm_trace.py)
  OUT="$PKG/libdinomc_trace.so"; BUILD="$PKG/build_trace"; EXTRA=(-DDMC_GEMM_TRACE_BUILD)
fi
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       -I"$ROOT/include" -I"$SRC" --expt-relaxed-constexpr)
mkdir -p "$BUILD"
objs=()
pids=()
for f in api gemm_sm100 gemm_simt rowops teacher ce ema clip adamw lars xrank; do
  o="$BUILD/$f.o"
  objs+=("$o")
  if [[ ! -f "$o" || "$SRC/$f.cu" -nt "$o" || "$SRC/dmc_common.cuh" -nt "$o" || "$SRC/dmc_ptx.cuh" -nt "$o" || "$ROOT/include/dinomc.h" -nt "$o" ]]; then
    "$NVCC" "${FLAGS[@]}" "${EXTRA[@]}" ${DMC_PTXAS_V:+-Xptxas -v} -c "$SRC/$f.cu" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -shared -o "$OUT" "${objs[@]}" -gencode arch=compute_100a,code=sm_100a -cudart static
echo "built $OUT"
