#!/usr/bin/env bash
# ncu --set full with source import of ONE launch of a microbenchmark kernel; exports the source/raw pages as CSV.
# usage: tools/ncu_src.sh <tag> <skip> <python cmd...>
tag=$1; skip=$2; shift 2
mkdir -p gpurun_out
( timeout 300 "$@" > gpurun_out/ncusrc_${tag}_plain.log 2>&1 ) || { echo "plain run failed"; tail -5 gpurun_out/ncusrc_${tag}_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s $skip -c 1 -f -o /tmp/ncusrc_$tag "$@" > gpurun_out/ncusrc_${tag}.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/ncusrc_$tag.ncu-rep --page source --csv > gpurun_out/ncusrc_${tag}_source.csv 2>/dev/null
ncu -i /tmp/ncusrc_$tag.ncu-rep --page raw --csv > gpurun_out/ncusrc_${tag}_raw.csv 2>/dev/null
ls -la gpurun_out/ncusrc_${tag}*
