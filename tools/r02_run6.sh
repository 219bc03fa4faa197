#!/usr/bin/env bash
# round 2, GPU call 6 (2 GPUs): instrumented peer exchange (host timing, capture traceback)
mkdir -p gpurun_out
( DMC_XRANK_DEBUG=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 4 --warmup 3 --graph 0 --no-cpu-baseline ) > gpurun_out/r02f_peer_eager.json 2> gpurun_out/r02f_peer_eager.err
echo "eager rc=$?"; grep -E "\[xrank\]|\[reducer\]|\[step host" gpurun_out/r02f_peer_eager.err | head -40
( DMC_XRANK_DEBUG=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02f_peer_graph.json 2> gpurun_out/r02f_peer_graph.err
echo "graph rc=$?"; grep -E "\[xrank\]|\[reducer\]" gpurun_out/r02f_peer_graph.err | head -10; grep -B2 -A40 "capture failed" gpurun_out/r02f_peer_graph.err | grep -v "^\[rank1\]" | head -90 | cut -c1-220
echo done
