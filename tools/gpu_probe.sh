#!/bin/bash
# usage: gpu_probe.sh <tag> [dbg values...]; conv GPU tests, fwd probe (unfused, fused + trace, ablations)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
tag=$1; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -5 > gpurun_out/${tag}_tests.log
: > gpurun_out/${tag}_probe.jsonl
X2GNN_FUSED=0 timeout 60 python tools/tile_probe.py >> gpurun_out/${tag}_probe.jsonl 2>gpurun_out/${tag}_err.log
X2GNN_TA_TRACE=gpurun_out/${tag}_trace.json timeout 60 python tools/tile_probe.py >> gpurun_out/${tag}_probe.jsonl 2>>gpurun_out/${tag}_err.log
for d in "$@"; do
  X2GNN_TA_DBG=$d timeout 40 python tools/tile_probe.py >> gpurun_out/${tag}_probe.jsonl 2>>gpurun_out/${tag}_err.log
done
cat gpurun_out/${tag}_tests.log gpurun_out/${tag}_probe.jsonl
tail -3 gpurun_out/${tag}_err.log
