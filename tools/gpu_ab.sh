#!/bin/bash
# usage: gpu_ab.sh <tag>: conv GPU tests, then the layer-step bench in fused (default) and unfused mode
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
tag=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -5 > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_tests.log
for m in tf32x3 tf32x3_fused; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-step --mode $m > gpurun_out/${tag}_bench_$m.json 2> gpurun_out/${tag}_bench_$m.err
done
python - <<PY
import json
for n in ("tf32x3","tf32x3_fused"):
    try:
        l=json.loads(open(f"gpurun_out/${tag}_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(l["ms_per_step"],4), round(l["value"]/1e6,1), round(l["roofline"]["frac"],4), l["roofline"].get("phase_ms_per_step"), "launches", l.get("gpu_launches_per_step"))
    except Exception as e:
        print(n, "ERR", e, open(f"gpurun_out/${tag}_bench_{n}.err").read()[-1500:])
PY
