#!/bin/bash
# usage: gpu_full.sh <tag>: whole GPU test suite, the default bench line, ncu launch lists of one layer step (both modes)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
tag=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_tests.log | tail -6
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 600 gpurun_out/${tag}_bench.err
for m in tf32x3 tf32x3_fused; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${tag}_launches_$m.csv python bench.py --layer-only --steps 2 --warmup 3 --mode $m > gpurun_out/${tag}_ncu_$m.log 2>&1
done
python - <<PY
import json
l=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
keep={k:l.get(k) for k in ("value","ms_per_step","e2e","gpu_launches_per_step","dp_grad_max_rel_err")}
keep["frac"]=l["roofline"]["frac"]; keep["phases"]=l["roofline"]["phase_ms_per_step"]; keep["traffic_source"]=l["roofline"].get("traffic_source")
print(json.dumps(keep))
for k in ("e2e_from_atoms","reference_on_gpu","ocelot_inference","ball500_sweep","segment_constant_edge_attr"):
    print(k, json.dumps(l.get(k))[:700])
t=l.get("train_step") or {}
print("train", {k:t.get(k) for k in ("mode","molecules_per_sec","ms_per_step","ms_per_step_median","error")})
PY
