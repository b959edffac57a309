for v in 0 9 12 16 0 9; do
  echo "# bps=$v" >> gpurun_out/k12.txt
  X2GNN_SRC_BPS=$v timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-train-step 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['roofline']['phase_ms_per_step']['attn_bwd_src'], l['segment_constant_edge_attr']['phase_ms_per_step']['attn_bwd_src'])" >> gpurun_out/k12.txt
done
X2GNN_SRC_BPS=9 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_large.py -m gpu -x -q > gpurun_out/t12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t12.log
tail -2 gpurun_out/t12.log
