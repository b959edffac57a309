timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t10.log
timeout 300 python bench.py --no-cpu-baseline --no-train-step > gpurun_out/b10.json 2> gpurun_out/b10.err
timeout 300 python bench.py --no-cpu-baseline --no-train-step --mode tf32 > gpurun_out/b10_tf32.json 2> gpurun_out/b10_tf32.err
X2GNN_CPASYNC=0 timeout 300 python bench.py --no-cpu-baseline --no-train-step > gpurun_out/b10_old.json 2> gpurun_out/b10_old.err
tail -3 gpurun_out/t10.log
