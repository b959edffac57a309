timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t13.log
timeout 300 python bench.py --no-cpu-baseline --no-train-step > gpurun_out/b13.json 2> gpurun_out/b13.err
tail -3 gpurun_out/t13.log
