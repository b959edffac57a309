timeout 300 python bench.py --no-cpu-baseline --no-train-step > gpurun_out/b16.json 2> gpurun_out/b16.err; tail -3 gpurun_out/b16.err
