timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-step > gpurun_out/ncu_ll.log 2>&1
