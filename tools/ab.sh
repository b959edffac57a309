timeout 600 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/t14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t14.log
python tools/prof_train.py > gpurun_out/prof_train2.txt 2>&1
git stash -q 2>/dev/null
tail -3 gpurun_out/t14.log; tail -1 gpurun_out/prof_train2.txt
