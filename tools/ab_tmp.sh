timeout 600 python -m pytest tests/test_gpu_bases.py tests/test_gpu_model.py tests/test_gpu_dropin.py -q -m gpu -x 2>&1 | tail -2 | cut -c1-300
timeout 300 python tools/train_step_ab.py 2>/dev/null | tail -1 | cut -c1-120
