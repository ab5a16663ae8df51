mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q -x --tb=short > gpurun_out/pytest_train.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_train.log
tail -60 gpurun_out/pytest_train.log
