mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "sfm" -s 2>&1 | tail -8
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
