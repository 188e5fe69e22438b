timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py 2>&1 | tail -15 > gpurun_out/r02_t1.log
