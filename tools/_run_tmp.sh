timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_h.log
timeout 300 python tools/stage_times.py 4096 >> gpurun_out/r02_h.log 2>&1
timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/r02_h.log
