RSPT_PYTEST_ARGS="--deselect tests/test_gpu_scale.py" bash tools/exp_run.sh > gpurun_out/r02_i4.log 2>&1
RSPT_INV_MODE=0 python tools/stage_times.py 4096 2>&1 | head -2 >> gpurun_out/r02_i4.log
