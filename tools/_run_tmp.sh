RSPT_PYTEST_ARGS="--deselect tests/test_gpu_scale.py" bash tools/exp_run.sh > gpurun_out/r02_d2.log 2>&1
