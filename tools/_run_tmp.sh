timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py 2>&1 | tail -5 > gpurun_out/r02_b1.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_b1_bench.json 2> gpurun_out/r02_b1_bench.err
tail -5 gpurun_out/r02_b1_bench.err >> gpurun_out/r02_b1.log
