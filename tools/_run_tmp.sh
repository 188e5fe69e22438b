timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_h.log
for r in 16 32 64 96 128; do echo "min_run $r" >> gpurun_out/r02_h.log; RSPT_PAIR_MIN_RUN=$r timeout 300 python tools/stage_times.py 4096 2>&1 | grep -v xdelta >> gpurun_out/r02_h.log; done
