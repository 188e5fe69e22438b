timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_l.log
for t in 256 512 1024; do echo "inv threads $t" >> gpurun_out/r02_l.log; RSPT_INV_THREADS=$t timeout 300 python tools/stage_times.py 4096 2>&1 | cut -c1-75,150-240 >> gpurun_out/r02_l.log; done
