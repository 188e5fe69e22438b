RSPT_DECODE_SEG_XOR=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_l.log
for t in 0 1; do echo "seg_xor $t" >> gpurun_out/r02_l.log; if [ $t = 1 ]; then export RSPT_DECODE_SEG_XOR=1; fi; timeout 300 python tools/stage_times.py 4096 2>&1 | cut -c1-75,150-240 >> gpurun_out/r02_l.log; done
