RSPT_TREE_LS=16 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_j.log
for w in 0 8 16 24 32; do echo "tree_ls $w" >> gpurun_out/r02_j.log; RSPT_TREE_LS=$w timeout 300 python tools/stage_times.py 4096 2>&1 | cut -c1-75,110-240 >> gpurun_out/r02_j.log; done
