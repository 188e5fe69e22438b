timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "allgather or place" 2>&1 | tail -3 > gpurun_out/r02_2gpu_e.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>/dev/null > gpurun_out/r02_bench_2gpu.json ) 2>> gpurun_out/r02_2gpu_e.log
