N=$1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/r02_f_bench_${N}gpu.err > gpurun_out/r02_f_bench_${N}gpu.json
