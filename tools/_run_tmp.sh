N=8
for c in 0 1024; do
  if [ $c = 0 ]; then unset RSPT_HOST_CHUNK_FRAMES; else export RSPT_HOST_CHUNK_FRAMES=$c; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --quick --no-cpu 2>/dev/null > gpurun_out/r02_e2e_${N}gpu_chunk$c.json
done
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1; lscpu | head -20 >> gpurun_out/r02_topo.txt
