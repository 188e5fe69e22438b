timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py 2>&1 | tail -3 > gpurun_out/r02_d3.log
for c in 0 96 192 296 444 592 1184 4096; do
  echo "chunk $c" >> gpurun_out/r02_d3.log
  RSPT_DECODE_CHUNK_FRAMES=$c python tools/stage_times.py 4096 2>&1 | head -3 | cut -c1-230 >> gpurun_out/r02_d3.log
done
