timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_i.log
for c in 0 48 96 192 384; do
  echo "chunk $c" >> gpurun_out/r02_i.log
  if [ $c = 0 ]; then unset RSPT_HOST_CHUNK_FRAMES; else export RSPT_HOST_CHUNK_FRAMES=$c; fi
  timeout 600 python bench.py --quick --no-cpu --steps 3 --warmup 3 --batches 4 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); x=j['packers']['xdelta_hzr']
print('value %.1f'%j['value'], 'e2e C %.2f (%.3f) D %.2f (%.3f)'%(x['compress']['e2e']['value'], x['compress']['e2e']['frac_of_copy_ceiling'], x['decompress']['e2e']['value'], x['decompress']['e2e']['frac_of_copy_ceiling']))" >> gpurun_out/r02_i.log
done
