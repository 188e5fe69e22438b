#!/bin/bash
# experiment driver (one gpurun call): default paths, then the opt-in spectral paths
set -o pipefail
echo "== default: parity + stress"; python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -m gpu -x -q 2>&1 | tail -3
echo "== default: stage times"; python tools/stage_times.py 4096
export RSPT_FWHT_FUSED=1 RSPT_WORDS_G4=1
echo "== opt-in: hadamard/dct parity"; python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_scale.py -m gpu -x -q -k "hadamard or dct or stream or golden or random or prdn" 2>&1 | tail -3
echo "== opt-in: stage times"; python tools/stage_times.py 4096 | grep "hadamard\|dct"
