#!/bin/bash
# validation batch (one gpurun call, one GPU): GPU test suite, smoke, one short bench line
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --quick --no-cpu --steps 10 --warmup 3 2>/dev/null | cut -c1-200
