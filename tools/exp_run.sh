#!/bin/bash
# validation batch (one gpurun call, one GPU): GPU test suite, smoke, one short bench line
timeout 900 python -m pytest tests -m gpu -x -q ${RSPT_PYTEST_ARGS:-} 2>&1 | tail -15
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --quick --no-cpu --steps 3 --warmup 3 --batches 4 2>/dev/null | cut -c1-250
timeout 300 python tools/stage_times.py 4096 2>&1 | tail -4
