#!/bin/bash
# Scratch command file for long gpurun invocations: `gpurun -- bash tools/_run.sh`.
# Default: the GPU parity suite and one short bench line.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --skip-config-legs 2>/dev/null | tail -1 | cut -c1-400
