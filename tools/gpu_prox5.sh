#!/bin/bash
# Developer tool: proximity tests on the GPU (3D, generated kernels, planar models), short.
bash tools/gpu_health.sh || exit 0
timeout 600 python -m pytest tests/test_proximity.py -m gpu -x -q -o faulthandler_timeout=150 2>&1 | tail -8
