#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2g}
timeout 1200 python -m pytest tests/test_recomb.py tests/test_gpu_dropin.py tests/test_gpu_infsites.py -m gpu -q --durations=5 > gpurun_out/pytest_$TAG.log 2>&1
tail -25 gpurun_out/pytest_$TAG.log
