#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2f}
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_$TAG.log 2>&1
tail -25 gpurun_out/pytest_$TAG.log
