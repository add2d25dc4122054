#!/bin/bash
for sel in "generated_problems and 20-5000-20-False" "generated_problems and 8-2000-20-False" "generated_problems and 50-4000"; do
echo "== $sel"
timeout 300 compute-sanitizer --tool synccheck --print-limit 2 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$sel" 2>&1 | grep "Barrier\|by thread\|Device Frame\|passed\|failed\|SUMMARY" | head -8
done
echo "== AWB_VERBOSE shapes"
AWB_VERBOSE=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "golden_vectors and T30" 2>&1 | grep "forward kernel"
AWB_VERBOSE=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "generated_problems and 20-5000-20-False" 2>&1 | grep "forward kernel"
