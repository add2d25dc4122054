#!/bin/bash
AWB_LIB=scripts/abl/lib_sanitize.so timeout 600 compute-sanitizer --tool synccheck --print-limit 5 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden_vectors or (generated_problems and (8-2000 or 20-5000-20-False or 12-3000))" 2>&1 | grep "Barrier\|by thread\|Device Frame\|passed\|failed\|SUMMARY" | head -12
