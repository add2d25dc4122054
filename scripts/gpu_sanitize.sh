#!/bin/bash
# compute-sanitizer over the parity suite at reduced sizes
mkdir -p gpurun_out
SEL='golden_vectors or (generated_problems and (8-2000 or 6-1500 or 3-500 or 12-3000)) or checkpointed_small or prior_given'
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py tests/test_recomb.py tests/test_gpu_packed.py tests/test_gpu_infsites.py tests/test_gpu_totalprob.py -m gpu -x -q -k "$SEL or device_sampler_matches_reference and 8-3000 or packed_equals or infsites or totalprob" \
      > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"
  grep -c "========= " gpurun_out/sanitizer_$tool.log
  grep "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" gpurun_out/sanitizer_$tool.log | tail -4
done
