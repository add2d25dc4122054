#!/bin/bash
# One GPU-box visit of round 2: synccheck, bench lines of every BASELINE config
# (config 3 with cpu_baseline + parity, 1/2 with the MCMC pair), the reference
# arm, the ncu launch list of the bench command, DRAM traffic of the forward
# launches, and one --set full capture of every kernel at a reduced size.
mkdir -p gpurun_out
TAG=${1:-r2}
SEL='golden_vectors or (generated_problems and (8-2000 or 6-1500 or 3-500 or 12-3000)) or checkpointed_small or prior_given'
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_$TAG.log 2>&1
tail -6 gpurun_out/pytest_$TAG.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_recomb.py tests/test_gpu_packed.py tests/test_gpu_infsites.py tests/test_gpu_totalprob.py -m gpu -q -k "$SEL or device_sampler_matches_reference and 8-3000 or packed_equals or infsites or totalprob" \
    > gpurun_out/sanitizer_synccheck.log 2>&1
echo "== synccheck rc=$?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_synccheck.log | tail -3
grep "at \|Device Frame" gpurun_out/sanitizer_synccheck.log | sort | uniq -c | sort -rn | head -8
timeout 900 python bench.py > gpurun_out/bench_${TAG}_config3.json 2> gpurun_out/bench_${TAG}_config3.err
for c in 1 2; do
  timeout 900 python bench.py --config $c > gpurun_out/bench_${TAG}_config$c.json 2> gpurun_out/bench_${TAG}_config$c.err
done
for c in 4 5; do
  timeout 900 python bench.py --config $c --cpu-sample-sites 50000 > gpurun_out/bench_${TAG}_config$c.json 2> gpurun_out/bench_${TAG}_config$c.err
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
for c in 3 1 2 4 5; do
python - <<PY
import json
try:
    l = json.load(open("gpurun_out/bench_${TAG}_config$c.json"))
    print("config $c value %.4e e2e %.4e ms %.1f frac %.4f fwd_only %.3e" % (l["value"], l["e2e"]["value"], l["ms_per_step"], l["roofline"]["frac"], l["forward_only"]["value"]), l["run"]["windows_per_gpu"], l.get("parity", {}).get("path_identical"), (l.get("cpu_baseline") or {}).get("value"), l.get("mcmc"))
except Exception as e:
    print("config $c failed", e)
PY
done
cat gpurun_out/bench_ref_$TAG.json | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none -k regex:awb_forward_fast -c 80 --csv \
    --log-file gpurun_out/traffic_fwd_$TAG.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_traffic_$TAG.log 2>&1
# one launch of every kernel, full sections: 444 windows x 20 k sites (three
# windows per SM: the dense forward kernel) and 148 windows (one per SM).  The
# reports are summarised HERE (gpurun_out/ only travels back up to 64 MiB).
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'awb_' -c 9 -o /tmp/full_dense_$TAG -f \
    python scripts/perf_probe.py --k 50 --sites 20000 --chains 444 --reps 1 \
    > gpurun_out/ncu_full_dense_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'awb_forward_fast|awb_emit|awb_traceback' -c 3 -o /tmp/full_one_$TAG -f \
    python scripts/perf_probe.py --k 50 --sites 20000 --chains 148 --reps 1 \
    > gpurun_out/ncu_full_one_$TAG.log 2>&1
for w in dense one; do
  ncu -i /tmp/full_${w}_$TAG.ncu-rep --page raw --csv > /tmp/full_${w}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py /tmp/full_${w}_raw.csv > gpurun_out/ncu_full_${w}_summary_$TAG.txt 2>&1
  ncu -i /tmp/full_${w}_$TAG.ncu-rep --page source --print-source cuda,sass --csv \
      -k regex:awb_forward_fast > /tmp/fwd_src_${w}.csv 2>/dev/null
  python scripts/ncu_roles.py /tmp/fwd_src_${w}.csv argweaver_b200/csrc/awb_forward_fast.cuh \
      > gpurun_out/forward_stalls_by_role_${w}_$TAG.txt 2>&1
  python scripts/ncu_lines.py /tmp/fwd_src_${w}.csv > gpurun_out/forward_hot_lines_${w}_$TAG.txt 2>&1
done
ls -la gpurun_out | tail -14; du -sh gpurun_out
