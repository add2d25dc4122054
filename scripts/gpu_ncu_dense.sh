#!/bin/bash
# --set full capture of the dense forward kernel (three windows per SM), the
# traceback and the emission kernel; summarised on the box
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'awb_forward_fast|awb_traceback|awb_emit' -c 3 -o /tmp/full_dense -f \
    python scripts/perf_probe.py --k 50 --sites 20000 --chains 444 --reps 1 --packed 1 \
    > gpurun_out/ncu_full_dense.log 2>&1
ncu -i /tmp/full_dense.ncu-rep --page raw --csv > /tmp/full_dense_raw.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/full_dense_raw.csv > gpurun_out/ncu_full_dense_k345_summary.txt 2>&1
ncu -i /tmp/full_dense.ncu-rep --page source --print-source cuda,sass --csv \
    -k regex:awb_forward_fast > /tmp/fwd_src_dense.csv 2>/dev/null
python scripts/ncu_roles.py /tmp/fwd_src_dense.csv argweaver_b200/csrc/awb_forward_fast.cuh \
    > gpurun_out/forward_stalls_by_role_dense.txt 2>&1
python scripts/ncu_lines.py /tmp/fwd_src_dense.csv > gpurun_out/forward_hot_lines_dense.txt 2>&1
grep -B2 -A22 "forward_fast" gpurun_out/ncu_full_dense_k345_summary.txt | head -40
cat gpurun_out/forward_stalls_by_role_dense.txt | head -30
