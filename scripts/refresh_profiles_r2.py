"""Copy the outputs of one `gpu_round2.sh <tag>` visit from gpurun_out/ into profiles/
(bench lines of every config, the reference arm, launch list, DRAM traffic of the
forward launches, ncu summaries, sanitizer summaries, SASS listing of the kernels)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
G = "gpurun_out/"
P = "profiles/"
U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6,
     "s": 1e9, "usecond": 1e3, "msecond": 1e6, "nsecond": 1, "second": 1e9}


def cp(src, dst):
    if os.path.exists(G + src):
        shutil.copy(G + src, P + dst)
        return True
    print("missing", src)
    return False


for c in (1, 2, 3, 4, 5):
    cp("bench_%s_config%d.json" % (tag, c),
       "r2_bench_n1.json" if c == 3 else "r2_bench_config%d.json" % c)
cp("bench_ref_%s.json" % tag, "r2_bench_reference_arm.json")
cp("launches_%s.csv" % tag, "r2_launches_bench.csv")
cp("traffic_fwd_%s.csv" % tag, "r2_forward_dram_traffic.csv")
for w in ("dense", "one"):
    cp("ncu_full_%s_summary_%s.txt" % (w, tag), "r2_ncu_full_summary_%s.txt" % w)
    cp("forward_stalls_by_role_%s_%s.txt" % (w, tag), "r2_forward_stalls_by_role_%s.txt" % w)
    cp("forward_hot_lines_%s_%s.txt" % (w, tag), "r2_forward_hot_lines_%s.txt" % w)

# launch list: time per kernel
tot_k, cnt_k = collections.Counter(), collections.Counter()
rows = list(csv.reader(open(P + "r2_launches_bench.csv")))
hdr = next(r for r in rows if r and r[0] == "ID")
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rows:
    if len(r) > vi and r[0].isdigit():
        n = re.sub(r"\(.*", "", r[ki])[:60]
        tot_k[n] += float(r[vi].replace(",", "")) * U.get(r[ui], 1) / 1e6
        cnt_k[n] += 1
s = sum(tot_k.values())
with open(P + "r2_launches_bench_summary.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none of "
            "`python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e` "
            "(two steps; per-launch times are serialised and cold-cache: shares, not "
            "absolutes)\n")
    for n, t in tot_k.most_common():
        line = "%-62s %5d launches %9.1f ms %5.1f %%" % (n, cnt_k[n], t, 100 * t / s)
        print(line)
        f.write(line + "\n")

# DRAM traffic of the forward launches of one step
rows = list(csv.reader(open(P + "r2_forward_dram_traffic.csv")))
hdr = next(r for r in rows if r and r[0] == "ID")
ii, ki, mi, ui, vi = (hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"),
                      hdr.index("Metric Unit"), hdr.index("Metric Value"))
by = collections.defaultdict(dict)
name = {}
for r in rows:
    if len(r) > vi and r[0].isdigit():
        by[int(r[ii])][r[mi]] = float(r[vi].replace(",", "")) * U.get(r[ui], 1)
        name[int(r[ii])] = re.sub(r"\(.*", "", r[ki])
bench = json.load(open(P + "r2_bench_n1.json"))
tot = lambda k: sum(v.get(k, 0) for v in by.values())
traffic = {
    "config": {"k": bench["config"]["k"], "ntimes": bench["config"]["ntimes"],
               "sites_per_window": bench["config"]["sites_per_window"],
               "windows_per_gpu": bench["run"]["windows_per_gpu"], "checkpoint": 1},
    "launches_summed": len(by),
    "kernels": sorted(set(name.values())),
    "dram_bytes_read": int(tot("dram__bytes_read.sum")),
    "dram_bytes_write": int(tot("dram__bytes_write.sum")),
    "kernel_time_ms_under_ncu": tot("gpu__time_duration.sum") / 1e6,
    "algorithmic_bytes": int(bench["roofline"]["algorithmic_bytes_per_step"]),
    "note": "sum over the forward-kernel launches of ONE step (first pass: one launch per "
            "segment; second pass: one launch per group of rebuilt segments), ncu "
            "--metrics dram__bytes_read.sum,dram__bytes_write.sum",
}
json.dump(traffic, open(P + "r2_forward_traffic.json", "w"), indent=2)
print("forward DRAM traffic: read %.1f GB, write %.1f GB, algorithmic %.1f GB (x%.2f)"
      % (traffic["dram_bytes_read"] / 1e9, traffic["dram_bytes_write"] / 1e9,
         traffic["algorithmic_bytes"] / 1e9,
         (traffic["dram_bytes_read"] + traffic["dram_bytes_write"]) /
         max(traffic["algorithmic_bytes"], 1)))

# (profiles/r2_compute_sanitizer.txt is written by hand from the logs of
# scripts/gpu_sanitize.sh and scripts/sync_probe.cu: it carries the probe's outcomes)

# SASS of the shipped kernels: the mnemonics that matter (bulk copies, mbarrier)
sass = subprocess.run(["cuobjdump", "-sass", "argweaver_b200/csrc/libargweaver_b200.so"],
                      capture_output=True, text=True).stdout
cur, ops = None, collections.defaultdict(collections.Counter)
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True,
                             text=True).stdout.strip().split("(")[0]
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", l)
    if m and cur:
        ops[cur][m.group(1).split(".")[0]] += 1
with open(P + "r2_sass_mnemonics.txt", "w") as f:
    f.write("cuobjdump -sass libargweaver_b200.so: instructions per kernel, and the "
            "counts of the mnemonics of interest\n")
    for k in sorted(ops):
        c = ops[k]
        f.write("%-70s %6d instr  DFMA %4d DADD %4d DMUL %4d SHFL %4d LDS %4d STS %4d BAR %3d "
                "UBLKCP %3d SYNCS %3d LDGSTS %3d\n"
                % (k[:70], sum(c.values()), c["DFMA"], c["DADD"], c["DMUL"], c["SHFL"],
                   c["LDS"], c["STS"], c["BAR"], c["UBLKCP"], c["SYNCS"], c["LDGSTS"]))
print(open(P + "r2_sass_mnemonics.txt").read()[:1500])
