"""Copy the outputs of one `gpu_round.sh <tag>` visit from gpurun_out/ into profiles/
(bench lines, launch list, DRAM traffic of the forward launches, ncu summaries)."""
import collections
import csv
import json
import shutil
import subprocess
import sys

tag = sys.argv[1]
G = "gpurun_out/"
U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}
rows = [r for r in csv.reader(open(G + "traffic_fwd_%s.csv" % tag)) if len(r) > 14 and r[0].isdigit()]
by = collections.defaultdict(dict)
for r in rows:
    by[int(r[0])][r[12]] = float(r[14].replace(",", "")) * U[r[13]]
ids = sorted(by)
bench = json.load(open(G + "bench_%s.json" % tag))
nseg = len(ids) // 2
p0, p1 = ids[:nseg], ids[nseg:]
tot = lambda ii, k: sum(by[i][k] for i in ii)
short = [i for i in p1 if by[i]["gpu__time_duration.sum"] < 1e5]
old = json.load(open("profiles/r1_forward_traffic.json"))
old.update({
    "launches_summed": nseg,
    "note": "sum over the %d forward-kernel launches of the first pass of one step (one per "
            "segment); the second pass repeats them except for the last segments of every "
            "window, whose tables stay resident (%d of its %d launches return at once; its "
            "launches took %.1f ms and wrote %.1f GB)"
            % (nseg, len(short), len(p1), tot(p1, "gpu__time_duration.sum") / 1e6,
               tot(p1, "dram__bytes_write.sum") / 1e9),
    "dram_bytes_read": int(tot(p0, "dram__bytes_read.sum")),
    "dram_bytes_write": int(tot(p0, "dram__bytes_write.sum")),
    "kernel_time_ms_under_ncu": tot(p0, "gpu__time_duration.sum") / 1e6,
    "algorithmic_bytes": int(bench["roofline"]["algorithmic_bytes_per_launch"]),
})
json.dump(old, open("profiles/r1_forward_traffic.json", "w"), indent=2)
shutil.copy(G + "traffic_fwd_%s.csv" % tag, "profiles/r1_forward_dram_traffic.csv")
shutil.copy(G + "launches_%s.csv" % tag, "profiles/r1_launches_bench.csv")
shutil.copy(G + "bench_%s.json" % tag, "profiles/r1_bench_n1.json")
shutil.copy(G + "bench_ref_%s.json" % tag, "profiles/r1_bench_reference_arm.json")
raw = subprocess.run(["ncu", "-i", G + "full_%s.ncu-rep" % tag, "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
open("/tmp/full_raw.csv", "w").write(raw)
open("profiles/r1_ncu_full_summary.txt", "w").write(
    subprocess.run([sys.executable, "scripts/ncu_summary.py", "/tmp/full_raw.csv"],
                   capture_output=True, text=True).stdout)
src = subprocess.run(["ncu", "-i", G + "full_%s.ncu-rep" % tag, "--page", "source",
                      "--print-source", "cuda,sass", "--csv", "-k", "regex:awb_forward_fast"],
                     capture_output=True, text=True).stdout
open("/tmp/fwd_src.csv", "w").write(src)
open("profiles/r1_forward_stalls_by_role.txt", "w").write(
    subprocess.run([sys.executable, "scripts/ncu_roles.py", "/tmp/fwd_src.csv",
                    "argweaver_b200/csrc/awb_forward_fast.cuh"], capture_output=True, text=True).stdout)
tot_k, cnt_k = collections.Counter(), collections.Counter()
for r in csv.reader(open("profiles/r1_launches_bench.csv")):
    if len(r) > 14 and r[0].isdigit():
        n = r[4].split("(")[0][:45]
        tot_k[n] += float(r[14].replace(",", "")) / 1e6
        cnt_k[n] += 1
s = sum(tot_k.values())
for n, t in tot_k.most_common():
    print("%-46s %4d launches %9.1f ms %5.1f%%" % (n, cnt_k[n], t, 100 * t / s))
print({k: bench[k] for k in ("value", "ms_per_step", "stage_ms")}, bench["e2e"]["value"],
      bench["e2e"]["ms_per_step"], bench["roofline"]["frac"])
