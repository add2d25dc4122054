"""Hottest source lines of one kernel, from
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > X.csv`
(samples, global-memory sectors, instructions per CUDA-C line)."""
import csv
import sys


def main(path, top=45):
    out, name, h = [], "", None
    f = lambda x: int(x) if x.strip().isdigit() else 0
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            name = r[1].split("/")[-1]
        elif r[0] == "Line No":
            h = r
            iN, iG, iI = (h.index("# Samples"), h.index("L2 Theoretical Sectors Global"),
                          h.index("Instructions Executed"))
        elif h and r[0].strip().isdigit() and len(r) > iG:
            out.append((f(r[iN]), f(r[iG]), f(r[iI]), name + ":" + r[0], r[1].strip()[:84]))
    tot, totg, toti = (sum(d[i] for d in out) or 1 for i in range(3))
    print("samples %d, global sectors %d, warp instructions %d" % (tot, totg, toti))
    for d in sorted(out, reverse=True)[:top]:
        print("%5.1f%% samp %5.1f%% sect %5.1f%% inst  %-22s| %s"
              % (100.0 * d[0] / tot, 100.0 * d[1] / totg, 100.0 * d[2] / toti, d[3], d[4]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
