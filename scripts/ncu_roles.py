"""Stall reasons of the forward kernel by warp role (norm / scribes / compute), from an
`ncu --page source --csv --print-source cuda,sass` dump of awb_forward_fast_kernel."""
import collections
import csv
import sys


def main(path, src_path):
    rows = list(csv.reader(open(path)))
    hdr, start = None, 0
    for i, r in enumerate(rows[:10]):
        if r and r[0] == "Line No":
            hdr, start = r, i
    num = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
    data = [r for r in rows[start + 1:] if len(r) >= len(hdr) and r[0].strip().isdigit()]
    stall = [(i, h) for i, h in enumerate(hdr)
             if h.startswith("stall_") and "Not Issued" not in h]
    src = {i: l for i, l in enumerate(open(src_path), 1)}

    def find(marker):
        return next(i for i, l in src.items() if marker in l)

    book0 = find("who keeps the books")
    scr0 = find("F-scribes: per-time sums and R between barrier 1")
    comp0 = find("// compute warps")
    reg = collections.OrderedDict((k, collections.Counter())
                                  for k in ("bookkeeping (lambdas + its own warp)", "scribe warps", "compute warps",
                                            "inlined (shuffles, asm)"))
    smp = collections.Counter()
    for r in data:
        ln = int(r[0])
        infile = ln in src and src[ln].strip()[:20] == r[1].strip()[:20]
        key = ("inlined (shuffles, asm)" if not infile or ln < book0 else
               "bookkeeping (lambdas + its own warp)" if ln < scr0 else "scribe warps" if ln < comp0 else
               "compute warps")
        smp[key] += num(r[6])
        for i, h in stall:
            reg[key][h] += num(r[i])
    tot = sum(smp.values())
    for k in reg:
        print("%-26s %5.1f %% of warp samples" % (k, 100.0 * smp[k] / tot))
        for h, v in reg[k].most_common(6):
            print("      %-24s %5.1f %%" % (h, 100.0 * v / max(smp[k], 1)))
    print("\nhottest lines")
    ix = {h: i for i, h in enumerate(hdr)}
    for r in sorted(data, key=lambda r: -num(r[6]))[:16]:
        top = sorted(((num(r[i]), h) for i, h in stall), reverse=True)[:2]
        print("%5s %5.2f %%  %-44s %s" % (r[0], 100.0 * num(r[6]) / tot,
                                          ", ".join("%s %d" % (h[6:], v) for v, h in top),
                                          r[1].strip()[:60]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
