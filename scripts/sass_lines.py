"""Static SASS instruction count per source line / region of one kernel, from
`nvdisasm -g -c <cubin>` (cubin: `cuobjdump -xelf all libargweaver_b200.so`).
Code that runs once per block is cold in the instruction caches; this shows
where a kernel's instructions are.  usage: sass_lines.py <nvdisasm.txt> <kernel-substring> [top]"""
import collections
import re
import sys


def main(path, kern, top=25):
    cnt = collections.Counter()
    cur, inside = None, False
    for line in open(path, errors="replace"):
        if line.startswith("//---") and ".text." in line:
            inside = kern in line
            cur = None
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+[A-Z@]", line) and cur:
            cnt[cur] += 1
    tot = sum(cnt.values())
    print("static SASS instructions: %d (%.1f KB)" % (tot, tot * 16 / 1024.0))
    files = collections.Counter()
    for (f, l), c in cnt.items():
        files[f] += c
    for f, c in files.most_common(6):
        print("  %5d  %s" % (c, f))
    for (f, l), c in cnt.most_common(top):
        print("%5d  %s:%d" % (c, f, l))
    return cnt


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
