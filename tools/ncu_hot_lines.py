"""Per-source-line executed-instruction profile of an .ncu-rep captured with --import-source on.
Usage: python tools/ncu_hot_lines.py sass.csv disasm.txt source.cuh
  sass.csv   = ncu -i rep --page source --csv --print-source sass
  disasm.txt = nvdisasm -g -c <cubin of the same build>"""
import collections
import csv
import re
import sys


def main():
    sass, disasm, srcfile = sys.argv[1:4]
    func, cur, maps = None, None, {}
    for ln in open(disasm).read().split("\n"):
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            func, cur = m.group(1), None
            maps[func] = {}
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m and cur and func:
            maps[func][int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(sass)))
    secs, i = [], 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            j, body = i + 2, []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                body.append(rows[j])
                j += 1
            secs.append((rows[i][1], rows[i + 1], body))
            i = j
        else:
            i += 1
    src = open(srcfile).read().split("\n")
    seen = set()
    for name, hdr, body in secs:
        if name in seen:
            continue
        seen.add(name)
        short = re.sub(r"\(.*", "", name).split("::")[-1]
        key = [k for k in maps if short in k]
        if not key:
            continue
        mp = maps[key[0]]
        ia, ii, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        base = int(body[0][ia], 16)
        agg, aggt, tot = collections.Counter(), collections.Counter(), 0
        for r in body:
            if len(r) <= it:
                continue
            k = mp.get(int(r[ia], 16) - base, ("?", 0))
            n = int(r[ii])
            agg[k] += n
            aggt[k] += int(r[it])
            tot += n
        print("=====", short, "warp instructions", tot)
        for k, n in agg.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 18):
            txt = src[k[1] - 1].strip()[:100] if k[0] == srcfile.split("/")[-1] else k[0]
            print("%5d %5.1f%%  thr/inst %4.1f  %s" % (k[1], 100 * n / tot, aggt[k] / max(n, 1), txt))


if __name__ == "__main__":
    main()
