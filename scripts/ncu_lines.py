#!/usr/bin/env python
"""Warp-stall samples of one kernel per CUDA source line.

  ncu -i rep.ncu-rep --page source --csv > src.csv          (SASS view: address, samples, stall columns)
  cuobjdump -xelf all libbseg.so ; nvdisasm -g -c X.cubin > X.sass   (SASS with //## File ... line markers)
  python scripts/ncu_lines.py src.csv X.sass <mangled-name-substring> [top]
"""
import csv
import re
import sys


def line_map(sass_path, func):
    off2line = {}
    inside = False
    cur = None
    for ln in open(sass_path, errors="replace"):
        if ln.startswith("//---") and ".text." in ln:
            inside = func in ln
            cur = None
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            inl = "inlined" in ln
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m and cur:
            off2line[int(m.group(1), 16)] = cur
    return off2line


def main():
    src, sass, func = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    o2l = line_map(sass, func)
    rows = list(csv.reader(open(src)))
    hk = next(k for k, r in enumerate(rows) if "# Samples" in r)
    hdr = rows[hk]
    ci = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    agg = {}
    for r in rows[hk + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        a = int(r[0], 16)
        if base is None:
            base = a
        key = o2l.get(a - base, ("?", 0))
        d = agg.setdefault(key, {"n": 0, "inst": 0})
        d["n"] += int(r[ci["# Samples"]] or 0)
        d["inst"] += int(r[ci["Instructions Executed"]] or 0)
        for s in stalls:
            v = int(r[ci[s]] or 0)
            if v:
                d[s] = d.get(s, 0) + v
    tot = sum(d["n"] for d in agg.values()) or 1
    print(f"total samples {tot}")
    for key, d in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top]:
        st = sorted(((v, k[6:]) for k, v in d.items() if k.startswith("stall_")), reverse=True)[:3]
        print(f"{100 * d['n'] / tot:5.1f}%  {key[0]}:{key[1]:<5d} inst {d['inst']:>8d}  " + ", ".join(f"{k} {v}" for v, k in st))


if __name__ == "__main__":
    main()
