"""Summarise an `ncu --page source --csv` dump: runs of SASS instructions with similar execution
counts (i.e. basic-block regions), their share of executed warp instructions and stall samples.
usage: python scripts/ncu_src_segments.py dump.csv [section_index] [min_share]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.005
sections, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        sections.append(cur)
    elif r and r[0] == "Address":
        hdr = r
    elif cur is not None and hdr is not None and len(r) == len(hdr):
        cur["rows"].append(r)
sec = sections[which]
data = sec["rows"]
isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
ith, isamp = hdr.index("Avg. Threads Executed"), hdr.index("# Samples")
tot = sum(int(r[iex]) for r in data)
tsamp = sum(int(r[isamp]) for r in data)
print(sec["name"], "| sass:", len(data), "| warp instr:", tot, "| samples:", tsamp, "| sections:", len(sections))
prev, start, segs = None, 0, []
for n, r in enumerate(data):
    ex = int(r[iex])
    if prev is not None and abs(ex - prev) > 0.02 * max(ex, prev, 1):
        segs.append((start, n - 1, prev))
        start = n
    prev = ex
segs.append((start, len(data) - 1, prev))
for s, e, c in segs:
    n = e - s + 1
    if n * c > min_share * tot:
        samp = sum(int(data[i][isamp]) for i in range(s, e + 1))
        print(f"sass[{s:4d}-{e:4d}] n={n:4d} count/inst={c:9d} instr-share={100*n*c/tot:5.1f}% "
              f"thr={data[s][ith]:>5s} stall-samples={100*samp/max(tsamp,1):5.1f}%  first: {data[s][isrc][:70]}")
if "--dump" in sys.argv:
    for n, r in enumerate(data):
        print(f"{n:5d} {int(r[iex]):10d} thr={r[ith]:>5s} samp={r[isamp]:>6s}  {r[isrc][:110]}")
