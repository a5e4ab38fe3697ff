"""Aggregate an `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` dump per CUDA source line:
samples, instructions executed, top stall reasons.  usage: python tools/ncu_lines.py dump.csv [top_n] [file_filter]"""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
flt = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
cur_file, hdr, agg = None, None, collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] in ("Function Name", "Kernel Name"):
        continue
    if r[0].strip().isdigit() and len(r) > 10:
        key = (cur_file, int(r[0]))
        d = agg.setdefault(key, {"src": r[1].strip()[:100], "samples": 0, "inst": 0, "stalls": collections.Counter()})
        ix = {h: i for i, h in enumerate(hdr)}
        def num(name):
            try:
                return int(float(r[ix[name]] or 0))
            except Exception:
                return 0
        # line-level row (Address empty) carries the aggregate
        if r[2] in ("", "-"):
            d["samples"] += num("# Samples")
            d["inst"] += num("Instructions Executed")
            for h in hdr:
                if h.startswith("stall_") and "Not Issued" not in h:
                    d["stalls"][h] += num(h)
tot = sum(d["samples"] for d in agg.values())
print("total samples", tot)
items = [(k, d) for k, d in agg.items() if not flt or flt in k[0]]
for (f, ln), d in sorted(items, key=lambda x: -x[1]["samples"])[:top]:
    st = ", ".join(f"{k[6:]} {v}" for k, v in d["stalls"].most_common(3) if v)
    print(f"{d['samples']:7d} {100*d['samples']/max(tot,1):5.1f}%  inst {d['inst']:9d}  {f}:{ln:<4d} {d['src'][:80]}  [{st}]")
