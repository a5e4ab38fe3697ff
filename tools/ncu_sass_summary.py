"""Summarise `ncu --page source --csv` output: instruction mix, stall reasons, hottest SASS ranges."""
import csv, sys, collections
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); stalls = collections.Counter(); total = 0; samples = 0
recs = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    op = r[ix["Source"]].split()
    op = [t for t in op if not t.startswith("@")][0] if op else "?"
    ops[op.split(".")[0]] += n; total += n; samples += s
    for h in hdr:
        if h.startswith("stall_") and "(Not Issued)" not in h:
            stalls[h] += int(r[ix[h]] or 0)
    recs.append((n, s, r[ix["Source"]].strip(), int(r[ix["L1 Wavefronts Shared"]] or 0), int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)))
print("total warp instructions", total, "samples", samples)
print("top opcodes:", [(k, f"{100*v/total:.1f}%") for k, v in ops.most_common(14)])
print("stalls:", [(k, f"{100*v/max(1,samples):.1f}%") for k, v in stalls.most_common(8)])
# hottest contiguous windows of 40 instructions by samples
W = 48
best = sorted(((sum(r[1] for r in recs[i:i+W]), i) for i in range(0, len(recs), W)), reverse=True)[:6]
for s, i in best:
    n = sum(r[0] for r in recs[i:i+W])
    print(f"--- window @{i}: samples {100*s/max(1,samples):.1f}% insts {100*n/total:.1f}%  e.g.:", "; ".join(r[2][:28] for r in recs[i:i+W:8]))
sh = sum(r[3] for r in recs); shi = sum(r[4] for r in recs)
print("shared wavefronts", sh, "ideal", shi)
