"""Per-kernel counts of the SASS mnemonics that prove which hardware paths a kernel uses (B200_PROFILING.md):
tcgen05 MMAs (UTC*MMA), TMEM loads (LDTM), TMA tensor loads / stores (UTMALDG / UTMASTG), 1-D bulk copies (UBLKCP),
legacy tensor-core MMAs (HMMA), ldmatrix (LDSM).  usage: python tools/sass_mnemonics.py [lib.so] > profiles/rN_sass_mnemonics.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "spotv2net_b200/libspotv2_gat.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "HMMA", "LDSM", "SYNCS", "BAR.SYNC"]
cur, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for p in pats:
        if re.search(r"\b" + re.escape(p) + r"\b|\b" + re.escape(p) + r"\.", line):
            counts[cur][p] += 1
            break
demangle = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass", lib, "- mnemonic counts per kernel (static instruction counts)")
for (name, c), d in zip(counts.items(), demangle):
    if not any(c[p] for p in pats[:11]):
        continue
    short = d[:d.rfind(">(") + 1] if ">(" in d else re.sub(r"\(.*", "", d); short = short.replace("void ", "").replace("spotv2::", "").replace("(anonymous namespace)::", "")
    print(f"{short[:78]:78s} " + "  ".join(f"{p} {c[p]}" for p in pats if c[p]))
