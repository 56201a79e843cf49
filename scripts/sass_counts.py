"""Per-kernel counts of the SASS mnemonics that show the Blackwell-native paths (tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM,
TMA tensor loads -> UTMALDG, 1-D bulk copies -> UBLKCP, cp.async -> LDGSTS) in the built library.
usage: python scripts/sass_counts.py [lib] > profiles/r02_sass_counts.txt   (runs on the CPU box: cuobjdump only)"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "aura_snn_rag_b200/libaura_hippo.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pat = {"UTCHMMA": r"\bUTCHMMA\b(?!\.2CTA)", "UTCHMMA.2CTA": r"UTCHMMA\.2CTA", "LDTM": r"\bLDTM", "UTMALDG": r"\bUTMALDG",
       "UBLKCP": r"\bUBLKCP", "LDGSTS": r"\bLDGSTS", "SYNCS(mbarrier)": r"\bSYNCS", "ATOMS": r"\bATOMS", "REDUX": r"\bREDUX"}
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pat.items():
            if re.search(p, line):
                counts[cur][k] += 1
cols = list(pat)
print(f"{'kernel':78s} " + " ".join(f"{c:>14s}" for c in cols))
tot = collections.Counter()
for name, c in counts.items():
    if sum(c.values()) == 0:
        continue
    tot.update(c)
    print(f"{name[:78]:78s} " + " ".join(f"{c[k]:14d}" for k in cols))
print(f"{'TOTAL':78s} " + " ".join(f"{tot[k]:14d}" for k in cols))
