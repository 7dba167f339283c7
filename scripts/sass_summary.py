"""Opcode evidence from the shipped library: per kernel, how many tcgen05 / TMA / TMEM / legacy-MMA instructions its SASS holds
(cuobjdump -sass; mnemonics per /opt/skills/guides/B200_PROFILING.md: UTCHMMA = tcgen05.mma kind::f16/tf32, UTMALDG / UTMASTG = TMA
load / store, LDTM = tcgen05.ld, HMMA = mma.sync).  usage: python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "safediffcon_b200", "libsafediffcon_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
PAT = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA(?!\.2CTA)"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
       ("LDTM", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("MUFU", r"\bMUFU"), ("LDGSTS", r"\bLDGSTS"),
       ("ATOM/RED", r"\b(ATOMG|RED)\b"), ("DADD/DFMA/DMUL", r"\b(DADD|DFMA|DMUL)\b")]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur and "/*" in line:
        counts[cur]["instr"] += 1
        for name, pat in PAT:
            if re.search(pat, line):
                counts[cur][name] += 1
                total[name] += 1
print(f"cuobjdump -sass {os.path.basename(lib)}: {len(counts)} kernels, sm_100a")
print("library totals:", ", ".join(f"{k} x{v}" for k, v in total.items()))
print()
hdr = ["instr"] + [n for n, _ in PAT]
print(f"{'kernel':70s} " + " ".join(f"{h[:9]:>9s}" for h in hdr))
for fn, c in counts.items():
    name = re.sub(r"\(.*", "", demangle(fn))[:70]
    print(f"{name:70s} " + " ".join(f"{c.get(h, 0):9d}" for h in hdr))
