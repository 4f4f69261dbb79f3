"""Per-kernel SASS evidence that the tensor-core kernels are tcgen05 / TMEM / TMA code (runs without a GPU):
cuobjdump -sass of the in-tree libsnvc_b200.so, mnemonic counts per kernel + the first instance of each.
    python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "snvc_b200", "libsnvc_b200.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS", "FFMA2", "HFMA2.BF16", "LDG.E.ENL2.256", "STG.E.ENL2.256"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, stats, first = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"snvc::\(anonymous namespace\)::", "", name)
        cur = name.split("(")[0]
        stats[cur] = collections.Counter(); first[cur] = {}
        continue
    if cur is None or "/*" not in line:
        continue
    ins = re.sub(r"/\*.*?\*/", "", line).strip()
    if not ins:
        continue
    stats[cur]["total"] += 1
    for k in KEYS:
        if re.search(r"\b" + re.escape(k) + r"\b", ins) or (k.endswith("256") and k in ins):
            if k == "UTCHMMA" and "UTCHMMA.2CTA" in ins:
                continue
            stats[cur][k] += 1
            first[cur].setdefault(k, ins[:110])
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a)  -- mnemonic counts per kernel")
tot = collections.Counter()
for k, c in stats.items():
    hits = {a: b for a, b in c.items() if a != "total"}
    tot.update(hits)
    if not any(x in hits for x in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM")):
        continue
    print(f"\n{k}\n    {c['total']} instructions; " + ", ".join(f"{a} x{b}" for a, b in sorted(hits.items())))
    for a in ("UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        if a in first[k]:
            print(f"    e.g. {first[k][a]}")
print("\n# whole library: " + ", ".join(f"{a} x{b}" for a, b in sorted(tot.items())))
print("\n# kernels without tensor-core / TMA instructions (HBM-bound gather / elementwise kernels):")
for k, c in stats.items():
    hits = {a: b for a, b in c.items() if a != "total"}
    if not any(x in hits for x in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM")):
        print(f"    {k[:100]}: {c['total']} instr" + ("; " + ", ".join(f"{a} x{b}" for a, b in sorted(hits.items())) if hits else ""))
